"""CUDA-graph cache for the K-step consistency loops of the product path.

The reference makes ONE library call per step (`model.unet(...)`, utils/generation.py:241-244); here a U-Net forward is
~350 (SD1.5) / ~920 (SDXL) kernel launches issued from Python, so at small batch (BASELINE configs[0]: 1 row,
configs[2]: 2 rows) the loop is launch-bound when run eagerly. `Generator.cons_generation` / `cons_inversion` and
`generation_sdxl.sample_deterministic` / `inverse_sample_deterministic` therefore capture the whole K-step loop
(K forwards + fused updates + the Python `AttentionStore` bookkeeping, which is plain tensor ops) into one CUDA
graph per (model, shapes, schedule, guidance, controller kind) on first use and replay it afterwards:

  first call   one eager run (fills TMA-descriptor / constant caches, allocates workspaces), capture, replay
  later calls  copy the inputs into the graph's static buffers, replay, hand out clones of the results

Only controllers whose effect is fully determined by their type are graphed: no controller, `EmptyControl`, and a
fresh plain `AttentionStore`. Edit controllers (stateful Python with per-edit tensors) and user subclasses run eagerly.
`ICD_CUDA_GRAPHS=0` disables the cache (every call runs eagerly); `ICD_MAX_GRAPHS` bounds it (LRU, default 8).
One in-flight call per process and device, like the rest of the library.
"""
import os
from collections import OrderedDict

import torch

MAX_GRAPHS = int(os.environ.get("ICD_MAX_GRAPHS", "8"))
_enabled = os.environ.get("ICD_CUDA_GRAPHS", "1") != "0"
stats = {"captures": 0, "replays": 0, "eager": 0}


def enabled():
    return _enabled and torch.cuda.is_available()


def set_enabled(flag):
    """Returns the previous setting."""
    global _enabled
    prev, _enabled = _enabled, bool(flag)
    return prev


def _cache_of(owner):
    """The cache lives ON the U-Net object: a captured graph holds raw pointers to that U-Net's packed weights, so
    it must not outlive it (and an `id()`-keyed global map could hand a stale graph to a new object)."""
    c = owner.__dict__.get("_icd_graphs")
    if c is None:
        c = owner.__dict__["_icd_graphs"] = OrderedDict()
    return c


def clear(owner):
    _cache_of(owner).clear()


def controller_signature(ctrl):
    """Hashable description of a controller whose whole effect on the loop is determined by its type, or None if
    the loop has to run eagerly with it."""
    from . import p2p
    if ctrl is None:
        return ("none",)
    if type(ctrl) is p2p.EmptyControl:
        return ("empty",)
    if type(ctrl) is p2p.AttentionStore:
        fresh = (not ctrl.attention_store and ctrl.cur_att_layer == 0
                 and all(len(v) == 0 for v in ctrl.step_store.values()))
        if fresh:
            return ("store", bool(ctrl.capture_self), int(ctrl.num_att_layers))
    return None


def proto_controller(ctrl):
    """A private controller of the same kind that lives with the captured graph (its stored maps are graph memory)."""
    from . import p2p
    if ctrl is None:
        return None
    if type(ctrl) is p2p.EmptyControl:
        return p2p.EmptyControl()
    proto = p2p.AttentionStore()
    proto.capture_self = ctrl.capture_self
    proto.num_att_layers = ctrl.num_att_layers
    return proto


def finish_controller(ctrl, proto, n_steps):
    """Leave the caller's controller in the state the eager loop would have left it in: step counter advanced by
    `n_steps`, `attention_store` holding the maps summed over the steps (fresh tensors: the graph's own buffers are
    overwritten by the next replay)."""
    from . import p2p
    if ctrl is None or type(ctrl) is p2p.EmptyControl:
        return
    ctrl.cur_step += n_steps
    ctrl.cur_att_layer = 0
    ctrl.step_store = ctrl.get_empty_store()
    ctrl.attention_store = {k: [m.clone() for m in v] for k, v in proto.attention_store.items()}


def run(owner, key, inputs, body):
    """`owner`: the B200UNet the loop runs (holds the cache). `body(*static_inputs) -> (outputs, aux)`: outputs is a list of CUDA tensors, aux any Python object created
    inside (kept alive with the graph). Returns (static outputs — valid until the next call with this key, aux)."""
    _cache = _cache_of(owner)
    activate = getattr(owner, "activate", None)     # unet.AdapterView: make its LoRA adapter the resident one
    if activate is not None:
        activate()
    entry = _cache.get(key)
    if entry is None:
        static_in = [x.clone() for x in inputs]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            body(*static_in)                    # eager warm-up outside the capture
        cur.wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            outs, aux = body(*static_in)
        entry = (graph, static_in, outs, aux)
        _cache[key] = entry
        stats["captures"] += 1
        while len(_cache) > max(1, MAX_GRAPHS):
            _cache.popitem(last=False)
    else:
        _cache.move_to_end(key)
        for dst, src in zip(entry[1], inputs):
            dst.copy_(src, non_blocking=True)
    entry[0].replay()
    stats["replays"] += 1
    return entry[2], entry[3]
