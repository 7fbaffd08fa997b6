"""CUDA-graph cache for the K-step consistency loops of the product path.

The reference makes ONE library call per step (`model.unet(...)`, utils/generation.py:241-244); here a U-Net forward is
~350 (SD1.5) / ~920 (SDXL) kernel launches issued from Python, so at small batch (BASELINE configs[0]: 1 row,
configs[2]: 2 rows) the loop is launch-bound when run eagerly. `Generator.cons_generation` / `cons_inversion` and
`generation_sdxl.sample_deterministic` / `inverse_sample_deterministic` therefore capture the whole K-step loop
(K forwards + fused updates + the Python `AttentionStore` bookkeeping, which is plain tensor ops) into one CUDA
graph per (model, shapes, schedule, guidance, controller kind) on first use and replay it afterwards:

  first call   one eager run (fills TMA-descriptor / constant caches, allocates workspaces), capture, replay
  later calls  copy the inputs into the graph's static buffers, replay, hand out clones of the results

Graphed controllers: none, `EmptyControl`, a fresh plain `AttentionStore`, and fresh prompt-to-prompt edit controllers
(`AttentionReplace` / `AttentionRefine` / `AttentionReweight`, with or without `LocalBlend`): their Python control flow
is a function of a few host-side values (class, prompt count, replace windows, blend thresholds — all part of the cache
key) and their per-edit tensors (token mapper, alphas, equalizer, word masks) are INPUTS of the graph, copied into its
static buffers before a replay — so the next edit with other prompts of the same shape replays the same graph
(BASELINE configs[2]: the edit loop is launch-bound when run eagerly). User subclasses run eagerly.
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


_EDIT_TENSORS = ("cross_replace_alpha", "mapper", "alphas", "equalizer")
_BLEND_TENSORS = ("alpha_layers", "substruct_layers")


_REQUIRE_CUDA = True      # CPU tests of the plumbing below switch this off


def _tensor_slots(obj, names, tag, sig, slots):
    for a in names:
        t = getattr(obj, a, None)
        if torch.is_tensor(t):
            if _REQUIRE_CUDA and not t.is_cuda:
                return False
            sig.append((tag, a, tuple(t.shape), str(t.dtype)))
            slots.append((tag, a))
    return True


def _edit_parts(ctrl):
    """(signature, [(owner tag, attribute)]) of a fresh prompt-to-prompt edit controller, or None."""
    from . import p2p
    if type(ctrl) not in (p2p.AttentionReplace, p2p.AttentionRefine, p2p.AttentionReweight):
        return None
    fresh = (ctrl.cur_step == 0 and ctrl.cur_att_layer == 0 and not ctrl.attention_store
             and all(len(v) == 0 for v in ctrl.step_store.values()))
    if not fresh:
        return None
    sig = ["edit", type(ctrl).__name__, int(ctrl.batch_size), int(ctrl.num_att_layers), bool(ctrl.capture_self),
           tuple(ctrl.num_self_replace), tuple(ctrl._cross_active), bool(p2p.LOW_RESOURCE)]
    slots = []
    if not _tensor_slots(ctrl, _EDIT_TENSORS, "ctrl", sig, slots):
        return None
    lb = ctrl.local_blend
    if lb is not None:
        if type(lb) is not p2p.LocalBlend or lb.counter != 0:
            return None
        sig.append(("blend", int(lb.start_blend), tuple(float(x) for x in lb.th)))
        if not _tensor_slots(lb, _BLEND_TENSORS, "blend", sig, slots):
            return None
    prev = getattr(ctrl, "prev_controller", None)
    if prev is not None:
        if type(prev) not in (p2p.AttentionReplace, p2p.AttentionRefine):
            return None
        sig.append(("prev", type(prev).__name__))
        if not _tensor_slots(prev, ("mapper", "alphas"), "prev", sig, slots):
            return None
    return tuple(sig), slots


def _slot_owner(ctrl, tag):
    return ctrl if tag == "ctrl" else (ctrl.local_blend if tag == "blend" else ctrl.prev_controller)


def controller_tensors(ctrl):
    """The per-edit device tensors of a graphable edit controller, in signature order (graph inputs); [] otherwise."""
    parts = _edit_parts(ctrl) if ctrl is not None else None
    if parts is None:
        return []
    return [getattr(_slot_owner(ctrl, tag), a) for tag, a in parts[1]]


def controller_signature(ctrl):
    """Hashable description of a controller whose whole effect on the loop is determined by it (and by the tensors
    `controller_tensors` lists), or None if the loop has to run eagerly with it."""
    from . import p2p
    parts = _edit_parts(ctrl) if ctrl is not None else None
    if parts is not None:
        return parts[0]
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


def proto_controller(ctrl, tensors=()):
    """A private controller of the same kind that lives with the captured graph (its stored maps are graph memory).
    `tensors`: the graph's static copies of `controller_tensors(ctrl)` for an edit controller."""
    import copy
    from . import p2p
    if ctrl is None:
        return None
    parts = _edit_parts(ctrl)
    if parts is not None:
        proto = copy.copy(ctrl)
        proto.step_store, proto.attention_store = proto.get_empty_store(), {}
        if proto.local_blend is not None:
            proto.local_blend = copy.copy(proto.local_blend)
        if getattr(proto, "prev_controller", None) is not None:
            proto.prev_controller = copy.copy(proto.prev_controller)
        assert len(tensors) == len(parts[1])
        for (tag, a), t in zip(parts[1], tensors):
            setattr(_slot_owner(proto, tag), a, t)
        return proto
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
    if getattr(ctrl, "local_blend", None) is not None:
        ctrl.local_blend.counter += n_steps


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
