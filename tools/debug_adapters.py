#!/usr/bin/env python
"""Debug aid: swap-mode adapter views vs resident fused models, forward by forward."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from invertible_cd_b200 import arch, graphs, loading  # noqa: E402

cfg = arch.small_sd15_config()
lr, lf = arch.synthetic_lora(cfg, r=8, seed=1, std=0.05), arch.synthetic_lora(cfg, r=8, seed=2, std=0.05)
ldm, rev, fwd = loading.load_models("synthetic:small_sd15:0", "cuda", lr, lf, r=8, w_embed_dim=512, dtype="fp16")
s_ldm, s_rev, s_fwd = loading.load_models("synthetic:small_sd15:0", "cuda", lr, lf, r=8, w_embed_dim=512, dtype="fp16",
                                          adapters="swap")
g = torch.Generator().manual_seed(0)
x = torch.randn(2, 4, 64, 64, generator=g).cuda()
ctx = torch.randn(2, 77, cfg.cross_attention_dim, generator=g).half().cuda()
wemb = rev.unet.guidance_embedding(rev.unet.cached_vector([3.0, 3.0]), 512)


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


for name, res, view in (("teacher", ldm, s_ldm), ("reverse", rev, s_rev), ("forward", fwd, s_fwd), ("reverse", rev, s_rev),
                        ("teacher", ldm, s_ldm)):
    a = res.unet(x, 500, encoder_hidden_states=ctx, timestep_cond=wemb)["sample"]
    b = view.unet(x, 500, encoder_hidden_states=ctx, timestep_cond=wemb)["sample"]
    torch.cuda.synchronize()
    print(name, "single forward: swap vs resident rel-L2", rel(b, a), "active:", s_ldm.unet._shared.active_adapter)
a = rev.unet(x, 500, encoder_hidden_states=ctx, timestep_cond=wemb)["sample"]
b = fwd.unet(x, 500, encoder_hidden_states=ctx, timestep_cond=wemb)["sample"]
c = ldm.unet(x, 500, encoder_hidden_states=ctx, timestep_cond=wemb)["sample"]
print("resident reverse vs forward rel-L2", rel(a, b), " reverse vs teacher", rel(a, c))
# same weights, two runs of the resident model: the determinism floor
print("resident reverse twice", rel(rev.unet(x, 500, encoder_hidden_states=ctx, timestep_cond=wemb)["sample"], a))

# ---- E3: every tensor reachable from the executor, swap (reverse active) vs resident reverse
from types import SimpleNamespace


def walk(obj, path, out, seen):
    if id(obj) in seen:
        return
    seen.add(id(obj))
    if torch.is_tensor(obj):
        out[path] = obj
    elif isinstance(obj, SimpleNamespace):
        for k, v in vars(obj).items():
            walk(v, f"{path}.{k}", out, seen)
    elif isinstance(obj, (list, tuple)):
        for i, v in enumerate(obj):
            walk(v, f"{path}[{i}]", out, seen)
    elif isinstance(obj, dict):
        for k, v in obj.items():
            if isinstance(k, str) and not k.startswith("_"):
                walk(v, f"{path}.{k}", out, seen)


shared = s_ldm.unet._shared
s_rev.unet.activate()
torch.cuda.synchronize()
ta, tb = {}, {}
walk({k: v for k, v in vars(shared).items() if not k.startswith("_")}, "unet", ta, set())
walk({k: v for k, v in vars(rev.unet).items() if not k.startswith("_")}, "unet", tb, set())
print("tensors:", len(ta), len(tb))
bad = 0
diffs = []
for k in tb:
    if k not in ta:
        print("missing in shared:", k); continue
    a, b = ta[k], tb[k]
    if a.shape != b.shape or not torch.equal(a, b):
        d = (a.float() - b.float()).abs().max().item() if a.shape == b.shape else -1
        nz = int((a != b).sum()) if a.shape == b.shape else -1
        diffs.append((nz / max(1, a.numel()), d, k, tuple(a.shape), nz))
        bad += 1
print("differing tensors:", bad)
for fr, d, k, shp, nz in sorted(diffs, reverse=True)[:12]:
    print("DIFF by fraction", k, shp, "max abs", d, "n", nz, "frac", fr)
for fr, d, k, shp, nz in sorted(diffs, key=lambda t: -t[1])[:8]:
    print("DIFF by max abs", k, shp, "max abs", d, "n", nz, "frac", fr)
b1 = shared.forward(x, 500, encoder_hidden_states=ctx, timestep_cond=wemb)["sample"]
a1 = rev.unet(x, 500, encoder_hidden_states=ctx, timestep_cond=wemb)["sample"]
print("after sync, shared.forward vs resident:", rel(b1, a1))
