#!/usr/bin/env python
"""How close is the on-GPU LoRA fuse (W_base + s * B.A by the tcgen05 GEMM, unet.B200UNet.set_adapter) to the host fuse
in fp32 (loading.fuse_lora)? Counts fp16-ulp differences per adapted module for two epilogue routes:
  mma   : s folded into B, base weights added on the tensor core (identity MMA) -> one fp16 rounding
  alpha : s passed as the GEMM's alpha, base weights added in the fp32 register epilogue"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from invertible_cd_b200 import arch, loading, ops  # noqa: E402
from invertible_cd_b200.unet import B200UNet  # noqa: E402

if __name__ == "__main__":
    cfg = arch.small_sd15_config()
    sd = arch.synthetic_state_dict(cfg, seed=0)
    lora = arch.synthetic_lora(cfg, r=8, seed=1, std=0.05)
    net = B200UNet(cfg, sd, "cuda")
    fused = B200UNet(cfg, loading.fuse_lora(sd, lora, r=8), "cuda")
    net.add_adapter("rev", lora)
    tot = {"mma": [0, 0, 0.0], "alpha": [0, 0, 0.0]}
    for mod, (Bp, At) in net._adapters["rev"].items():
        _, getter, row0, nrows, _ = net._wmap[mod]
        base = net._base_copy[getter().data_ptr()][row0:row0 + nrows]
        want = fused._wmap[mod][1]()[row0:row0 + nrows]
        for route in ("mma", "alpha"):
            out = torch.empty_like(base)
            if route == "mma":
                ops.linear(Bp, At, residual=base, out=out)
            else:
                ops.linear((Bp.float() * 8).half(), At, residual=base, out=out, alpha=0.125)
            d = (out.view(torch.int16).int() - want.view(torch.int16).int())
            tot[route][0] += int((d != 0).sum())
            tot[route][1] += d.numel()
            tot[route][2] += float(d.sum())              # bit patterns: > 0 = magnitude larger than the host fuse
    for route, (nz, n, bias) in tot.items():
        print(f"{route}: {nz}/{n} elements differ from the host fuse ({100.0 * nz / n:.3f} %), "
              f"mean signed difference {bias / n:+.5f} ulp")
