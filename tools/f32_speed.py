#!/usr/bin/env python
"""Speed of the fp32 validation path: a few SGEMM shapes (TFLOP/s on the FMA pipe) and one full SD1.5 row-forward in fp32
vs fp16 (eager launches, CUDA events)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from invertible_cd_b200 import arch, ops_f32  # noqa: E402
from invertible_cd_b200.unet import B200UNet  # noqa: E402
from tools._timing import time_us  # noqa: E402

if __name__ == "__main__":
    for M, K, N in [(4096, 320, 320), (4096, 1280, 1280), (1024, 1280, 5120), (4096, 2880, 320)]:
        a, w = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
        out = torch.empty(M, N, device="cuda")
        us = time_us(lambda: ops_f32.linear(a, w, out=out), 10)
        print(f"sgemm M={M} N={N} K={K}: {us:8.1f} us  {2.0 * M * N * K / us / 1e6:6.2f} TFLOP/s")
    B, H, W, Cin, Cout = 1, 64, 64, 320, 320
    x, w = torch.randn(B * H * W, Cin, device="cuda"), torch.randn(Cout, 9 * Cin, device="cuda")
    us = time_us(lambda: ops_f32.conv3x3(x, w, B, H, W), 10)
    print(f"sgemm conv3x3 {H}x{W}x{Cin}->{Cout}: {us:8.1f} us  {2.0 * B * H * W * Cout * 9 * Cin / us / 1e6:6.2f} TFLOP/s")
    cfg = arch.sd15_config(time_cond_proj_dim=512)
    sd = arch.synthetic_state_dict(cfg, seed=0, device="cuda")
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(1, 4, 64, 64, generator=g).cuda()
    ctx = torch.randn(1, 77, 768, generator=g).cuda()
    for prec in ("fp16", "fp32"):
        net = B200UNet(cfg, sd, "cuda", precision=prec)
        wemb = net.guidance_embedding(net.cached_vector([19.0]), 512)
        fn = lambda: net(lat, 999, encoder_hidden_states=ctx.to(net.dtype), timestep_cond=wemb)
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"SD1.5 row-forward, {prec}: {e0.elapsed_time(e1) / 3:8.2f} ms (eager launches)")
        del net
