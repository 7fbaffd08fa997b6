#!/usr/bin/env python
"""Micro-benchmark of GroupNorm(+SiLU) / LayerNorm on the SD1.5 (B=8) shapes: achieved GB/s vs algorithmic bytes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from invertible_cd_b200 import ops  # noqa: E402
from tools._timing import time_us  # noqa: E402

if __name__ == "__main__":
    ws = torch.empty(8 * 128 * 64, device="cuda")
    sd15 = [(4096, 320, 0), (4096, 640, 320), (4096, 320, 320), (1024, 640, 0), (1024, 1280, 640),
            (256, 1280, 0), (256, 1280, 1280), (64, 1280, 1280)]
    sdxl = [(16384, 320, 0), (4096, 640, 0), (4096, 640, 320), (1024, 1280, 0), (1024, 1280, 1280), (1024, 1280, 640)]
    cases = [(8, s) for s in sd15]
    if "--all" in sys.argv:
        cases += [(4, s) for s in sdxl] + [(2, s) for s in sd15[:1] + sd15[3:6]] + [(1, s) for s in sd15[:1] + sd15[3:6]]
    for B, (HW, C0, C1) in cases:
        x0 = torch.randn(B * HW, C0, device="cuda").half()
        x1 = torch.randn(B * HW, C1, device="cuda").half() if C1 else None
        C = C0 + C1
        g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        out = torch.empty(B * HW, C, device="cuda", dtype=torch.float16)
        fn = lambda: ops.groupnorm(x0, B, HW, g, b, 1e-5, True, ws, x1=x1, out=out)
        us = time_us(fn, 20)
        byt = 2.0 * B * HW * C * 2
        print(f"GN  B={B} HW={HW:5d} C={C0}+{C1}: {us:7.1f} us  {byt / us / 1e3:7.1f} GB/s (1R+1W)  {byt / 1e6:6.1f} MB")
    B = 8
    for rows, C in [(32768, 320), (8192, 640), (2048, 1280)]:
        x = torch.randn(rows, C, device="cuda").half()
        g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        out = torch.empty_like(x)
        fn = lambda: ops.layernorm(x, g, b, 1e-5, out=out)
        us = time_us(fn, 20)
        byt = 2.0 * rows * C * 2
        print(f"LN  rows={rows:6d} C={C}: {us:7.1f} us  {byt / us / 1e3:7.1f} GB/s")
