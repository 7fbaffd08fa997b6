#!/usr/bin/env python
"""Micro-benchmark of the fused attention kernel on the SD1.5 / SDXL shapes (CUDA events, L2-warm)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from invertible_cd_b200 import ops  # noqa: E402
from tools._timing import time_us  # noqa: E402

SHAPES = [(8, 8, 4096, 4096, 40), (8, 8, 1024, 1024, 80), (8, 8, 256, 256, 160), (8, 8, 4096, 77, 40),
          (8, 8, 1024, 77, 80), (8, 8, 256, 77, 160),
          (4, 10, 4096, 4096, 64), (4, 20, 1024, 1024, 64), (4, 20, 1024, 77, 64)]

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", type=int, default=-1)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--no-capture", action="store_true", help="cross-attention shapes without the probability write-out")
    ap.add_argument("--cross-only", action="store_true")
    a = ap.parse_args()
    for i, (B, H, Nq, Nk, D) in enumerate(SHAPES):
        if (a.only >= 0 and i != a.only) or (a.cross_only and Nk > 128):
            continue
        q = torch.randn(B * Nq, H * D, device="cuda").half()
        k = torch.randn(B * Nk, H * D, device="cuda").half()
        v = torch.randn(B * Nk, H * D, device="cuda").half()
        out = torch.empty_like(q)
        probs = torch.empty(B * H, Nq, (Nk + 7) // 8 * 8, device="cuda", dtype=torch.float16) if Nk <= 128 and not a.no_capture else None
        fn = lambda: ops.attention(q, k, v, B, H, Nq, Nk, D, D ** -0.5, out=out, probs_out=probs)
        us = time_us(fn, a.iters)
        fl = 4.0 * B * H * Nq * Nk * D
        tiles = B * H * ((Nq + 127) // 128) * ((Nk + 127) // 128)
        print(f"[{i}] B={B} H={H} Nq={Nq} Nk={Nk} D={D}{' +capture' if probs is not None else ''}: {us:9.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  "
              f"{us * 1e-6 * 1.9e9 * 148 / tiles:7.0f} SM-cycles/tile")
