#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 GEMM / implicit-conv kernel on the shapes of one SD1.5 (B=8) U-Net forward.
Usage: python tools/gemm_bench.py [--only IDX] [--iters N]   (CUDA events, L2-warm; prints TFLOP/s per shape)"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from invertible_cd_b200 import ops  # noqa: E402
from tools._timing import time_us  # noqa: E402
from invertible_cd_b200.packing import pack_geglu  # noqa: E402

SHAPES = [  # (kind, M, N, K, extras)
    ("lin", 32768, 320, 320, "bias+res"),
    ("lin", 8192, 640, 640, "bias+res"),
    ("lin", 2048, 1280, 1280, "bias+res"),
    ("lin", 32768, 960, 320, ""),
    ("lin", 32768, 2560, 320, "geglu"),
    ("lin", 8192, 5120, 640, "geglu"),
    ("lin", 32768, 320, 1280, "bias+res"),
    ("conv", 32768, 320, 320, "64"),      # B=8, 64x64
    ("conv", 8192, 640, 640, "32"),
    ("conv", 2048, 1280, 1280, "16"),
    ("conv", 512, 1280, 1280, "8"),
    ("conv", 512, 1280, 2560, "8"),
    ("lin", 4096, 4096, 4096, ""),
    # SDXL (B=4, 1024^2): the 32x32 level of the transformer stacks
    ("lin", 4096, 1280, 1280, "bias+res"),
    ("lin", 4096, 1280, 5120, "bias+res"),
    ("lin", 4096, 3840, 1280, ""),
    ("lin", 4096, 10240, 1280, "geglu"),
    ("lin", 16384, 640, 640, "bias+res"),
    ("lin", 16384, 1920, 640, ""),
    ("lin", 8192, 1920, 640, ""),
    ("lin", 2048, 3840, 1280, ""),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", type=int, default=-1)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--bn", type=int, default=0)
    ap.add_argument("--bm", type=int, default=0)
    args = ap.parse_args()
    dev = "cuda"
    for idx, (kind, M, N, K, extra) in enumerate(SHAPES):
        if args.only >= 0 and idx != args.only:
            continue
        if kind == "lin":
            a = torch.randn(M, K, device=dev).half()
            w = (torch.randn(N, K, device=dev) * K ** -0.5).half()
            bias = torch.randn(N, device=dev) if "bias" in extra or "geglu" in extra else None
            res = torch.randn(M, N, device=dev).half() if "res" in extra else None
            geglu = "geglu" in extra
            if geglu:
                w, bias = pack_geglu(w, bias, 256)
            out = torch.empty(M, N // 2 if geglu else N, device=dev, dtype=torch.float16)
            fn = lambda: ops.linear(a, w, bias=bias, residual=res, out=out, geglu=geglu,
                                    force_bn=256 if geglu else args.bn, force_bm=0 if geglu else args.bm)
            flops = 2.0 * M * N * K
        else:
            HW = int(extra)
            B = M // (HW * HW)
            x = torch.randn(M, K, device=dev).half()
            w = (torch.randn(N, 9 * K, device=dev) * (9 * K) ** -0.5).half()
            bias = torch.randn(N, device=dev)
            out = torch.empty(M, N, device=dev, dtype=torch.float16)
            fn = lambda: ops.conv3x3(x, w, B, HW, HW, bias=bias, out=out, force_bn=args.bn, force_bm=args.bm)
            flops = 2.0 * M * N * K * 9
        us = time_us(fn, args.iters)
        print(f"[{idx:2d}] {kind:4s} M={M:6d} N={N:5d} K={K * (9 if kind == 'conv' else 1):6d} {extra:9s} "
              f"bn={ops.pick_bn(M, N, 1, 'geglu' in extra, False, 256 if 'geglu' in extra else args.bn):3d} "
              f"{us:9.1f} us  {flops / us / 1e6:8.1f} TFLOP/s")


if __name__ == "__main__":
    main()
