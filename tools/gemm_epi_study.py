#!/usr/bin/env python
"""Where does the time of the K = N = C transformer linears go? Graph-timed variants of the same problem: epilogue
features on/off (none / bias / bias+residual) x tile shapes. Usage: python tools/gemm_epi_study.py [--one]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from invertible_cd_b200 import ops  # noqa: E402
from tools._timing import time_us  # noqa: E402

SHAPES = [(32768, 320, 320), (8192, 640, 640), (2048, 1280, 1280), (4096, 1280, 1280), (16384, 640, 640),
          (32768, 320, 1280), (8192, 640, 2560), (2048, 1280, 5120), (4096, 1280, 5120), (16384, 640, 2560),
          (32768, 960, 320), (8192, 1920, 640), (16384, 1920, 640), (512, 1280, 1280), (65536, 320, 320)]
TILES = [(0, 0), (128, 160), (256, 160), (128, 128), (256, 128), (128, 256), (256, 64), (256, 256)]


def main():
    one = "--one" in sys.argv
    dev = "cuda"
    for M, N, K in SHAPES[:1] if one else SHAPES:
        a = torch.randn(M, K, device=dev).half()
        w = (torch.randn(N, K, device=dev) * K ** -0.5).half()
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev).half()
        out = torch.empty(M, N, device=dev, dtype=torch.float16)
        flops = 2.0 * M * N * K
        if one:
            for _ in range(3):
                ops.linear(a, w, bias=bias, residual=res, out=out)
            torch.cuda.synchronize()
            return
        for bm, bn in TILES:
            row = []
            for name, kw in (("plain", {}), ("bias", dict(bias=bias)), ("bias+res", dict(bias=bias, residual=res))):
                try:
                    us = time_us(lambda: ops.linear(a, w, out=out, force_bm=bm, force_bn=bn, **kw), 20)
                    row.append(f"{name} {us:6.1f} us {flops / us / 1e6:6.0f} TF")
                except Exception as e:
                    row.append(f"{name} ERR {str(e)[:40]}")
            print(f"M={M:6d} N={N:5d} K={K:5d} bm={bm:3d} bn={bn:3d} | " + " | ".join(row), flush=True)
        # memory floor: a copy of the same bytes (A + out [+ res])
        x = torch.empty(M * (K + 2 * N), device=dev, dtype=torch.float16)
        y = torch.empty_like(x)
        us = time_us(lambda: y.copy_(x), 20)
        print(f"    copy of (A + res + out) bytes x2 ({x.numel() * 4 / 1e6:.0f} MB moved): {us:6.1f} us", flush=True)


if __name__ == "__main__":
    main()
