#!/usr/bin/env python
"""Timeline of one GEMM launch from in-kernel clock64 stamps (debug build: make -C invertible_cd_b200/csrc GPROF=1;
run with ICD_LIB_PATH=invertible_cd_b200/libicd_b200_gprof.so). Prints, for a few CTAs, cycles relative to kernel entry:
setup done, pdl_wait passed, first operands landed, per tile: MMAs issued / accumulator ready / epilogue drained."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from invertible_cd_b200 import _lib, ops  # noqa: E402

SHAPES = [(32768, 320, 320), (4096, 1280, 1280)]


def main():
    lib = _lib.load()
    lib.icd_gemm_prof_buffer.restype = ctypes.POINTER(ctypes.c_longlong)
    for M, N, K in SHAPES:
        for extra in ("plain", "bias+res"):
            a = torch.randn(M, K, device="cuda").half()
            w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
            bias = torch.randn(N, device="cuda") if extra != "plain" else None
            res = torch.randn(M, N, device="cuda").half() if extra != "plain" else None
            out = torch.empty(M, N, device="cuda", dtype=torch.float16)
            for _ in range(3):
                ops.linear(a, w, bias=bias, residual=res, out=out)
            torch.cuda.synchronize()
            buf = lib.icd_gemm_prof_buffer()
            print(f"== M={M} N={N} K={K} {extra}")
            for cta in (73,):
                s = [buf[cta * 32 + i] for i in range(32)]
                t0 = s[0]
                rel = lambda i: (s[i] - t0) if s[i] else None
                tiles = []
                for it in range(7):
                    if s[8 + 3 * it] and s[8 + 3 * it] > t0:
                        tiles.append((rel(8 + 3 * it), rel(9 + 3 * it), rel(10 + 3 * it)))
                print(f"  cta {cta:3d}: setup {rel(1)} pdl {rel(2)} first-operands {rel(4)} producer-done {rel(3)} "
                      f"mma-done {rel(5)} epi-done {rel(6)} end {rel(7)} | tiles (mma issued, acc ready, drained): {tiles}")
            for i in range(256 * 32):
                buf[i] = 0


if __name__ == "__main__":
    main()
