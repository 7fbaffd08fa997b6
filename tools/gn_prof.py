#!/usr/bin/env python
"""Phase timeline of the single-pass GroupNorm (needs `make VARIANT=gnprof EXTRA_DEFS=-DICD_GN_PROFILE` and
ICD_LIB_PATH=.../libicd_b200_gnprof.so): cycles between the clock64 stamps of thread 0, averaged over the first 256 CTAs."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from invertible_cd_b200 import _lib, ops  # noqa: E402

NAMES = ["pdl_wait", "load+stash", "reduce+publish", "grid barrier", "finalize", "normalise+store"]
lib = _lib.load()
lib.icd_gn_prof_buffer.restype = ctypes.POINTER(ctypes.c_longlong)
buf = lib.icd_gn_prof_buffer()
B = 8
ws = torch.empty(B * 128 * 64, device="cuda")
for HW, C0, C1 in [(4096, 320, 0), (1024, 640, 0), (4096, 320, 320), (256, 1280, 0)]:
    x0 = torch.randn(B * HW, C0, device="cuda").half()
    x1 = torch.randn(B * HW, C1, device="cuda").half() if C1 else None
    C = C0 + C1
    g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    out = torch.empty(B * HW, C, device="cuda", dtype=torch.float16)
    for _ in range(5):                      # warm (L2-resident input, as in the step)
        ops.groupnorm(x0, B, HW, g, b, 1e-5, True, ws, x1=x1, out=out)
    torch.cuda.synchronize()
    n = 256
    st = [[buf[i * 8 + k] for k in range(7)] for i in range(n)]
    st = [s for s in st if s[6] > s[0] > 0]
    avg = [sum(s[k + 1] - s[k] for s in st) / max(1, len(st)) for k in range(6)]
    tot = sum(s[6] - s[0] for s in st) / max(1, len(st))
    print(f"GN B={B} HW={HW} C={C0}+{C1}: CTAs {len(st)}  total {tot:.0f} cycles  " +
          "  ".join(f"{nm} {v:.0f}" for nm, v in zip(NAMES, avg)))
