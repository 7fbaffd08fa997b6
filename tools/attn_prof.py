#!/usr/bin/env python
"""Per-phase cycle breakdown of one softmax warp of the attention kernel (needs a `make PROF=1` build of the library)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from invertible_cd_b200 import _lib, ops  # noqa: E402

for (B, H, Nq, Nk, D) in [(8, 8, 4096, 4096, 40), (4, 10, 4096, 4096, 64), (4, 20, 1024, 1024, 64), (8, 8, 1024, 1024, 80)]:
    q = torch.randn(B * Nq, H * D, device="cuda").half()
    k = torch.randn(B * Nk, H * D, device="cuda").half()
    v = torch.randn(B * Nk, H * D, device="cuda").half()
    out = torch.empty_like(q)
    for _ in range(3):
        ops.attention(q, k, v, B, H, Nq, Nk, D, D ** -0.5, out=out)
    torch.cuda.synchronize()
    print(f"B={B} H={H} Nq={Nq} Nk={Nk} D={D} MT={os.environ.get('ICD_ATTN_MT')} POLY={os.environ.get('ICD_ATTN_POLY')}")
    sys.stdout.flush()
    _lib.load().icd_attention_prof_dump(ctypes.c_int((Nk + 63) // 64))
