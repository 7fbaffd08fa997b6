"""Comparator only (never on the product path): cuBLAS (torch.matmul) vs gemm_tc on a few GEMM shapes, GPU time via
CUDA-graph replay."""
import sys, torch
sys.path.insert(0, '/root/repo')
from invertible_cd_b200 import ops

def t(fn, it=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(it): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3

for M, N, K in [(512, 1280, 11520), (2048, 1280, 11520), (8192, 640, 640), (2048, 1280, 1280), (32768, 320, 320),
                (32768, 320, 2880), (4096, 4096, 4096), (8192, 640, 5760)]:
    a = torch.randn(M, K, device='cuda').half(); w = (torch.randn(N, K, device='cuda') * K ** -0.5).half()
    out = torch.empty(M, N, device='cuda', dtype=torch.float16)
    t_ours = t(lambda: ops.linear(a, w, out=out))
    t_cublas = t(lambda: torch.matmul(a, w.t(), out=out))
    fl = 2.0 * M * N * K
    print(f"M={M:6d} N={N:5d} K={K:6d}: ours {t_ours:7.1f} us ({fl / t_ours / 1e6:7.1f} TF)   cuBLAS {t_cublas:7.1f} us ({fl / t_cublas / 1e6:7.1f} TF)")
