import sys, torch
sys.path.insert(0, '/root/repo')
from invertible_cd_b200 import ops
def t(fn, it=10):
    """GPU time per call: the calls are captured into one CUDA graph (no host launch overhead in the timing)."""
    for _ in range(2): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(it): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
M, N, K = 512, 1280, 11520
x = torch.randn(M, 1280, device='cuda').half(); w = (torch.randn(N, K, device='cuda') * K ** -0.5).half()
out = torch.empty(M, N, device='cuda', dtype=torch.float16); o32 = torch.empty(M, N, device='cuda')
for sp in (1, 2, 4, 7, 14):
    for bm, bn in ((128, 64), (128, 128), (128, 256), (256, 256), (256, 128)):
        try:
            us = t(lambda: ops.conv3x3(x, w, 8, 8, 8, out=out, force_bn=bn, force_bm=bm, force_splits=sp if sp > 1 else 0) if sp > 1 else ops.gemm_raw(a0=x, a_mode=1, K0=1280, a0_ld=1280, B=8, H=8, W=8, b=w, b_ld=K, ZB1=1, M=M, N=N, K=1280, Z=1, alpha=1.0, out=out, ldc=N, out_fp32=0, out_mode=0, force_bn=bn, force_bm=bm, ws=None, ws_bytes=0, rows_per_img=64))
            print(f"splits={sp:2d} bm={bm} bn={bn}: {us:8.1f} us")
        except Exception as e:
            print(sp, bm, bn, "ERR", str(e)[:80])
a = torch.randn(4096, 4096, device='cuda').half(); wl = torch.randn(4096, 4096, device='cuda').half()
print("4096^3 fp16 out", t(lambda: ops.linear(a, wl)), "fp32 out (direct)", t(lambda: ops.linear(a, wl, out_fp32=True)))
