import sys, torch
sys.path.insert(0, '/root/repo')
from invertible_cd_b200 import ops
M, N, K = 512, 1280, 11520
a = torch.randn(M, K, device='cuda').half(); w = (torch.randn(N, K, device='cuda') * K ** -0.5).half()
out = torch.empty(M, N, device='cuda', dtype=torch.float16)
for _ in range(3):
    ops.linear(a, w, out=out, force_bn=128, force_bm=128)
torch.cuda.synchronize()
