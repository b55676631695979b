import sys, torch
sys.path.insert(0, '/root/repo')
from astrophotography_b200 import kernels
n, h, w = (int(sys.argv[1]) if len(sys.argv) > 1 else 100), 2048, 9576
g = torch.Generator(device='cuda'); g.manual_seed(1)
cube = torch.empty((n, h, w), dtype=torch.float32, device='cuda')
for i in range(n):
    cube[i].normal_(1000.0, 12.0, generator=g)
    hits = torch.rand((h, w), device='cuda', generator=g) < 1e-4
    cube[i][hits] += 5000.0
for _ in range(3):
    kernels.stack_reduce(cube, method="average", k_lo=5.0, k_hi=5.0, maxiters=1, cen="median", dev="mad_std")
torch.cuda.synchronize()
