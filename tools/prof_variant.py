import sys, torch
sys.path.insert(0, '/root/repo')
from astrophotography_b200 import kernels
prefer = None if sys.argv[1] == "default" else sys.argv[1]
n, h, w = (int(sys.argv[2]) if len(sys.argv) > 2 else 100), (int(sys.argv[3]) if len(sys.argv) > 3 else 2048), 9576
g = torch.Generator(device='cuda'); g.manual_seed(1)
cube = torch.empty((n, h, w), dtype=torch.float32, device='cuda')
for i in range(n):
    cube[i].normal_(1000.0, 12.0, generator=g)
# cosmic-ray-like hits (0.01 % of samples)
hits = torch.rand((n, h, w), device='cuda', generator=g) < 1e-4
cube[hits] += 5000.0
del hits
out = None
for _ in range(3):
    res = kernels.stack_reduce(cube, method="average", k_lo=3.0, k_hi=3.0, maxiters=5, cen="mean", dev="std", prefer=prefer)
torch.cuda.synchronize()
print(kernels.stack_last_staging())
