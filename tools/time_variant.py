"""Time one staging variant of the headline kappa-sigma stack on a synthetic cube (dev tool).

    python tools/time_variant.py registers_tensormap [n h w]
"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astrophotography_b200 import kernels
prefer = None if sys.argv[1] == "default" else sys.argv[1]
n, h, w = (int(x) for x in sys.argv[2:5]) if len(sys.argv) >= 5 else (100, 6388, 9576)
g = torch.Generator(device='cuda'); g.manual_seed(1)
cube = torch.empty((n, h, w), dtype=torch.float32, device='cuda')
for i in range(n):
    cube[i].normal_(1000.0, 12.0, generator=g)
    hits = torch.rand((h, w), device='cuda', generator=g) < 1e-4     # cosmic-ray-like hits
    cube[i][hits] += 5000.0
kw = dict(method="average", k_lo=3.0, k_hi=3.0, maxiters=5, cen="mean", dev="std", prefer=prefer)
res = kernels.stack_reduce(cube, **kw)
out = res
for _ in range(3):
    kernels.stack_reduce(cube, **kw)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
K = 10
ev[0].record()
for _ in range(K):
    kernels.stack_reduce(cube, **kw)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / K
gb = (4 * n + 5) * h * w / 1e9
print(f"{prefer} staging={kernels.stack_last_staging()} env={os.environ.get('APGPU_TMAP_TILES_PER_WARP')} n={n} {h}x{w}: {ms:.3f} ms  {gb/ms*1e3:.0f} GB/s  {gb/ms*1e3/6459:.3f} of 6459")
