"""Time the sorted (median / median-MAD) kernels on synthetic cubes (dev tool)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astrophotography_b200 import kernels
g = torch.Generator(device="cuda"); g.manual_seed(1)
cfgs = [(30, 4096, 4096), (64, 6388, 9576), (100, 6388, 9576), (200, 6388, 9576)]
if len(sys.argv) > 1:
    cfgs = [tuple(int(v) for v in sys.argv[1:4])]
for (n, h, w) in cfgs:
    cube = torch.empty((n, h, w), dtype=torch.float32, device="cuda")
    for i in range(n):
        cube[i].normal_(1000.0, 12.0, generator=g)
    for name, kw in (("median", dict(method="median", maxiters=0, want_nrej=False)),
                     ("medmad", dict(method="average", k_lo=5.0, k_hi=5.0, maxiters=1, cen="median", dev="mad_std"))):
        out = {}
        for _ in range(2):
            kernels.stack_reduce(cube, out=out, **kw)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            kernels.stack_reduce(cube, out=out, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(n, name, round(ms, 3), "ms", round((4 * n + 5) * h * w / ms / 1e6), "GB/s", round((4 * n + 5) * h * w / ms / 1e6 / 6459, 3))
    del cube
