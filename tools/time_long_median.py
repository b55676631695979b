"""dev tool: plain median of long stacks (lane-cooperative kernel) vs the generic kernel, 4096 x 4096 frames."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from astrophotography_b200 import kernels
PEAK = bench.peaks()[0]
dev = torch.device("cuda", 0)
h = w = 4096
cube = bench.synth_cube_device(torch, 512, h, w, dev, seed=1000)
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out = {}
for n in (224, 256, 320, 384, 448, 512):
    ms = timeit(lambda: kernels.stack_reduce(cube[:n], out=out, method="median", maxiters=0, want_nrej=False))
    nb = (4 * n + 4) * h * w
    print(f"N={n} {kernels.stack_kernel_name(n, 'median', maxiters=0)}  {ms:8.3f} ms  {nb/ms/1e6:6.0f} GB/s  {nb/ms/1e6/PEAK:.3f} of peak", flush=True)
REF = dict(method="average", k_lo=5.0, k_hi=5.0, maxiters=1, cen="median", dev="mad_std")
out2 = {}
for n in (256, 512):
    ms = timeit(lambda: kernels.stack_reduce(cube[:n], out=out2, **REF))
    nb = (4 * n + 5) * h * w
    print(f"N={n} {kernels.stack_kernel_name(n)}  {ms:8.3f} ms  {nb/ms/1e6:6.0f} GB/s  {nb/ms/1e6/PEAK:.3f} of peak", flush=True)
    ms = timeit(lambda: kernels.stack_reduce(cube[:n], out=out2, method="median", maxiters=0, want_nrej=False, want_uncert=True))
    print(f"N={n} {kernels.stack_kernel_name(n, 'median', maxiters=0, want_uncert=True)}  {ms:8.3f} ms  {(4*n+8)*h*w/ms/1e6/PEAK:.3f} of peak", flush=True)
ms = timeit(lambda: kernels.stack_reduce(cube[:256], out=out, method="median", maxiters=0, want_nrej=False, force_generic=True), reps=1)
print(f"N=256 generic {ms:8.3f} ms  {(4*256+4)*h*w/ms/1e6/PEAK:.3f} of peak")
