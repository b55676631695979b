"""dev tool: ONE profiled launch of every kernel of the path (for `ncu --profile-from-start off`): warm-up launches
run outside the cudaProfilerStart/Stop range.

    python tools/prof_path_kernels.py stack   # headline kappa-sigma kernel at 100 x (9576 x 6388), the other stack
                                              # kernels at 100 x 4096^2 (keeps ncu's save/restore small)
    python tools/prof_path_kernels.py frame   # calibrate, fix_badpix, fused calibrate + repair, flat normalisation,
                                              # whole-image statistics + threshold mask at 9576 x 6388
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from astrophotography_b200 import kernels, synth

dev = torch.device("cuda", 0)
K = bench.HEADLINE
REF = dict(method="average", k_lo=5.0, k_hi=5.0, maxiters=1, cen="median", dev="mad_std")


def profiled(fn, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


what = sys.argv[1] if len(sys.argv) > 1 else "frame"
if what == "stack":
    n, h, w = 100, 6388, 9576
    cube = bench.synth_cube_device(torch, n, h, w, dev, seed=1000)
    out = {}
    profiled(lambda: kernels.stack_reduce(cube, out=out, **K))
    del cube, out
    torch.cuda.empty_cache()
    h = w = 4096
    cube = bench.synth_cube_device(torch, n, h, w, dev, seed=1000)
    out = {}
    profiled(lambda: kernels.stack_reduce(cube, out=out, **REF))
    profiled(lambda: kernels.stack_reduce(cube, out=out, method="median", maxiters=0, want_nrej=False))
    profiled(lambda: kernels.stack_reduce(cube, out=out, method="median", maxiters=0, want_nrej=False, want_uncert=True))
    c16 = torch.empty((n, h, w), dtype=torch.int16, device=dev)
    for i in range(n):
        c16[i] = cube[i].clamp(0, 65535).round().to(torch.int32).to(torch.int16)
    del cube
    torch.cuda.empty_cache()
    u16 = c16.view(torch.uint16)
    profiled(lambda: kernels.stack_reduce(u16, out=out, **K))
    profiled(lambda: kernels.stack_reduce(u16, out=out, **REF))
else:
    h, w = 6388, 9576
    g = torch.Generator(device=dev); g.manual_seed(3)
    raw = torch.randint(0, 65535, (h, w), dtype=torch.int32, device=dev, generator=g).to(torch.int16).view(torch.uint16)
    bias = torch.empty((h, w), device=dev).normal_(1000.0, 12.0, generator=g)
    dark = torch.empty((h, w), device=dev).normal_(1040.0, 12.0, generator=g)
    rawf = torch.empty((h, w), device=dev).normal_(3000.0, 50.0, generator=g)
    flat = (30000.0 * (1 + 0.01 * torch.randn((h, w), device=dev, generator=g))).contiguous()
    mask = torch.from_numpy(synth.badpix_mask((h, w), auto_fraction=1e-3)).to(dev)
    cal = torch.empty((h, w), dtype=torch.float32, device=dev)
    nflat, _norm = kernels.flat_normalise(flat)
    profiled(lambda: kernels.flat_normalise(flat))
    profiled(lambda: kernels.calibrate(raw, bias, dark, nflat, 1.0 / 3.0, True, out=cal))
    profiled(lambda: kernels.calibrate(rawf, bias, dark, nflat, 1.0 / 3.0, True, out=cal))
    profiled(lambda: kernels.fix_badpix(cal, mask, 2))
    profiled(lambda: kernels.calibrate_repair(raw, bias, dark, nflat, 1.0 / 3.0, True, mask=mask, deltapix=2, out=cal, out_big_endian=True))
    profiled(lambda: kernels.sigma_clipped_stats(dark, 4.0, maxiters=1), warm=1)
    profiled(lambda: kernels.threshold_mask(dark, 950.0, 1130.0))
torch.cuda.synchronize()
