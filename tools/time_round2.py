"""Time the round-2 kernels on synthetic cubes (dev tool): median/MAD clip, median + uncertainty,
uint16 frames.  Usage: python tools/time_round2.py [quick]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                     # noqa: E402  (synthetic cube: hot pixels + cosmic hits)
from astrophotography_b200 import kernels                        # noqa: E402

PEAK = bench.peaks()[0]
MEDMAD = dict(method="average", k_lo=5.0, k_hi=5.0, maxiters=1, cen="median", dev="mad_std")
KAPPA = dict(method="average", k_lo=3.0, k_hi=3.0, maxiters=5, cen="mean", dev="std")


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(tag, ms, nbytes, extra=""):
    gbs = nbytes / ms / 1e6
    print(f"{tag:58s} {ms:8.3f} ms {gbs:7.0f} GB/s  {gbs / PEAK:5.3f} of peak {extra}", flush=True)


def main():
    quick = "quick" in sys.argv
    dev = torch.device("cuda", 0)
    cfgs = [(30, 4096, 4096), (64, 4096, 4096), (100, 6388, 9576)] if not quick else [(30, 2048, 2048)]
    for (n, h, w) in cfgs:
        cube = bench.synth_cube_device(torch, n, h, w, dev, seed=1000)
        if "clean" in sys.argv:                      # no hot pixels / cosmic hits: nothing takes the slow path
            for i in range(n):
                cube[i].normal_(1000.0, 12.0)
        px = h * w
        out32, out64 = {}, {}
        for name, kw, out in (("f32 out", dict(), out32), ("f64 out + uncert (ApMasterCal default)", dict(out_f64=True, want_uncert=True), out64)):
            ms = timeit(lambda: kernels.stack_reduce(cube, out=out, **MEDMAD, **kw))
            report(f"N={n} {h}x{w} medmad, {name}", ms, (4 * n + 5) * px)
        ms = timeit(lambda: kernels.stack_reduce(cube, out=out32, method="median", maxiters=0, want_nrej=False))
        report(f"N={n} {h}x{w} median", ms, (4 * n + 4) * px)
        ms = timeit(lambda: kernels.stack_reduce(cube, out=out64, method="median", maxiters=0, want_nrej=False, want_uncert=True))
        report(f"N={n} {h}x{w} median + MAD uncertainty (f64)", ms, (4 * n + 16) * px)
        ms = timeit(lambda: kernels.stack_reduce(cube, out=out32, **KAPPA))
        report(f"N={n} {h}x{w} kappa-sigma f32 frames", ms, (4 * n + 5) * px)
        # uint16 frames of the same data
        c16 = torch.empty((n, h, w), dtype=torch.int16, device=dev)
        for i in range(n):
            c16[i] = cube[i].clamp(0, 65535).round().to(torch.int32).to(torch.int16)
        del cube
        u16 = c16.view(torch.uint16)
        ms = timeit(lambda: kernels.stack_reduce(u16, out=out32, **KAPPA))
        report(f"N={n} {h}x{w} kappa-sigma u16 frames", ms, (2 * n + 5) * px)
        ms = timeit(lambda: kernels.stack_reduce(u16, out=out32, **MEDMAD))
        report(f"N={n} {h}x{w} medmad u16 frames", ms, (2 * n + 5) * px)
        ms = timeit(lambda: kernels.stack_reduce(u16, out=out32, method="median", maxiters=0, want_nrej=False))
        report(f"N={n} {h}x{w} median u16 frames", ms, (2 * n + 4) * px)
        del c16, u16
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
