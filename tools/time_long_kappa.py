"""Time the lane-cooperative kappa-sigma kernels on long stacks (dev tool).
Usage: python tools/time_long_kappa.py [N ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                     # noqa: E402
from astrophotography_b200 import kernels                        # noqa: E402
from tools.time_round2 import timeit, report, KAPPA              # noqa: E402


def main():
    ns = [int(x) for x in sys.argv[1:] if x.isdigit()] or [128, 200, 256, 512]
    dev = torch.device("cuda", 0)
    for (h, w) in ((1472, 2184), (4096, 4096)):
        for n in ns:
            cube = bench.synth_cube_device(torch, n, h, w, dev, seed=1000)
            out = {}
            ms = timeit(lambda: kernels.stack_reduce(cube, out=out, **KAPPA))
            nrej = out["nrej"]
            report(f"N={n} {h}x{w} kappa-sigma ({kernels.stack_kernel_name(n, **KAPPA)})", ms, (4 * n + 5) * h * w,
                   f"nrej>0: {float((nrej.to(torch.int32) > 0).float().mean()):.4f}")
            del cube, out
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
