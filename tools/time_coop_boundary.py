"""Single-thread register kernel against the lane-cooperative one for kappa-sigma stacks just above N = 100
(dev tool; decided MEANCLIP_COOP_MIN_N in csrc/stack.cu).  Usage: python tools/time_coop_boundary.py [N ...]"""
import sys, torch
sys.path.insert(0, "/root/repo")
import bench
from astrophotography_b200 import kernels
from tools.time_round2 import timeit, report, KAPPA
dev = torch.device("cuda", 0)
h, w = 4096, 4096
for n in ([int(x) for x in sys.argv[1:]] or [104, 112, 128]):
    cube = bench.synth_cube_device(torch, n, h, w, dev, seed=1000)
    out = {}
    for pref in (None, "registers"):
        ms = timeit(lambda: kernels.stack_reduce(cube, out=out, prefer=pref, **KAPPA))
        report(f"N={n} prefer={pref} {kernels.stack_kernel_name(n, prefer=pref, **KAPPA)}", ms, (4 * n + 5) * h * w)
    del cube, out
    torch.cuda.empty_cache()
