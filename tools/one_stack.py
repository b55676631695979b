"""Run one stack_reduce per parameter set on a synthetic cube (to be profiled: ncu -k regex:stack_).
Usage: python tools/one_stack.py N H W [kappa|medmad|median] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                     # noqa: E402
from astrophotography_b200 import kernels                        # noqa: E402
from tools.time_round2 import KAPPA, MEDMAD                      # noqa: E402

n, h, w = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
what = sys.argv[4] if len(sys.argv) > 4 else "kappa"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
kw = {"kappa": KAPPA, "medmad": MEDMAD, "median": dict(method="median", maxiters=0, want_nrej=False)}[what]
cube = bench.synth_cube_device(torch, n, h, w, torch.device("cuda", 0), seed=1000)
out = {}
for _ in range(reps):
    kernels.stack_reduce(cube, out=out, **kw)
torch.cuda.synchronize()
print("done", kernels.stack_kernel_name(n, **{k: v for k, v in kw.items() if k != "want_nrej"}))
