"""dev tool: one profiled launch of the lane-cooperative median (N from argv, default 256) at 4096 x 4096."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from astrophotography_b200 import kernels
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cube = bench.synth_cube_device(torch, n, 4096, 4096, torch.device("cuda", 0), seed=1000)
out = {}
for _ in range(2):
    kernels.stack_reduce(cube, out=out, method="median", maxiters=0, want_nrej=False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
kernels.stack_reduce(cube, out=out, method="median", maxiters=0, want_nrej=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
