import subprocess, sys
code = '''
import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
from astrophotography_b200 import kernels
h, w, row0, nrows, n = map(int, sys.argv[1:6])
st = np.random.default_rng(1).normal(1000, 12, size=(n, h, w)).astype(np.float32)
cube = torch.from_numpy(st).cuda()
out = kernels.stack_reduce(cube, method="average", k_lo=3.0, k_hi=3.0, maxiters=5, cen="mean", dev="std", row0=row0, nrows=nrows)
torch.cuda.synchronize()
print("ok staging", kernels.stack_last_staging())
'''
for args in [(8,76,3,4,100),(8,75,4,4,100),(8,75,3,4,100),(8,75,0,8,100),(16,75,1,8,64),(8,74,2,4,100),(12,68,2,9,300)]:
    r = subprocess.run([sys.executable, "-c", code, *map(str,args)], capture_output=True, text=True)
    print(args, "pix0*4 % 16 =", (args[2]*args[1]*4) % 16, (r.stdout.strip() or r.stderr.strip().splitlines()[-1])[:100])
