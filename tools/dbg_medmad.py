import sys, numpy as np, torch
sys.path.insert(0, '.')
from astrophotography_b200 import kernels
st = np.array([5005.6304, 5006.399, 5015.5317], dtype=np.float32).reshape(3,1,1)
st = np.repeat(st, 4, axis=2)
r = kernels.stack_reduce(torch.from_numpy(st).cuda(), want_uncert=False)
torch.cuda.synchronize()
print(r['data'].cpu().numpy(), r['nrej'].cpu().numpy())
