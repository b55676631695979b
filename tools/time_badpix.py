"""dev tool: fix_badpix / fused calibrate+repair timings for different masks (9576 x 6388)."""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astrophotography_b200 import kernels, synth
dev = torch.device("cuda", 0)
h, w = 6388, 9576
g = torch.Generator(device=dev); g.manual_seed(3)
data = torch.empty((h, w), device=dev).normal_(1000.0, 12.0, generator=g)
raw = torch.randint(0, 65535, (h, w), dtype=torch.int32, device=dev, generator=g).to(torch.int16).view(torch.uint16)
bias = torch.empty((h, w), device=dev).normal_(1000.0, 12.0, generator=g)
dark = torch.empty((h, w), device=dev).normal_(1040.0, 12.0, generator=g)
nflat = torch.empty((h, w), device=dev).normal_(1.0, 0.01, generator=g)
out = torch.empty((h, w), dtype=torch.float32, device=dev)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
full = synth.badpix_mask((h, w), auto_fraction=1e-3)
rnd = (np.random.default_rng(1).random((h, w)) < 1.35e-3).astype(np.uint8)
cols = np.zeros((h, w), np.uint8); cols[:, [1000, 1001, 5000]] = 2
for name, m in (("empty", np.zeros((h, w), np.uint8)), ("random 0.135 %", rnd), ("3 bad columns", cols), ("bench mask", full)):
    md = torch.from_numpy(m).to(dev)
    t1 = timeit(lambda: kernels.fix_badpix(data, md, 2))
    t2 = timeit(lambda: kernels.calibrate_repair(raw, bias, dark, nflat, 1.0 / 3.0, True, mask=md, deltapix=2, out=out))
    print(f"{name:16s} nbad={int(m.astype(bool).sum()):7d}  fix_badpix {t1*1e3:7.1f} us ({9*h*w/t1/1e6/6459:.2f})   fused {t2*1e3:7.1f} us ({19*h*w/t2/1e6/6459:.2f})", flush=True)
t = timeit(lambda: kernels.calibrate(raw, bias, dark, nflat, 1.0 / 3.0, True, out=out))
print(f"calibrate alone {t*1e3:.1f} us; copy_ {timeit(lambda: out.copy_(data))*1e3:.1f} us")
