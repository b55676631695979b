"""BASELINE config 5: throughput sweep over frame count N x frame size, for the median, the reference's
median/MAD clip, the kappa-sigma clip and calibrate-only, on one GPU (device-resident inputs).

    python tools/sweep.py > profiles/rNN_sweep_1gpu.md
"""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astrophotography_b200 import kernels

PEAK = 6459.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
SIZES = [(1024, 1024), (1472, 2184), (4096, 4096), (4000, 6000), (6388, 9576)]
NS = [8, 16, 32, 64, 100, 128, 200, 256, 512]
MODES = {
    "median": dict(method="median", maxiters=0, want_nrej=False),
    "medmad 5s x1 (ApMasterCal)": dict(method="average", k_lo=5.0, k_hi=5.0, maxiters=1, cen="median", dev="mad_std"),
    "kappa-sigma 3s x5": dict(method="average", k_lo=3.0, k_hi=3.0, maxiters=5, cen="mean", dev="std"),
}
MAX_BYTES = 60e9


def timeit(fn, reps):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    print(f"| frame | N | mode | kernel | ms | Mpix-frames/s | GB/s | of {PEAK:.0f} GB/s |")
    print("|---|---:|---|---|---:|---:|---:|---:|")
    for (h, w) in SIZES:
        nmax = max(n for n in NS if n * h * w * 4 <= MAX_BYTES)
        cube = torch.empty((nmax, h, w), dtype=torch.float32, device="cuda")
        for i in range(nmax):
            cube[i].normal_(1000.0, 12.0, generator=g)
            hits = torch.rand((h, w), device="cuda", generator=g) < 1e-4
            cube[i][hits] += 5000.0
        for n in NS:
            if n > nmax:
                continue
            sub = cube[:n]
            for mode, kw in MODES.items():
                out = {}
                ms = timeit(lambda: kernels.stack_reduce(sub, out=out, **kw), 3 if n * h * w > 2e9 else 10)
                nbytes = (4 * n + (4 if mode == "median" else 5)) * h * w
                name = kernels.stack_kernel_name(n, **{k: v for k, v in kw.items() if k != "want_nrej"})
                print(f"| {w}x{h} | {n} | {mode} | {name}/{kernels.stack_last_staging()} | {ms:.3f} | "
                      f"{n * h * w / ms / 1e3:.0f} | {nbytes / ms / 1e6:.0f} | {nbytes / ms / 1e6 / PEAK:.2f} |")
        # calibrate-only
        raw, bias, dark, flat = cube[0], cube[1], cube[2], cube[3].abs() + 1.0
        cal = torch.empty((h, w), dtype=torch.float32, device="cuda")
        ms = timeit(lambda: kernels.calibrate(raw, bias, dark, flat, 1.0 / 3.0, True, out=cal), 20)
        print(f"| {w}x{h} | 1 | calibrate-only (f32 raw) | calibrate_vec4 | {ms:.4f} | {h * w / ms / 1e3:.0f} | "
              f"{20 * h * w / ms / 1e6:.0f} | {20 * h * w / ms / 1e6 / PEAK:.2f} |")
        del cube, sub
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
