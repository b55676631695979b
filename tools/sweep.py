"""BASELINE config 5: throughput sweep over frame count N x frame size, for the median, the reference's
median/MAD clip, the kappa-sigma clip and calibrate-only (device-resident inputs).

    python tools/sweep.py > profiles/rNN_sweep_1gpu.md
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P \
        tools/sweep.py > profiles/rNN_sweep_8gpu.md

Under torchrun every rank owns its row band of every frame (pipeline.row_band), the ranks start together and
the slowest rank's CUDA-event time counts; the table holds the whole job's Mpix-frames/s.
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


WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
dist = None


def say(line):
    if RANK == 0:
        print(line, flush=True)


def timeit(fn, reps):
    fn(); fn()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    ms = _timeit(fn, reps)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def _timeit(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    global dist
    if WORLD > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        saved = os.dup(1)
        os.dup2(2, 1)                       # NCCL announces itself on stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        dist.barrier()
        sys.stdout.flush()
        os.dup2(saved, 1)
    from astrophotography_b200 import pipeline
    g = torch.Generator(device="cuda"); g.manual_seed(5 + RANK)
    say(f"{WORLD} GPU(s), rows of every frame sharded over the ranks; fractions are per GPU, of {PEAK:.0f} GB/s\n")
    say(f"| frame | N | mode | kernel | ms | Mpix-frames/s (all GPUs) | GB/s per GPU | of {PEAK:.0f} GB/s |")
    say("|---|---:|---|---|---:|---:|---:|---:|")
    for (hfull, w) in SIZES:
        r0, r1, _, _ = pipeline.row_band(hfull, WORLD, RANK)
        h = r1 - r0
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
                say(f"| {w}x{hfull} | {n} | {mode} | {name}/{kernels.stack_last_staging()} | {ms:.3f} | "
                    f"{n * hfull * w / ms / 1e3:.0f} | {nbytes / ms / 1e6:.0f} | {nbytes / ms / 1e6 / PEAK:.2f} |")
        # calibrate-only
        raw, bias, dark, flat = cube[0], cube[1], cube[2], cube[3].abs() + 1.0
        cal = torch.empty((h, w), dtype=torch.float32, device="cuda")
        ms = timeit(lambda: kernels.calibrate(raw, bias, dark, flat, 1.0 / 3.0, True, out=cal), 20)
        say(f"| {w}x{hfull} | 1 | calibrate-only (f32 raw) | calibrate_vec4 | {ms:.4f} | {hfull * w / ms / 1e3:.0f} | "
            f"{20 * h * w / ms / 1e6:.0f} | {20 * h * w / ms / 1e6 / PEAK:.2f} |")
        del cube, sub
        torch.cuda.empty_cache()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
