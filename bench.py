#!/usr/bin/env python3
"""Benchmark of the FITS-reduction hot path on B200 (see BASELINE.json / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--small]

Headline (``value``): Mpix-frames/s of the kappa-sigma-clipped mean stack
(kappa=3, <=5 iterations, BASELINE.json config 3's algorithm) over
100 x (9576 x 6388) float32 frames resident in HBM -- the workload BASELINE.json's
target is quoted on.  One step = one pass of the stack reducer over the whole
cube (24.5 GB, far larger than the 126 MB L2, so no L2 flush is needed between
steps).  ``e2e`` is the same metric through the host-buffer API
(``HostStackCombiner``): frames in pinned host memory, H2D + kernel + D2H of the
result inside the timed region.  ``roofline`` is for the stack kernel
(algorithmic bytes 4*N+5 per pixel over the CUDA-event kernel time, against the
measured copy bandwidth in MEASURED_PEAKS.json).  ``cpu_baseline`` times the
numpy oracle port of the same algorithm on a bounded row sample on the host.
``variants`` reports the other kernels of the path (the reference's own
ApMasterCal median/MAD setting, plain median, calibrate, bad-pixel repair).

Multi-GPU (torchrun, one process per GPU): weak scaling -- every rank reduces
its own 100 x (9576 x 6388) row band of a taller mosaic; no data-path collective
(SURVEY.md section 8e); barrier + max-over-ranks timing.

``--impl reference``: the reference's CPU implementation of the same workload
(the numpy oracle port of ccdproc.combine / astropy.sigma_clip -- the reference
itself is numpy-based and ccdproc/astropy cannot be installed here), all host
threads, each step a bounded row sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HEADLINE = dict(method="average", k_lo=3.0, k_hi=3.0, maxiters=5, cen="mean", dev="std")
FULL = dict(n=100, h=6388, w=9576)          # 100 x (9576 x 6388): W=9576 columns, H=6388 rows
SMALL = dict(n=100, h=512, w=2048)          # --small: quick functional run
METRIC = "Mpix-frames/sec sigma-clip stack (kappa=3, 5 iters) of 100x(9576x6388) f32; % HBM roofline"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------
# synthetic data, generated on the device for the big cube
# ---------------------------------------------------------------------------
def synth_cube_device(torch, n, h, w, device, seed):
    """N dark-like frames: 1000 + N(0,12) + fixed hot pixels + per-frame cosmic hits."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    cube = torch.empty((n, h, w), dtype=torch.float32, device=device)
    nhot = int(round(5e-4 * h * w))
    hot_idx = torch.randint(0, h * w, (nhot,), generator=g, device=device)
    hot_amp = 5000.0 * (0.5 + 0.5 * torch.rand(nhot, generator=g, device=device))
    ncos = int(round(1e-4 * h * w))
    for k in range(n):
        f = cube[k]
        f.normal_(1000.0, 12.0, generator=g)
        f.view(-1).index_add_(0, hot_idx, hot_amp)
        cidx = torch.randint(0, h * w, (ncos,), generator=g, device=device)
        camp = 500.0 + 29500.0 * torch.rand(ncos, generator=g, device=device)
        f.view(-1).index_add_(0, cidx, camp)
    return cube


def synth_cube_host(n, h, w, seed):
    rng = np.random.default_rng(seed)
    cube = rng.normal(1000.0, 12.0, size=(n, h, w)).astype(np.float32)
    nhot = int(round(5e-4 * h * w))
    hot = rng.integers(0, h * w, nhot)
    cube.reshape(n, -1)[:, hot] += (5000.0 * rng.uniform(0.5, 1.0, nhot)).astype(np.float32)
    ncos = int(round(1e-4 * h * w))
    for k in range(n):
        ci = rng.integers(0, h * w, ncos)
        cube[k].reshape(-1)[ci] += rng.uniform(500, 30000, ncos).astype(np.float32)
    return cube


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:            # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
# reference arm / cpu baseline (oracle port on host cores)
# ---------------------------------------------------------------------------
def cpu_combine_rows(cube_rows, params, threads):
    """Oracle port on a (N, rows, W) host sample, split over ``threads`` row bands."""
    from oracle import combine_oracle as C
    rows = cube_rows.shape[1]
    if threads <= 1:
        C.combine(cube_rows, want_uncert=False, **params)
        return
    from concurrent.futures import ThreadPoolExecutor
    bounds = np.linspace(0, rows, threads + 1).astype(int)
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(lambda i: C.combine(cube_rows[:, bounds[i]:bounds[i + 1]], want_uncert=False, **params),
                    [i for i in range(threads) if bounds[i + 1] > bounds[i]]))


def run_reference(args, shape):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, h, w = shape["n"], shape["h"], shape["w"]
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    rows = min(h, max(cores, 4 * cores))           # bounded sample: a few seconds per step
    cube = synth_cube_host(n, rows, w, seed=1000)
    for _ in range(args.warmup):
        cpu_combine_rows(cube, HEADLINE, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_combine_rows(cube, HEADLINE, cores)
    dt = time.perf_counter() - t0
    mpf = n * rows * w / 1e6
    value = mpf * args.steps / dt
    sample = (f"{n} frames x {rows} rows x {w} cols per step ({mpf:.1f} Mpix-frames = {100.0 * rows / h:.2f} % of the frame, "
              f"scaled), numpy oracle port, {cores} threads over row bands")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpix-frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": bench_config(shape), "threads": cores, "rows_per_step": rows,
        "cpu_baseline": {"value": value, "unit": "Mpix-frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpix-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def bench_config(shape):
    """The ``config`` object, identical in the GPU arm and the reference arm (the driver compares them)."""
    return {"workload": workload_name(shape), **HEADLINE,
            "l2": "inputs (24.5 GB per GPU) >> L2 (126 MB): no flush needed",
            "reference_arm": "numpy oracle port of ccdproc.combine / astropy.sigma_clip (oracle/combine_oracle.py), "
                             "all host threads over row bands, 4 rows per thread per step, scaled to the full frame"}


def recorded_traffic(kname, shape):
    """DRAM bytes per launch of the dominant kernel from THIS round's ncu capture (profiles/r02_traffic.json,
    refreshed with the capture); None when this kernel / shape has no capture -- never a stale constant."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)["traffic_bytes"].get(f"{kname}@{shape['n']}x{shape['h']}x{shape['w']}")
    except (OSError, KeyError, ValueError):
        return None


def workload_name(shape):
    return f"kappa-sigma-clipped mean stack {shape['n']}x({shape['w']}x{shape['h']}) float32"


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def time_steps(torch, fn, steps, warmup, dist=None):
    """CUDA-event timing on the current stream; barrier + sync on both sides; max over ranks."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    return ms


def run_gpu(args, shape):
    import torch
    from astrophotography_b200 import _native, kernels, pipeline, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL announces its version on stdout at communicator creation: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    device = torch.device("cuda", local_rank)
    n, h, w = shape["n"], shape["h"], shape["w"]
    mpix = h * w / 1e6
    hbm_peak, peak_src = peaks()

    # one equally spaced cube; the variants also time the 200-frame stack of BASELINE config 4
    n_alloc = n if (args.no_variants or args.small) else max(n, 200)
    cube_all = synth_cube_device(torch, n_alloc, h, w, device, seed=1000 + rank)
    cube = cube_all[:n]
    out = {"data": torch.empty((h, w), dtype=torch.float32, device=device),
           "nrej": torch.empty((h, w), dtype=torch.uint8, device=device)}
    torch.cuda.synchronize()

    def step():
        kernels.stack_reduce(cube, out=out, **HEADLINE)

    sampler = ClockSampler(local_rank)
    launches0 = _native.launch_count()
    sampler.start()
    ms = time_steps(torch, step, args.steps, args.warmup, dist)
    clocks = sampler.stop()
    headline_result = out["data"].clone()
    launches = (_native.launch_count() - launches0) * args.steps // (args.steps + args.warmup)
    ms_per_step = ms / args.steps
    value = world * n * mpix / (ms_per_step * 1e-3)
    alg_bytes = (4 * n + 5) * h * w
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
    STAGING = {-1: "", 0: "/direct-loads", 1: "/bulk-copy-cta", 2: "/cp.async-warp", 3: "/tensormap-tma-warp",
               4: "/lane-split-cp.async", 5: "/warp-coop-swizzled-tensormap-tma"}
    kname = kernels.stack_kernel_name(n, **HEADLINE) + STAGING.get(kernels.stack_last_staging(), "")

    # ---- other kernels of the path (rank 0 reporting only; short runs) ----
    variants = {}

    def add_variant(name, fn, units_mpix, nbytes, steps=3, unit="Mpix-frames/s"):
        vms = time_steps(torch, fn, steps, 2) / steps
        variants[name] = {"value": units_mpix / (vms * 1e-3), "unit": unit, "ms": vms,
                          "hbm_gbs": nbytes / (vms * 1e-3) / 1e9,
                          "frac_of_peak": nbytes / (vms * 1e-3) / 1e9 / hbm_peak}

    if not args.no_variants:
        ref = dict(method="average", k_lo=5.0, k_hi=5.0, maxiters=1, cen="median", dev="mad_std")
        add_variant("stack_medmad_5sigma_1pass_apmastercal[%s]" % kernels.stack_kernel_name(n, **ref),
                    lambda: kernels.stack_reduce(cube, out=out, **ref), n * mpix, alg_bytes)
        med = dict(method="median", maxiters=0)
        add_variant("stack_median[%s]" % kernels.stack_kernel_name(n, **med),
                    lambda: kernels.stack_reduce(cube, out=out, want_nrej=False, **med), n * mpix, (4 * n + 4) * h * w)
        add_variant("stack_kappa_sigma_registers_direct_loads[%s]" % kernels.stack_kernel_name(n, prefer="registers", **HEADLINE),
                    lambda: kernels.stack_reduce(cube, out=out, prefer="registers_direct", **HEADLINE), n * mpix, alg_bytes)
        add_variant("stack_kappa_sigma_registers_cpasync_staged[%s]" % kernels.stack_kernel_name(n, prefer="registers", **HEADLINE),
                    lambda: kernels.stack_reduce(cube, out=out, prefer="registers_cpasync", **HEADLINE), n * mpix, alg_bytes)
        add_variant("stack_kappa_sigma_registers_tensormap_staged[%s]" % kernels.stack_kernel_name(n, prefer="registers", **HEADLINE),
                    lambda: kernels.stack_reduce(cube, out=out, prefer="registers_tensormap", **HEADLINE), n * mpix, alg_bytes)
        add_variant("stack_kappa_sigma_registers_tma_staged[%s]" % kernels.stack_kernel_name(n, prefer="registers", **HEADLINE),
                    lambda: kernels.stack_reduce(cube, out=out, prefer="registers_tma", **HEADLINE), n * mpix, alg_bytes)
        add_variant("stack_kappa_sigma_shared[%s]" % kernels.stack_kernel_name(n, prefer="shared", **HEADLINE),
                    lambda: kernels.stack_reduce(cube, out=out, prefer="shared", **HEADLINE), n * mpix, alg_bytes)
        add_variant("stack_median_with_mad_uncertainty[%s]" % kernels.stack_kernel_name(n, want_uncert=True, **med),
                    lambda: kernels.stack_reduce(cube, out=out, want_nrej=False, want_uncert=True, **med), n * mpix, (4 * n + 8) * h * w)
        # uint16 frames (what BITPIX=16 camera files hold): the same stack rounded, 2 bytes per pixel-frame
        c16 = torch.empty((n, h, w), dtype=torch.int16, device=device)
        for i in range(n):
            c16[i] = cube[i].clamp(0, 65535).round().to(torch.int32).to(torch.int16)
        u16 = c16.view(torch.uint16)
        add_variant("stack_kappa_sigma_u16_frames[%s]" % kernels.stack_kernel_name(n, **HEADLINE),
                    lambda: kernels.stack_reduce(u16, out=out, **HEADLINE), n * mpix, (2 * n + 5) * h * w)
        add_variant("stack_medmad_5sigma_1pass_apmastercal_u16_frames[%s]" % kernels.stack_kernel_name(n, **ref),
                    lambda: kernels.stack_reduce(u16, out=out, **ref), n * mpix, (2 * n + 5) * h * w)
        del c16, u16
        for nn in (16, 30, 64, 128, 200):
            if nn > n_alloc:
                continue
            cn = cube_all[:nn]
            kernels.stack_reduce(cn, out=out, **HEADLINE)
            tag = kernels.stack_kernel_name(nn, **HEADLINE) + STAGING.get(kernels.stack_last_staging(), "")
            add_variant("stack_kappa_sigma_N%d[%s]" % (nn, tag),
                        lambda cn=cn: kernels.stack_reduce(cn, out=out, **HEADLINE), nn * mpix, (4 * nn + 5) * h * w)
        # BASELINE config 2 end to end on the device: master bias + dark + flat (30 frames of 4096x4096 each, the
        # reference's ApMasterCal setting: 5-sigma median/MAD clip, one pass, mean), flat normalisation, then
        # calibration and bad-pixel repair of one uint16 science frame
        if h >= 4096 and w >= 4096:
            h2 = w2 = 4096
            c2 = cube_all[:90].reshape(-1)[: 90 * h2 * w2].view(90, h2, w2)
            raw2 = torch.randint(0, 65535, (h2, w2), dtype=torch.int32, device=device).to(torch.int16).view(torch.uint16)
            mask2 = torch.from_numpy(synth.badpix_mask((h2, w2), auto_fraction=1e-3)).to(device)
            m_out = [{"data": torch.empty((h2, w2), dtype=torch.float32, device=device),
                      "nrej": torch.empty((h2, w2), dtype=torch.uint8, device=device)} for _ in range(3)]
            cal2 = torch.empty((h2, w2), dtype=torch.float32, device=device)

            def config2():
                for k in range(3):
                    kernels.stack_reduce(c2[30 * k: 30 * k + 30], out=m_out[k], **ref)
                nflat2, _ = kernels.flat_normalise(m_out[2]["data"])
                kernels.calibrate(raw2, m_out[0]["data"], m_out[1]["data"], nflat2, 1.0 / 3.0, True, out=cal2)
                kernels.fix_badpix(cal2, mask2, 2)

            px2 = h2 * w2
            add_variant("config2_masters_3x30x4096x4096_medmad_then_calibrate_repair[%s]" % kernels.stack_kernel_name(30, **ref),
                        config2, 90 * px2 / 1e6, (3 * (4 * 30 + 5) + 8 + 18 + 9) * px2)
            del c2, raw2, mask2, m_out, cal2
        raw = torch.randint(0, 65535, (h, w), dtype=torch.int32, device=device).to(torch.int16).view(torch.uint16)
        bias, dark = cube[0], cube[1]
        flat = (30000.0 * (1 + 0.01 * torch.randn((h, w), device=device))).contiguous()
        nflat, _ = kernels.flat_normalise(flat)
        cal = torch.empty((h, w), dtype=torch.float32, device=device)
        add_variant("calibrate_u16_fused", lambda: kernels.calibrate(raw, bias, dark, nflat, 1.0 / 3.0, True, out=cal),
                    mpix, 18 * h * w, steps=5, unit="Mpix/s")
        rawf = cube[2]
        add_variant("calibrate_f32_fused", lambda: kernels.calibrate(rawf, bias, dark, nflat, 1.0 / 3.0, True, out=cal),
                    mpix, 20 * h * w, steps=5, unit="Mpix/s")
        mask = torch.from_numpy(synth.badpix_mask((h, w), auto_fraction=1e-3)).to(device)
        add_variant("fix_badpix_dp2", lambda: kernels.fix_badpix(cal, mask, 2), mpix, 9 * h * w, steps=5, unit="Mpix/s")
        add_variant("flat_norm_nanmean", lambda: kernels.flat_norm(flat), mpix, 4 * h * w, steps=5, unit="Mpix/s")
        # the batch driver's launch: calibrate + repair fused (no intermediate image), FITS byte order out
        add_variant("calibrate_repair_fused_u16_dp2", lambda: kernels.calibrate_repair(raw, bias, dark, nflat, 1.0 / 3.0, True,
                                                                                      mask=mask, deltapix=2, out=cal, out_big_endian=True),
                    mpix, 19 * h * w, steps=5, unit="Mpix/s")
        add_variant("calibrate_then_repair_two_launches_u16_dp2",
                    lambda: kernels.fix_badpix(kernels.calibrate(raw, bias, dark, nflat, 1.0 / 3.0, True, out=cal), mask, 2),
                    mpix, 27 * h * w, steps=5, unit="Mpix/s")
        del raw, flat, nflat, cal, mask
        # the batch driver end to end through FILES (rank 0): ApCalibrate.calibrate_many (masters resident, three
        # streams, fused calibrate + repair, FITS byte order produced on the GPU) against one ApCalibrate.calibrate
        # call per frame with the same resident masters; 4096 x 4096 uint16 frames in a temporary directory
        if rank == 0 and not args.small:
            variants.update(batch_driver_variant(torch, device))

    # ---- end to end through the host-buffer API (pinned host frames) ----
    e2e = None
    if not args.no_e2e:
        avail = _mem_available_bytes()
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        need = n * h * w * 4
        n_host = n
        if avail is not None and need * local_world > 0.55 * avail:
            n_host = max(4, int(0.55 * avail / local_world / (h * w * 4)))
        host_frames, keep = [], []
        for i in range(min(n, n_host)):
            arr, t = pipeline.pinned_empty((h, w), np.float32)
            t.copy_(cube[i])
            host_frames.append(arr)
            keep.append(t)
        torch.cuda.synchronize()
        frames = [host_frames[i % len(host_frames)] for i in range(n)]
        comb = pipeline.HostStackCombiner(n, h, w, **HEADLINE, band_bytes=2 << 30, device=device)
        esteps = max(1, min(args.steps, 3))
        comb.combine(frames)                                     # warm-up
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            res = comb.combine(frames)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        # check the e2e result against the device-resident result
        same = bool(np.array_equal(res["data"], headline_result.cpu().numpy(), equal_nan=True)) if n_host == n else None
        e2e = {"value": world * n * mpix * esteps / dt, "unit": "Mpix-frames/s",
               "h2d_bytes_per_step": comb.h2d_bytes, "d2h_bytes_per_step": comb.d2h_bytes,
               "steps": esteps, "ms_per_step": 1e3 * dt / esteps, "api": "pipeline.HostStackCombiner.combine",
               "host_frames_distinct": len(host_frames), "matches_device_path": same,
               "pcie_gbs": comb.h2d_bytes * esteps / dt / 1e9}
        del comb, keep, host_frames, frames, res
        _release_pinned_cache(torch)

    # ---- the same through uint16 host frames (raw BITPIX=16 frames: half the PCIe and HBM bytes) ----
    e2e_u16 = None
    if not args.no_e2e:
        host16, keep16 = [], []
        for i in range(min(n, n_host)):
            arr, t = pipeline.pinned_empty((h, w), np.uint16)
            t.view(torch.int16).copy_(cube[i].clamp(0, 65535).round().to(torch.int32).to(torch.int16))
            host16.append(arr)
            keep16.append(t)
        torch.cuda.synchronize()
        frames16 = [host16[i % len(host16)] for i in range(n)]
        comb16 = pipeline.HostStackCombiner(n, h, w, **HEADLINE, band_bytes=2 << 30, device=device, dtype=np.uint16)
        comb16.combine(frames16)                                 # warm-up
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            comb16.combine(frames16)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e_u16 = {"value": world * n * mpix * esteps / dt, "unit": "Mpix-frames/s",
                   "h2d_bytes_per_step": comb16.h2d_bytes, "d2h_bytes_per_step": comb16.d2h_bytes,
                   "steps": esteps, "ms_per_step": 1e3 * dt / esteps,
                   "api": "pipeline.HostStackCombiner.combine(dtype=uint16)",
                   "frames": "the same frames rounded to uint16 (what a BITPIX=16 camera file holds)",
                   "pcie_gbs": comb16.h2d_bytes * esteps / dt / 1e9}
        del comb16, keep16, host16, frames16
        _release_pinned_cache(torch)

    # ---- strong scaling: ONE stack of BASELINE config 4 (200 frames of 9576x6388 float32) row-sharded over the
    # ranks through the product API; every rank uploads only its band, the result lands in one shared host array ----
    strong = None
    if not args.no_strong:
        ns = 200 if not args.small else 24
        sc = pipeline.ShardedStackCombiner(ns, h, w, dist=dist, dtype=np.float32, want_nrej=True, band_bytes=2 << 30,
                                           **HEADLINE)
        r0, r1 = sc.band_rows()
        brows = r1 - r0
        avail = _mem_available_bytes()
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        per_frame = brows * w * 4
        n_dist = ns
        if avail is not None and ns * per_frame * local_world > 0.5 * avail:
            n_dist = max(4, int(0.5 * avail / local_world / max(per_frame, 1)))
        bands, keepb = [], []
        for i in range(min(ns, n_dist)):
            arr, t = pipeline.pinned_empty((brows, w), np.float32)
            t.copy_(cube_all[i % cube_all.shape[0], r0:r1])
            bands.append(arr)
            keepb.append(t)
        torch.cuda.synchronize()
        band_frames = [bands[i % len(bands)] for i in range(ns)]
        sc.combine(band_frames)                                   # warm-up (ends with a barrier)
        ssteps = max(1, min(args.steps, 2))
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(ssteps):
            sres = sc.combine(band_frames)                        # H2D of the band + reduce + D2H + barrier
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        h2d_all = ns * h * w * 4
        full_ok = None
        if rank == 0:
            full_ok = bool(np.isfinite(sres["data"]).all()) and sres["data"].shape == (h, w)
        strong = {"scaling": "strong", "workload": f"kappa-sigma-clipped mean stack {ns}x({w}x{h}) float32, rows sharded over {world} GPU(s)",
                  "value": ns * mpix * ssteps / dt, "unit": "Mpix-frames/s", "n_gpus": world,
                  "ms_per_step": 1e3 * dt / ssteps, "steps": ssteps,
                  "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": 5 * h * w,
                  "h2d_gbs_aggregate": h2d_all * ssteps / dt / 1e9, "h2d_gbs_per_gpu": h2d_all * ssteps / dt / 1e9 / world,
                  "rows_per_rank": brows, "host_frames_distinct": len(bands), "result_complete_on_rank0": full_ok,
                  "api": "pipeline.ShardedStackCombiner.combine (band upload, reduce, D2H into the shared host result, barrier)"}
        del sres, band_frames, bands, keepb
        sc.close()
        _release_pinned_cache(torch)

    # ---- CPU baseline: oracle port, single thread, bounded sample (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rows = min(h, 96)
        sample = cube[:, :rows].cpu().numpy()
        t0 = time.perf_counter()
        cpu_combine_rows(sample, HEADLINE, 1)
        dt = time.perf_counter() - t0
        cpu = {"value": n * rows * w / 1e6 / dt, "unit": "Mpix-frames/s", "cores": 1, "kind": "port",
               "sample": f"{n} frames x {rows} rows x {w} cols ({n * rows * w / 1e6:.1f} Mpix-frames) of the same cube, "
                         f"numpy oracle port (oracle/combine_oracle.py), 1 thread, {dt:.1f} s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mpix-frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(shape),
            "kernel_info": {"kernel": kname, "per_gpu": f"{n}x({w}x{h})"},
            "clocks": clocks, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": recorded_traffic(kname, shape), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel": kname,
                         "step": "the kernel above + stack_marked_kernel, the scan launch that finishes the pixels it "
                                 "marked (DESIGN 3.3b): `achieved` divides by the time of both, `traffic` is the sum of both"},
            "e2e": e2e, "e2e_u16": e2e_u16, "strong": strong, "cpu_baseline": cpu, "variants": variants,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def batch_driver_variant(torch, device, nframes=8, shape=(4096, 4096)):
    import shutil
    import tempfile
    import astrophotography_b200 as ap
    from astrophotography_b200 import fitsio, synth
    tmp = tempfile.mkdtemp(prefix="apgpu_bench_")
    try:
        rng = np.random.default_rng(3)
        hdr = fitsio.new_header
        fitsio.write_image(os.path.join(tmp, "mbias.fits"), rng.normal(1000, 5, shape).astype(np.float32), hdr({}))
        fitsio.write_image(os.path.join(tmp, "mdark.fits"), rng.normal(1040, 6, shape).astype(np.float32), hdr({"EXPTIME": 900.0}))
        fitsio.write_image(os.path.join(tmp, "mflat.fits"), synth.flat_frame(shape), hdr({}))
        fitsio.write_image(os.path.join(tmp, "mask.fits"), synth.badpix_mask(shape, auto_fraction=1e-3), hdr({}))
        raws = []
        base = synth.science_frame(shape, nstars=20)
        for k in range(nframes):
            path = os.path.join(tmp, f"raw{k}.fits")
            fitsio.write_image(path, np.roll(base, 17 * k, axis=1), hdr({"EXPTIME": 300.0, "PEDESTAL": -100}))
            raws.append(path)
        cal = ap.ApCalibrate(os.path.join(tmp, "mbias.fits"), os.path.join(tmp, "mdark.fits"), os.path.join(tmp, "mflat.fits"),
                             os.path.join(tmp, "mask.fits"), "ERROR", True)
        outs = [os.path.join(tmp, f"cal{k}.fits") for k in range(nframes)]
        cal.calibrate_many(raws[:2], outs[:2], 2)                     # warm-up (page cache, pinned buffers)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cal.calibrate_many(raws, outs, 2)
        torch.cuda.synchronize()
        t_many = time.perf_counter() - t0
        t0 = time.perf_counter()
        for r, o in zip(raws, outs):
            cal.calibrate(r, o, 2, None, False)
        torch.cuda.synchronize()
        t_one = time.perf_counter() - t0
        mpix = shape[0] * shape[1] / 1e6
        note = f"{nframes} uint16 frames of {shape[1]}x{shape[0]} read from and written to FITS files in a temporary directory"
        return {
            "batch_driver_calibrate_many_files": {"value": nframes / t_many, "unit": "frames/s", "ms": 1e3 * t_many / nframes,
                                                  "mpix_per_s": nframes * mpix / t_many, "note": note},
            "per_frame_calibrate_files_same_resident_masters": {"value": nframes / t_one, "unit": "frames/s",
                                                                "ms": 1e3 * t_one / nframes, "mpix_per_s": nframes * mpix / t_one,
                                                                "note": note},
        }
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def _release_pinned_cache(torch):
    """Give cached page-locked blocks back to the OS between the host-buffer sections."""
    fn = getattr(torch._C, "_host_emptyCache", None)
    if fn is not None:
        fn()


def _mem_available_bytes():
    try:
        with open("/proc/meminfo") as f:
            for ln in f:
                if ln.startswith("MemAvailable:"):
                    return int(ln.split()[1]) * 1024
    except OSError:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--small", action="store_true", help="small cube (functional check, not a valid bench number)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the row-sharded 200-frame strong-scaling record")
    args = ap.parse_args()
    shape = SMALL if args.small else FULL
    if args.impl == "reference":
        run_reference(args, shape)
    else:
        run_gpu(args, shape)


if __name__ == "__main__":
    main()
