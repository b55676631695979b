"""Host-to-device copy ceiling of one box (dev tool; run under torchrun with 1..8 ranks, one GPU each).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/h2d_probe.py

Every rank copies the same amount of page-locked host memory to its GPU, all ranks at once (barrier before,
device events, max over ranks), for a few variants: one copy stream / two copy streams, buffers allocated
after binding the process to the CPUs that nvidia-smi reports as local to its GPU, write-combined memory.
Rank 0 prints one JSON line per variant with the per-GPU and aggregate GB/s, plus the box's NUMA layout."""
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

GB = 1 << 30


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as exc:      # noqa: BLE001
        return f"<{exc}>"


def local_cpus(index):
    """CPU affinity list that `nvidia-smi topo -m` prints for GPU `index`, e.g. '0-31,64-95' -> set of ints."""
    out = sh("nvidia-smi topo -m")
    for ln in out.splitlines():
        parts = ln.split()
        if parts and parts[0] == f"GPU{index}":
            for tok in parts[1:]:
                if tok[0].isdigit() and ("-" in tok or "," in tok):
                    cpus = set()
                    for rng in tok.split(","):
                        a, _, b = rng.partition("-")
                        cpus.update(range(int(a), int(b or a) + 1))
                    return cpus
    return None


def timed_copy(dsts, srcs, streams, reps):
    torch.cuda.synchronize()
    dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = [torch.cuda.Event(enable_timing=True) for _ in streams]
    e0.record()
    for s in streams:
        s.wait_event(e0)
    for _ in range(reps):
        for k, (d, src) in enumerate(zip(dsts, srcs)):
            with torch.cuda.stream(streams[k % len(streams)]):
                d.copy_(src, non_blocking=True)
    for s, e in zip(streams, e1):
        e.record(s)
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(e) for e in e1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    saved = os.dup(1)
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    dist.barrier()
    sys.stdout.flush()
    os.dup2(saved, 1)
    nchunk, chunk = 16, GB // 4                    # 16 x 256 MiB = 4 GiB per rank and repetition
    dsts = [torch.empty(chunk, dtype=torch.uint8, device="cuda") for _ in range(nchunk)]
    if rank == 0:
        print(json.dumps({"probe": "layout", "world": world, "numactl": sh("numactl -H 2>/dev/null | head -20"),
                          "lscpu": sh("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'"),
                          "topo": sh("nvidia-smi topo -m"), "meminfo": sh("grep -E 'MemTotal|MemAvailable' /proc/meminfo")}))
    results = []
    for variant in ("default", "two_streams", "bound_to_local_cpus", "bound_two_streams"):
        if variant.startswith("bound"):
            cpus = local_cpus(lr)
            if cpus:
                try:
                    os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or os.sched_getaffinity(0))
                except OSError:
                    pass
        srcs = [torch.empty(chunk, dtype=torch.uint8, pin_memory=True) for _ in range(nchunk)]
        for sbuf in srcs:
            sbuf.fill_(1)                           # first touch by this (possibly bound) process
        streams = [torch.cuda.Stream() for _ in range(2 if "two" in variant else 1)]
        timed_copy(dsts, srcs, streams, 1)
        reps = 3
        ms = timed_copy(dsts, srcs, streams, reps)
        gbs = nchunk * chunk * reps / ms / 1e6
        results.append({"probe": "h2d", "variant": variant, "world": world, "gbs_per_gpu": gbs, "gbs_aggregate": gbs * world,
                        "bytes_per_rank": nchunk * chunk * reps})
        del srcs
        fn = getattr(torch._C, "_host_emptyCache", None)
        if fn:
            fn()
    if rank == 0:
        for r in results:
            print(json.dumps(r))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
