"""Row-band sharded combine on real GPUs: one process per rank, each uploads only its band, reduces it with the
CUDA kernels and writes its band of the result into the shared host array (pipeline.ShardedStackCombiner).

* world 2 on ONE GPU (gloo carries the segment name and the barriers, both ranks compute on cuda:0): runs on
  the single-GPU test box;
* world 2 / 4 / 8 over NCCL, one GPU per rank: runs wherever that many GPUs are visible.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu

PARAMS = dict(method="average", k_lo=3.0, k_hi=3.0, maxiters=5, cen="mean", dev="std")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _frames(n, h, w, dtype):
    rng = np.random.default_rng(4242)                     # identical frames on every rank
    st = rng.normal(1000, 12, (n, h, w)).astype(np.float32)
    hits = rng.random((n, h, w)) < 0.003
    st[hits] += rng.uniform(500, 30000, int(hits.sum())).astype(np.float32)
    if np.dtype(dtype) == np.uint16:
        return np.clip(np.rint(st), 0, 65535).astype(np.uint16)
    return st


def _worker(rank, world, port, backend, n, h, w, dtype, outdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dev = rank if backend == "nccl" else 0
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from astrophotography_b200 import _native, pipeline
    st = _frames(n, h, w, dtype)
    sc = pipeline.ShardedStackCombiner(n, h, w, dist=dist, dtype=dtype, want_nrej=True, band_bytes=1 << 20, **PARAMS)
    r0, r1 = sc.band_rows()
    launches0 = _native.launch_count()
    res = sc.combine([st[i, r0:r1] for i in range(n)])     # a rank only touches its own rows
    assert _native.launch_count() > launches0 or r1 == r0
    assert all(reg.ok for reg in sc._regs), "the shared result band could not be page-locked"
    if rank == 0:
        np.savez(os.path.join(outdir, "sharded.npz"), **{k: np.array(v) for k, v in res.items()})
    sc.close()
    dist.barrier()
    dist.destroy_process_group()


def _check(tmp_path, n, h, w, dtype):
    from oracle import combine_oracle as C
    got = np.load(tmp_path / "sharded.npz")
    exp = C.combine(_frames(n, h, w, dtype).astype(np.float32), want_uncert=False, **PARAMS)
    assert got["data"].shape == (h, w)
    assert np.array_equal(got["nrej"].astype(np.int64), exp["nrej"])
    ok = np.abs(got["data"].astype(np.float64) - exp["data"]) <= 1e-6 * np.maximum(np.abs(exp["data"]), 12.0)
    assert ok.all()


@pytest.mark.parametrize("dtype", [np.float32, np.uint16])
def test_sharded_combine_two_ranks_one_gpu(cuda, tmp_path, dtype):
    n, h, w = 40, 37, 264                                   # odd row count: bands of 19 and 18 rows
    mp.spawn(_worker, args=(2, _free_port(), "gloo", n, h, w, dtype, str(tmp_path)), nprocs=2, join=True)
    _check(tmp_path, n, h, w, dtype)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_combine_nccl(cuda, tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    n, h, w = 200, 66, 520                                  # BASELINE config 4's frame count, small frames
    mp.spawn(_worker, args=(world, _free_port(), "nccl", n, h, w, np.float32, str(tmp_path)), nprocs=world, join=True)
    _check(tmp_path, n, h, w, np.float32)
