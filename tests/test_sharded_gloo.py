"""World-size-2 gloo test (CPU) of the row-band sharded combine: partition, per-rank
reduce, gather on rank 0.  The per-band compute is injected (the numpy oracle) because
there is no GPU here; on the GPU box the same driver runs HostStackCombiner per rank."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, h, w, outdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from astrophotography_b200 import pipeline
    from oracle import combine_oracle as C
    rng = np.random.default_rng(77)                       # identical frames on every rank
    frames = [rng.normal(1000, 12, (h, w)).astype(np.float32) for _ in range(n)]
    frames[3][5, 5] = 30000.0

    def reduce_band(bands):
        r = C.combine(np.stack(bands), "average", 3.0, 3.0, 5, "mean", "std", want_uncert=False)
        return {"data": r["data"].astype(np.float32), "nrej": r["nrej"].astype(np.uint16)}

    res = pipeline.combine_sharded(frames, reduce_band=reduce_band, dist=dist)
    if rank == 0:
        np.savez(os.path.join(outdir, "sharded.npz"), **res)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("h", [33, 8])
def test_sharded_combine_world2_gloo(tmp_path, h):
    world, n, w = 2, 9, 17
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, h, w, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "sharded.npz")
    from oracle import combine_oracle as C
    rng = np.random.default_rng(77)
    frames = [rng.normal(1000, 12, (h, w)).astype(np.float32) for _ in range(n)]
    frames[3][5, 5] = 30000.0
    exp = C.combine(np.stack(frames), "average", 3.0, 3.0, 5, "mean", "std", want_uncert=False)
    assert got["data"].shape == (h, w)
    assert np.array_equal(got["data"], exp["data"].astype(np.float32))
    assert np.array_equal(got["nrej"].astype(np.int64), exp["nrej"])
