"""Host-buffer pipelines: frames in host memory -> combined master in host memory.

This is the call a user (``ApMasterCal.make_master``) or a binding added to the
reference makes when the frames live in host RAM, and what ``bench.py`` times as
``e2e``: the stack is cut into row bands (the scheme ``ccdproc.combine`` uses for
``mem_limit``, reference ``scripts/ap_combine_darks.py:398,418`` -- but each
frame is read once, not once per band); band ``k+1`` is uploaded on a copy
stream while band ``k`` is reduced on the compute stream (two device buffers,
CUDA events), and each band of the result goes back to pinned host memory as
soon as it is done.  Multi-GPU: one process per GPU, each owning a contiguous
row band of every frame (``row_band``); no data-path collective is needed
because every output pixel depends only on the same pixel of the N frames.
"""
from __future__ import annotations

import numpy as np

from . import _native, kernels


def row_band(nrows: int, world: int, rank: int, halo: int = 0):
    """Contiguous row band ``[r0, r1)`` of rank ``rank`` out of ``world`` and the
    band extended by ``halo`` rows on each side where they exist (bad-pixel
    repair reads a ``deltapix`` halo of the input; the stack needs none).
    Rows are split as evenly as possible, the first ``nrows % world`` ranks
    taking one extra row."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(nrows, world)
    r0 = rank * base + min(rank, extra)
    r1 = r0 + base + (1 if rank < extra else 0)
    return r0, r1, max(0, r0 - halo), min(nrows, r1 + halo)


def pinned_empty(shape, dtype=np.float32):
    """A numpy array backed by page-locked host memory (fast, truly async H2D/D2H)."""
    torch = _native.require_cuda()
    tdtype = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
              np.dtype(np.uint8): torch.uint8, np.dtype(np.uint16): torch.uint16,
              np.dtype(np.int16): torch.int16}[np.dtype(dtype)]
    t = torch.empty(tuple(shape), dtype=tdtype, pin_memory=True)
    if tdtype == torch.uint16:
        return t.view(torch.int16).numpy().view(np.uint16), t
    return t.numpy(), t


class HostStackCombiner:
    """Reusable double-buffered host->device->host stack reducer.

    ``combine(frames)`` takes N host frames (a sequence of (H,W) float32 numpy
    arrays or one (N,H,W) array; pinned memory gives full PCIe speed) and
    returns host arrays ``data`` (+ ``nrej``, ``uncert``, ``allmasked`` when
    requested).  Device buffers are allocated once per instance.
    """

    def __init__(self, n, h, w, method="average", k_lo=5.0, k_hi=5.0, maxiters=1, cen="median",
                 dev="mad_std", out_f64=False, want_nrej=True, want_uncert=False,
                 want_allmasked=False, band_bytes=2 << 30, device=None):
        torch = _native.require_cuda()
        self.torch = torch
        self.n, self.h, self.w = int(n), int(h), int(w)
        self.params = dict(method=method, k_lo=k_lo, k_hi=k_hi, maxiters=maxiters, cen=cen, dev=dev)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.band_rows = int(max(1, min(self.h, band_bytes // (self.n * self.w * 4))))
        self.nbands = (self.h + self.band_rows - 1) // self.band_rows
        odt = torch.float64 if out_f64 else torch.float32
        self.cube = [torch.empty((self.n, self.band_rows, self.w), dtype=torch.float32, device=self.device)
                     for _ in range(2)]
        self.outs = []
        for _ in range(2):
            o = {"data": torch.empty((self.band_rows, self.w), dtype=odt, device=self.device)}
            if want_nrej:
                o["nrej"] = torch.empty((self.band_rows, self.w),
                                        dtype=torch.uint8 if self.n <= 255 else torch.uint16, device=self.device)
            if want_uncert:
                o["uncert"] = torch.empty((self.band_rows, self.w), dtype=odt, device=self.device)
            if want_allmasked:
                o["allmasked"] = torch.empty((self.band_rows, self.w), dtype=torch.uint8, device=self.device)
            self.outs.append(o)
        self.want = dict(want_nrej=want_nrej, want_uncert=want_uncert, want_allmasked=want_allmasked)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.compute_stream = torch.cuda.Stream(device=self.device)
        np_odt = np.float64 if out_f64 else np.float32
        self.host_out = {"data": pinned_empty((self.h, self.w), np_odt)}
        if want_nrej:
            self.host_out["nrej"] = pinned_empty((self.h, self.w), np.uint8 if self.n <= 255 else np.uint16)
        if want_uncert:
            self.host_out["uncert"] = pinned_empty((self.h, self.w), np_odt)
        if want_allmasked:
            self.host_out["allmasked"] = pinned_empty((self.h, self.w), np.uint8)
        self.h2d_bytes = self.n * self.h * self.w * 4
        self.d2h_bytes = sum(arr.nbytes for arr, _ in self.host_out.values())

    def _frame_tensor(self, frames, i):
        torch = self.torch
        f = frames[i]
        if isinstance(f, torch.Tensor):
            return f
        if f.dtype != np.float32 or not f.flags.c_contiguous:
            raise RuntimeError("HostStackCombiner: frames must be C-contiguous float32")
        return torch.from_numpy(f)

    def combine(self, frames):
        torch = self.torch
        if len(frames) != self.n:
            raise RuntimeError(f"HostStackCombiner: expected {self.n} frames, got {len(frames)}")
        host = [self._frame_tensor(frames, i) for i in range(self.n)]
        for f in host:
            if tuple(f.shape) != (self.h, self.w):
                raise RuntimeError("HostStackCombiner: frame shape mismatch")
        uploaded = [torch.cuda.Event() for _ in range(2)]
        reduced = [None, None]
        for b in range(self.nbands):
            buf = b & 1
            r0 = b * self.band_rows
            r1 = min(self.h, r0 + self.band_rows)
            rows = r1 - r0
            with torch.cuda.stream(self.copy_stream):
                if reduced[buf] is not None:          # buffer still being read by band b-2
                    self.copy_stream.wait_event(reduced[buf])
                cube = self.cube[buf]
                for i in range(self.n):
                    cube[i, :rows].copy_(host[i][r0:r1], non_blocking=True)
                uploaded[buf].record(self.copy_stream)
            with torch.cuda.stream(self.compute_stream):
                self.compute_stream.wait_event(uploaded[buf])
                kernels.stack_reduce(self.cube[buf], row0=0, nrows=rows, out=self.outs[buf],
                                     **self.params, **self.want)
                for key, (_, pinned) in self.host_out.items():
                    src = self.outs[buf][key][:rows]
                    dst = pinned[r0:r1]
                    if dst.dtype != src.dtype:        # uint16 pinned buffers are int16-backed
                        src = src.view(dst.dtype)
                    dst.copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.compute_stream)
                reduced[buf] = ev
        self.compute_stream.synchronize()
        return {k: arr for k, (arr, _) in self.host_out.items()}


def combine_sharded(frames, reduce_band=None, dist=None, **combine_kw):
    """Row-band sharded combine across the ranks of ``torch.distributed`` (one process per GPU).

    Every rank holds (or can read) the same N host frames, reduces only its own
    row band (``row_band``) and the bands are gathered on rank 0; there is no
    data-path collective because each output pixel depends only on the same
    pixel of the N frames.  ``reduce_band(list_of_band_arrays)
    -> dict`` defaults to the GPU ``HostStackCombiner``; tests inject the CPU
    oracle to exercise this host logic under the ``gloo`` backend.

    Returns the dict of full-frame host arrays on rank 0 and ``None`` elsewhere.
    """
    import torch
    if dist is None:
        import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    n = len(frames)
    h, w = frames[0].shape
    r0, r1, _, _ = row_band(h, world, rank)
    bands = [np.ascontiguousarray(f[r0:r1]) for f in frames]
    if reduce_band is None:
        def reduce_band(bs):
            comb = HostStackCombiner(n, r1 - r0, w, **combine_kw)
            return {k: np.array(v, copy=True) for k, v in comb.combine(bs).items()}
    local = reduce_band(bands) if r1 > r0 else {}
    if world == 1:
        return local
    max_rows = (h + world - 1) // world
    keys = sorted(local.keys()) if r1 > r0 else None
    key_lists = [None] * world
    dist.all_gather_object(key_lists, (keys, {k: str(local[k].dtype) for k in (keys or [])}))
    keys, dtypes = next(kl for kl in key_lists if kl[0] is not None)
    backend = dist.get_backend()
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    out = {} if rank == 0 else None
    for k in keys:
        dt = np.dtype(dtypes[k])
        pad = np.zeros((max_rows, w), dtype=dt)
        if r1 > r0:
            pad[: r1 - r0] = local[k]
        t = torch.from_numpy(pad.view(np.uint8).reshape(max_rows, w * dt.itemsize)).to(device)   # raw bytes: any dtype, any backend
        recv = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, recv, dst=0)
        if rank == 0:
            full = np.empty((h, w), dtype=dt)
            for rk in range(world):
                a0, a1, _, _ = row_band(h, world, rk)
                part = recv[rk].cpu().numpy().view(dt).reshape(max_rows, w)
                full[a0:a1] = part[: a1 - a0]
            out[k] = full
    return out
