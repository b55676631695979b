"""Host-buffer pipelines: frames in host memory -> combined master in host memory.

This is the call a user (``ApMasterCal.make_master``) or a binding added to the
reference makes when the frames live in host RAM, and what ``bench.py`` times as
``e2e``: the stack is cut into row bands (the scheme ``ccdproc.combine`` uses for
``mem_limit``, reference ``scripts/ap_combine_darks.py:398,418`` -- but each
frame is read once, not once per band); band ``k+1`` is uploaded on a copy
stream while band ``k`` is reduced on the compute stream (two device buffers,
CUDA events), and each band of the result goes back to pinned host memory as
soon as it is done.  Frames may be float32 or raw uint16 (2 bytes per pixel over
PCIe and HBM; converted inside the kernels' load phase).

Multi-GPU (``ShardedStackCombiner``): one process per GPU, each owning a contiguous
row band of every frame (``row_band``).  A rank uploads only its band, reduces it
and writes its band of the result straight into ONE host array shared by all
ranks (POSIX shared memory, page-locked in every process): no device round trip,
no data-path collective -- every output pixel depends only on the same pixel of
the N frames.  ``torch.distributed`` carries only the name
of the shared segment and the barriers.
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


_TORCH_DTYPES = None


def _torch_dtype(torch, dtype):
    global _TORCH_DTYPES
    if _TORCH_DTYPES is None:
        _TORCH_DTYPES = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
                         np.dtype(np.uint8): torch.uint8, np.dtype(np.uint16): torch.uint16,
                         np.dtype(np.int16): torch.int16}
    return _TORCH_DTYPES[np.dtype(dtype)]


def pinned_empty(shape, dtype=np.float32):
    """A numpy array backed by page-locked host memory (fast, truly async H2D/D2H).
    Returns ``(array, tensor)``; the tensor owns the memory."""
    torch = _native.require_cuda()
    tdtype = _torch_dtype(torch, dtype)
    t = torch.empty(tuple(shape), dtype=tdtype, pin_memory=True)
    if tdtype == torch.uint16:
        return t.view(torch.int16).numpy().view(np.uint16), t
    return t.numpy(), t


def host_tensor(torch, arr):
    """A CPU tensor sharing memory with a C-contiguous numpy array (uint16 arrays are
    wrapped as int16: only the bytes matter to a copy)."""
    if isinstance(arr, torch.Tensor):
        return arr
    if not arr.flags.c_contiguous:
        raise RuntimeError("host frames must be C-contiguous")
    if arr.dtype == np.uint16:
        return torch.from_numpy(arr.view(np.int16))
    return torch.from_numpy(arr)


class HostStackCombiner:
    """Reusable double-buffered host->device->host stack reducer.

    ``combine(frames)`` takes N host frames -- a sequence of (H,W) numpy arrays or one
    (N,H,W) array, float32 or uint16 as declared by ``dtype`` (pinned memory gives full
    PCIe speed) -- and returns host arrays ``data`` (+ ``nrej``, ``uncert``, ``allmasked``
    when requested).  Device buffers are allocated once per instance.  ``host_out`` may
    supply the result arrays (e.g. slices of a shared segment that the caller page-locked);
    by default they are pinned arrays owned by the instance.
    """

    def __init__(self, n, h, w, method="average", k_lo=5.0, k_hi=5.0, maxiters=1, cen="median",
                 dev="mad_std", out_f64=False, want_nrej=True, want_uncert=False,
                 want_allmasked=False, band_bytes=2 << 30, device=None, dtype=np.float32,
                 u16_format="native", host_out=None):
        torch = _native.require_cuda()
        self.torch = torch
        self.n, self.h, self.w = int(n), int(h), int(w)
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float32), np.dtype(np.uint16)):
            raise RuntimeError(f"HostStackCombiner: frames must be float32 or uint16, not {self.dtype}")
        self.u16_format = u16_format
        self.params = dict(method=method, k_lo=k_lo, k_hi=k_hi, maxiters=maxiters, cen=cen, dev=dev)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        item = self.dtype.itemsize
        rows = int(max(1, min(self.h, band_bytes // (self.n * self.w * item))))
        if self.dtype == np.uint16 and rows < self.h:
            # a uint16 band must start on a 16-byte boundary to keep the tensor-map kernels (8 pixels)
            step = 8 // int(np.gcd(8, self.w))
            rows = max(step, rows - rows % step)
        self.band_rows = rows
        self.nbands = (self.h + self.band_rows - 1) // self.band_rows
        odt = torch.float64 if out_f64 else torch.float32
        cdt = torch.float32 if self.dtype == np.float32 else torch.int16
        self.cube = [torch.empty((self.n, self.band_rows, self.w), dtype=cdt, device=self.device) for _ in range(2)]
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
        spec = {"data": np_odt}
        if want_nrej:
            spec["nrej"] = np.uint8 if self.n <= 255 else np.uint16
        if want_uncert:
            spec["uncert"] = np_odt
        if want_allmasked:
            spec["allmasked"] = np.uint8
        self.out_spec = spec
        self._keep = []
        self.host_out = {}
        for key, dt in spec.items():
            if host_out is not None:
                arr = host_out[key]
                if arr.shape != (self.h, self.w) or arr.dtype != np.dtype(dt) or not arr.flags.c_contiguous:
                    raise RuntimeError(f"HostStackCombiner: host_out[{key!r}] must be a C-contiguous {(self.h, self.w)} {np.dtype(dt)} array")
                self.host_out[key] = (arr, host_tensor(torch, arr))
            else:
                arr, t = pinned_empty((self.h, self.w), dt)
                self.host_out[key] = (arr, t.view(torch.int16) if t.dtype == torch.uint16 else t)
        self.h2d_bytes = self.n * self.h * self.w * item
        self.d2h_bytes = sum(arr.nbytes for arr, _ in self.host_out.values())

    def _frame_tensor(self, frames, i):
        f = frames[i]
        if isinstance(f, self.torch.Tensor):
            return f
        if f.dtype != self.dtype or not f.flags.c_contiguous:
            raise RuntimeError(f"HostStackCombiner: frames must be C-contiguous {self.dtype}")
        return host_tensor(self.torch, f)

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
                                     u16_format=self.u16_format, **self.params, **self.want)
                for key, (_, pinned) in self.host_out.items():
                    src = self.outs[buf][key][:rows]
                    dst = pinned[r0:r1]
                    if dst.dtype != src.dtype:        # uint16 host arrays are int16-backed
                        src = src.view(dst.dtype)
                    dst.copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.compute_stream)
                reduced[buf] = ev
        self.compute_stream.synchronize()
        return {k: arr for k, (arr, _) in self.host_out.items()}


# ---------------------------------------------------------------------------
# page-locking memory that torch did not allocate (shared segments)
# ---------------------------------------------------------------------------
class _HostRegistration:
    """cudaHostRegister over a numpy array's memory for the life of the object."""

    def __init__(self, torch, arr):
        self.rt = torch.cuda.cudart()
        self.ptr = arr.__array_interface__["data"][0]
        self.nbytes = arr.nbytes
        self.ok = False
        if self.nbytes:
            err = self.rt.cudaHostRegister(self.ptr, self.nbytes, 0)
            code = int(getattr(err, "value", err))
            self.ok = code == 0
            if not self.ok:                         # still correct, only slower (staged copies)
                self.error = code

    def close(self):
        if self.ok:
            self.rt.cudaHostUnregister(self.ptr)
            self.ok = False


class SharedHostArrays:
    """The same named host arrays mapped by every process of one node (POSIX shared memory).

    Rank 0 creates the segment, the other ranks attach by name (``create_or_attach`` exchanges the
    name through ``torch.distributed``); ``arrays[key]`` are numpy views.  The creating process unlinks
    the segment on ``close``."""

    def __init__(self, spec, shape, name=None, create=True, shared_tracker=False):
        from multiprocessing import shared_memory
        self.spec = {k: np.dtype(v) for k, v in spec.items()}
        self.shape = tuple(shape)
        npix = int(np.prod(self.shape))
        offs, total = {}, 0
        for k, dt in self.spec.items():
            total = (total + 4095) // 4096 * 4096          # page-aligned planes
            offs[k] = total
            total += npix * dt.itemsize
        total = max(total, 1)
        if create:
            self.shm = shared_memory.SharedMemory(create=True, size=total)
        else:
            self.shm = shared_memory.SharedMemory(name=name)
            # Python < 3.13 registers a segment with the resource tracker on attach as well: an independent
            # process (torchrun rank) must take that back or its tracker unlinks the segment at exit; a child
            # spawned by the creator shares the creator's tracker, where the name is registered once anyway
            if not shared_tracker:
                try:
                    from multiprocessing import resource_tracker
                    resource_tracker.unregister(self.shm._name, "shared_memory")
                except Exception:                          # noqa: BLE001
                    pass
        self.name = self.shm.name
        self.owner = create
        self.arrays = {k: np.ndarray(self.shape, dtype=dt, buffer=self.shm.buf, offset=offs[k])
                       for k, dt in self.spec.items()}

    @classmethod
    def create_or_attach(cls, spec, shape, dist=None):
        if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
            return cls(spec, shape)
        rank = dist.get_rank()
        box = [None]
        seg = None
        if rank == 0:
            seg = cls(spec, shape)
            box = [seg.name]
        dist.broadcast_object_list(box, src=0)
        if rank != 0:
            seg = cls(spec, shape, name=box[0], create=False)
        return seg

    def close(self):
        self.arrays = {}
        if self.owner:                                     # (a mapped segment survives its unlink)
            try:
                self.shm.unlink()
            except FileNotFoundError:
                pass
            self.owner = False
        try:
            self.shm.close()
        except BufferError:                                # a view is still alive somewhere: the GC unmaps it later
            pass


class ShardedStackCombiner:
    """Row-band sharded combine across the ranks of ``torch.distributed`` (one process per GPU).

    Every rank constructs one with the same arguments.  ``combine(band_frames)`` takes THIS rank's row
    band ``[r0, r1)`` of the N host frames (``band_rows()``; a rank never needs the other rows), uploads
    and reduces it, and writes the band of the result straight into a host array shared by all ranks;
    after the closing barrier ``result()`` is the full-frame dict on every rank (rank 0 is the one that
    writes the master file).  ``reduce_band`` replaces the GPU ``HostStackCombiner`` in the gloo/CPU test
    of this host logic (it gets the list of band frames and returns a dict of band arrays)."""

    def __init__(self, n, h, w, dist=None, reduce_band=None, out_f64=False, want_nrej=True, want_uncert=False,
                 want_allmasked=False, dtype=np.float32, **combine_kw):
        if dist is None:
            import torch.distributed as dist
        self.dist = dist
        self.live = dist.is_initialized()
        self.world = dist.get_world_size() if self.live else 1
        self.rank = dist.get_rank() if self.live else 0
        self.n, self.h, self.w = int(n), int(h), int(w)
        self.r0, self.r1, _, _ = row_band(self.h, self.world, self.rank)
        np_odt = np.float64 if out_f64 else np.float32
        spec = {"data": np_odt}
        if want_nrej:
            spec["nrej"] = np.uint8 if self.n <= 255 else np.uint16
        if want_uncert:
            spec["uncert"] = np_odt
        if want_allmasked:
            spec["allmasked"] = np.uint8
        self.shared = SharedHostArrays.create_or_attach(spec, (self.h, self.w), dist)
        self.bands = {k: a[self.r0:self.r1] for k, a in self.shared.arrays.items()}
        self.reduce_band = reduce_band
        self.comb = None
        self._regs = []
        if reduce_band is None and self.r1 > self.r0:
            torch = _native.require_cuda()
            self._regs = [_HostRegistration(torch, a) for a in self.bands.values()]
            self.comb = HostStackCombiner(self.n, self.r1 - self.r0, self.w, out_f64=out_f64, want_nrej=want_nrej,
                                          want_uncert=want_uncert, want_allmasked=want_allmasked, dtype=dtype,
                                          host_out=self.bands, **combine_kw)
        self.h2d_bytes = self.comb.h2d_bytes if self.comb else 0
        self.d2h_bytes = self.comb.d2h_bytes if self.comb else 0

    def band_rows(self):
        return self.r0, self.r1

    def combine(self, band_frames, barrier=True):
        if self.r1 > self.r0:
            if self.comb is not None:
                self.comb.combine(band_frames)
            else:
                res = self.reduce_band(band_frames)
                for k, dst in self.bands.items():
                    dst[...] = res[k]
        if barrier and self.live:
            self.dist.barrier()
        return self.result()

    def result(self):
        return dict(self.shared.arrays)

    def close(self):
        for r in self._regs:
            r.close()
        self._regs = []
        self.comb = None
        self.bands = {}
        if self.live:
            self.dist.barrier()                 # nobody unlinks while another rank still reads
        self.shared.close()


# ---------------------------------------------------------------------------
# files -> master: read each frame once, straight into page-locked memory, while the previous ones upload
# ---------------------------------------------------------------------------
class FileFrames:
    """The N frames of a stack as FITS files (``fitsio.ImageLayout`` each).  When every file is a BITPIX=16 /
    BZERO=32768 image (raw camera frames, reference doc/fits_metadata.md:70-76) the data units travel to the GPU
    as they are on disk (``u16_format='fits'``: 2 bytes per pixel, no host conversion); anything else is read
    through ``fitsio.read_image`` and converted to float32 on the host like the reference's ``_read_fits``."""

    def __init__(self, paths, fitsio):
        self.fitsio = fitsio
        self.paths = [str(p) for p in paths]
        self.layouts = [fitsio.ImageLayout(p) for p in self.paths]
        for p, lay in zip(self.paths, self.layouts):
            if len(lay.shape) != 2:
                raise RuntimeError(f"Error, {p} is not a 2-D image.")
            if lay.shape != self.layouts[0].shape:
                raise RuntimeError(f"Error, {p} has shape {lay.shape}, expected {self.layouts[0].shape}.")
        self.n = len(self.paths)
        self.h, self.w = self.layouts[0].shape
        self.raw_u16 = all(lay.raw_u16 for lay in self.layouts)
        self.dtype = np.dtype(np.uint16 if self.raw_u16 else np.float32)
        self.u16_format = "fits" if self.raw_u16 else "native"

    def read_band(self, i, r0, r1, out):
        """Rows ``[r0, r1)`` of frame ``i`` into ``out`` (a (r1-r0, W) array of ``self.dtype``)."""
        if self.raw_u16:
            self.layouts[i].read_rows_raw(r0, r1, out)
        else:
            data, _ = self.fitsio.read_image(self.paths[i], 0)
            out[...] = data[r0:r1]               # numpy converts to float32 (reference: data.astype(np.float32))


def combine_files(src, r0=0, r1=None, host_out=None, out_f64=False, want_nrej=True, want_uncert=False,
                  want_allmasked=False, device=None, hbm_fraction=0.6, ring=3, **params):
    """Rows ``[r0, r1)`` of the combined master of the files in ``src`` (a ``FileFrames``).

    Frame ``i+1`` is read from its file (``readinto`` page-locked staging memory) while frame ``i`` travels to
    the device cube; one ``stack_reduce`` launch follows and the result planes come back into ``host_out``
    (arrays of shape ``(r1-r0, W)``; pinned arrays of this call by default).  Stacks larger than
    ``hbm_fraction`` of the free HBM are reduced in row bands, re-reading the files once per band."""
    torch = _native.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    n, w = src.n, src.w
    r1 = src.h if r1 is None else r1
    rows = r1 - r0
    item = src.dtype.itemsize
    np_odt = np.float64 if out_f64 else np.float32
    spec = {"data": np_odt}
    if want_nrej:
        spec["nrej"] = np.uint8 if n <= 255 else np.uint16
    if want_uncert:
        spec["uncert"] = np_odt
    if want_allmasked:
        spec["allmasked"] = np.uint8
    keep = []
    if host_out is None:
        host_out = {}
        for k, dt in spec.items():
            arr, t = pinned_empty((rows, w), dt)
            host_out[k] = arr
            keep.append(t)
    free_b, _total = torch.cuda.mem_get_info(dev)
    band = int(max(1, min(rows, hbm_fraction * free_b // (n * w * item + 40 * w))))
    if band < rows and src.raw_u16:
        step = 8 // int(np.gcd(8, w))
        band = max(step, band - band % step)
    cdt = torch.float32 if src.dtype == np.float32 else torch.int16
    cube = torch.empty((n, band, w), dtype=cdt, device=dev)
    outs = {k: torch.empty((band, w), dtype=_torch_dtype(torch, dt) if np.dtype(dt) != np.uint16 else torch.uint16, device=dev)
            for k, dt in spec.items()}
    stage = [pinned_empty((band, w), src.dtype) for _ in range(ring)]
    free_ev = [None] * ring
    copy_stream = torch.cuda.Stream(device=dev)
    for b0 in range(0, rows, band):
        b1 = min(rows, b0 + band)
        nb = b1 - b0
        for i in range(n):
            slot = i % ring
            if free_ev[slot] is not None:
                free_ev[slot].synchronize()          # the upload that last used this staging buffer is done
            arr, t = stage[slot]
            src.read_band(i, r0 + b0, r0 + b1, arr[:nb])
            with torch.cuda.stream(copy_stream):
                cube[i, :nb].copy_(host_tensor(torch, arr)[:nb] if t.dtype == torch.uint16 else t[:nb], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                free_ev[slot] = ev
        torch.cuda.current_stream().wait_stream(copy_stream)
        kernels.stack_reduce(cube, row0=0, nrows=nb, out=outs, u16_format=src.u16_format, out_f64=out_f64,
                             want_nrej=want_nrej, want_uncert=want_uncert, want_allmasked=want_allmasked, **params)
        for k in spec:
            dst = host_tensor(torch, host_out[k])[b0:b1]
            s_ = outs[k][:nb]
            dst.copy_(s_.view(dst.dtype) if dst.dtype != s_.dtype else s_, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    return host_out


def _file_shard_worker(rank, world, seg_name, spec, paths, opts, devices):
    """One process per GPU of ``combine_files_sharded``: rows ``row_band(H, world, rank)`` of every file."""
    import torch
    from . import fitsio
    torch.cuda.set_device(devices[rank])
    src = FileFrames(paths, fitsio)
    seg = SharedHostArrays(spec, (src.h, src.w), name=seg_name, create=False, shared_tracker=True)
    r0, r1, _, _ = row_band(src.h, world, rank)
    if r1 > r0:
        bands = {k: a[r0:r1] for k, a in seg.arrays.items()}
        regs = [_HostRegistration(torch, a) for a in bands.values()]
        combine_files(src, r0, r1, host_out=bands, **opts)
        for reg in regs:
            reg.close()
        del bands
    seg.close()


def combine_files_sharded(paths, gpus, out_f64=False, want_nrej=True, want_uncert=False, want_allmasked=False,
                          devices=None, **params):
    """``combine_files`` over ``gpus`` GPUs of this box: one spawned process per GPU reads and reduces its row band
    of every file and writes it into a shared host array; returns the full-frame dict (copies).  ``devices``
    (tests): the CUDA device index of every rank, default ``range(gpus)``."""
    import torch
    import torch.multiprocessing as mp
    from . import fitsio
    devices = list(range(gpus)) if devices is None else list(devices)
    if len(devices) != gpus or max(devices) >= torch.cuda.device_count():
        raise RuntimeError(f"combine_files_sharded: {gpus} GPUs requested, {torch.cuda.device_count()} visible")
    lay = fitsio.ImageLayout(paths[0])
    n = len(paths)
    np_odt = np.float64 if out_f64 else np.float32
    spec = {"data": np_odt}
    if want_nrej:
        spec["nrej"] = np.uint8 if n <= 255 else np.uint16
    if want_uncert:
        spec["uncert"] = np_odt
    if want_allmasked:
        spec["allmasked"] = np.uint8
    seg = SharedHostArrays(spec, lay.shape)
    try:
        opts = dict(out_f64=out_f64, want_nrej=want_nrej, want_uncert=want_uncert, want_allmasked=want_allmasked, **params)
        mp.spawn(_file_shard_worker, args=(gpus, seg.name, {k: np.dtype(v).str for k, v in spec.items()},
                                           [str(p) for p in paths], opts, devices), nprocs=gpus, join=True)
        return {k: np.array(v, copy=True) for k, v in seg.arrays.items()}
    finally:
        seg.close()


def combine_sharded(frames, reduce_band=None, dist=None, **combine_kw):
    """Convenience wrapper: every rank holds (or can read) the same N host frames; the full-frame result
    dict (copies) is returned on rank 0 and ``None`` elsewhere."""
    n = len(frames)
    h, w = frames[0].shape
    sc = ShardedStackCombiner(n, h, w, dist=dist, reduce_band=reduce_band, dtype=frames[0].dtype, **combine_kw)
    r0, r1 = sc.band_rows()
    res = sc.combine([f[r0:r1] for f in frames])
    out = {k: np.array(v, copy=True) for k, v in res.items()} if sc.rank == 0 else None
    sc.close()
    return out
