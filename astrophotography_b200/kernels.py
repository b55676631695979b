"""Array-level entry points: torch CUDA tensors in, torch CUDA tensors out.

Thin, validation-only wrappers over the C ABI (``include/apgpu.h``).  torch is
used for device memory and streams only; every arithmetic operation of the hot
path happens inside ``libapgpu.so``.  All calls are asynchronous on torch's
current CUDA stream.
"""
from __future__ import annotations

import ctypes

from . import _native

METHODS = {"median": 0, "average": 1, "mean": 1, "min": 2, "max": 3}
CENFUNCS = {"mean": 0, "median": 1}
DEVFUNCS = {"std": 0, "mad_std": 1}
_FORCE_GENERIC = 1
U16_FORMATS = {"native": 0, "fits": 1}        # include/apgpu.h: APGPU_U16_NATIVE / APGPU_U16_FITS_BZERO
_PREFER = {None: 0, "registers": 2, "shared": 4, "registers_tma": 2 | 8, "tma": 8,
           "registers_direct": 2 | 16, "direct": 16, "registers_cpasync": 2 | 32, "cpasync": 32,
           "registers_tensormap": 2 | 64, "tensormap": 64}


def _stream(torch):
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _check_image(torch, t, name, dtype=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a C-contiguous tensor")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")


def stack_kernel_name(n, method="average", k_lo=5.0, k_hi=5.0, maxiters=1, cen="median",
                      dev="mad_std", want_uncert=False, out_f64=False, force_generic=False, prefer=None):
    lib = _native.load()
    return lib.apgpu_stack_kernel_name(
        int(n), METHODS[method], float(k_lo), float(k_hi), _maxiters(maxiters), CENFUNCS[cen],
        DEVFUNCS[dev], int(want_uncert), int(out_f64),
        (_FORCE_GENERIC if force_generic else 0) | _PREFER[prefer]).decode()


def stack_last_staging():
    """-1 generic kernel, 0 direct loads, 1 CTA-wide bulk / tensor-map copies, 2 warp cp.async, 3 warp tensor-map TMA,
    4 lane-split cp.async, 5 warp-cooperative swizzled tensor-map TMA (see include/apgpu.h)."""
    return int(_native.load().apgpu_stack_last_staging())


def _maxiters(maxiters):
    return -1 if maxiters is None else int(maxiters)


def stack_reduce(frames, method="average", k_lo=5.0, k_hi=5.0, maxiters=1, cen="median",
                 dev="mad_std", row0=0, nrows=None, out_f64=False, want_nrej=True,
                 want_uncert=False, want_allmasked=False, force_generic=False, out=None, prefer=None,
                 u16_format="native"):
    """Per-pixel combine of N frames (``apgpu_stack_reduce_f32`` / ``apgpu_stack_reduce_u16``).

    ``frames``: a (N,H,W) CUDA tensor or a sequence of N (H,W) CUDA tensors, float32 or
    uint16 (int16 tensors are taken as the raw 16-bit samples).  uint16 frames are
    converted inside the load phase of the kernels; ``u16_format="fits"`` means the
    samples are the data unit of a BITPIX=16 / BZERO=32768 FITS file as stored on disk
    (big-endian int16 + 32768).  Defaults are the reference's ApMasterCal settings
    (scripts/ap_combine_darks.py:394-399): average after one 5-sigma median/MAD clip.
    Returns a dict of (H,W) tensors: ``data``, and when requested ``nrej`` (uint8, or
    uint16 for N>255), ``uncert``, ``allmasked``.  Only rows ``[row0, row0+nrows)`` are
    written.
    """
    torch = _native.require_cuda()
    lib = _native.load()
    sample_dtypes = (torch.float32, torch.uint16, torch.int16)
    if isinstance(frames, torch.Tensor):
        if frames.dim() != 3:
            raise RuntimeError("stack_reduce: frame cube must be (N, H, W)")
        _check_image(torch, frames, "frames")
        flist = [frames[i] for i in range(frames.shape[0])]
    else:
        flist = list(frames)
        for i, f in enumerate(flist):
            _check_image(torch, f, f"frames[{i}]")
    n = len(flist)
    if n == 0:
        raise RuntimeError("stack_reduce: no frames")
    dt = flist[0].dtype
    if dt not in sample_dtypes or any(f.dtype != dt for f in flist):
        raise RuntimeError(f"stack_reduce: frames must all be float32 or all be uint16, got {dt}")
    h, w = flist[0].shape
    for f in flist:
        if tuple(f.shape) != (h, w):
            raise RuntimeError("stack_reduce: frames differ in shape")
    if method not in METHODS or cen not in CENFUNCS or dev not in DEVFUNCS:
        raise RuntimeError(f"stack_reduce: bad method/cen/dev {method}/{cen}/{dev}")
    if u16_format not in U16_FORMATS:
        raise RuntimeError(f"stack_reduce: bad u16_format {u16_format}")
    if nrows is None:
        nrows = h - row0
    device = flist[0].device
    res = out if out is not None else {}
    if "data" not in res:
        res["data"] = torch.empty((h, w), dtype=torch.float64 if out_f64 else torch.float32, device=device)
    if want_nrej and "nrej" not in res:
        res["nrej"] = torch.empty((h, w), dtype=torch.uint8 if n <= 255 else torch.uint16, device=device)
    if want_uncert and "uncert" not in res:
        res["uncert"] = torch.empty_like(res["data"])
    if want_allmasked and "allmasked" not in res:
        res["allmasked"] = torch.empty((h, w), dtype=torch.uint8, device=device)
    ptrs = (ctypes.c_void_p * n)(*[f.data_ptr() for f in flist])
    nrej = res.get("nrej") if want_nrej else None
    flags = (_FORCE_GENERIC if force_generic else 0) | _PREFER[prefer]
    tail = (int(row0), int(nrows), METHODS[method], float(k_lo), float(k_hi),
            _maxiters(maxiters), CENFUNCS[cen], DEVFUNCS[dev],
            _ptr(res["data"]), int(res["data"].dtype == torch.float64),
            _ptr(nrej), int(nrej is not None and nrej.dtype == torch.uint16),
            _ptr(res.get("uncert") if want_uncert else None),
            _ptr(res.get("allmasked") if want_allmasked else None),
            flags, _stream(torch))
    if dt == torch.float32:
        st = lib.apgpu_stack_reduce_f32(ptrs, n, h, w, *tail)
        _native.check(st, "apgpu_stack_reduce_f32")
    else:
        st = lib.apgpu_stack_reduce_u16(ptrs, U16_FORMATS[u16_format], n, h, w, *tail)
        _native.check(st, "apgpu_stack_reduce_u16")
    return res


def flat_norm(flat):
    """``np.nanmean(flat)`` as a 1-element float32 CUDA tensor (bit-exact numpy tree)."""
    torch = _native.require_cuda()
    lib = _native.load()
    _check_image(torch, flat, "flat", torch.float32)
    npix = flat.numel()
    nbytes = int(lib.apgpu_flat_norm_workspace_bytes(npix))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=flat.device)
    norm = torch.empty(1, dtype=torch.float32, device=flat.device)
    _native.check(lib.apgpu_flat_norm_f32(_ptr(flat), npix, _ptr(ws), nbytes, _ptr(norm), _stream(torch)),
                  "apgpu_flat_norm_f32")
    return norm


def flat_normalise(flat, norm=None):
    """``flat / np.nanmean(flat)`` (ApCalibrate._generate_flat, core/ApCalibrate.py:178-190)."""
    torch = _native.require_cuda()
    lib = _native.load()
    if norm is None:
        norm = flat_norm(flat)
    out = torch.empty_like(flat)
    _native.check(lib.apgpu_flat_divide_f32(_ptr(flat), _ptr(norm), _ptr(out), flat.numel(), _stream(torch)),
                  "apgpu_flat_divide_f32")
    return out, norm


def calibrate(raw, bias, dark, normflat=None, exp_ratio=1.0, dark_still_biased=False,
              pedestal=None, out=None):
    """Fused bias / scaled-dark / flat calibration (core/ApCalibrate.py:439-474).

    ``raw`` may be float32 (already converted, pedestal applied) or uint16
    (conversion and optional ``pedestal`` fused into the kernel)."""
    torch = _native.require_cuda()
    lib = _native.load()
    _check_image(torch, raw, "raw")
    _check_image(torch, bias, "bias", torch.float32)
    _check_image(torch, dark, "dark", torch.float32)
    if normflat is not None:
        _check_image(torch, normflat, "normflat", torch.float32)
    for name, t in (("bias", bias), ("dark", dark), ("normflat", normflat)):
        if t is not None and tuple(t.shape) != tuple(raw.shape):
            raise RuntimeError(f"calibrate: {name} shape {tuple(t.shape)} != raw shape {tuple(raw.shape)}")
    if out is None:
        out = torch.empty(raw.shape, dtype=torch.float32, device=raw.device)
    npix = raw.numel()
    if raw.dtype == torch.float32:
        if pedestal:
            raise RuntimeError("calibrate: pedestal fusion is for uint16 raw frames only")
        st = lib.apgpu_calibrate_f32(_ptr(raw), _ptr(bias), _ptr(dark), _ptr(normflat),
                                     float(exp_ratio), int(bool(dark_still_biased)), _ptr(out), npix,
                                     _stream(torch))
    elif raw.dtype == torch.uint16:
        has_ped = pedestal is not None and float(pedestal) != 0.0
        st = lib.apgpu_calibrate_u16(_ptr(raw), float(pedestal or 0.0), int(has_ped), _ptr(bias),
                                     _ptr(dark), _ptr(normflat), float(exp_ratio),
                                     int(bool(dark_still_biased)), _ptr(out), npix, _stream(torch))
    else:
        raise RuntimeError(f"calibrate: raw dtype {raw.dtype} not supported (float32 or uint16)")
    _native.check(st, "apgpu_calibrate")
    return out


RAW_KINDS = {"f32": 0, "u16": 1, "u16_fits": 2}       # include/apgpu.h: APGPU_RAW_*


def calibrate_repair(raw, bias, dark, normflat=None, exp_ratio=1.0, dark_still_biased=False, pedestal=None,
                     mask=None, deltapix=2, min_valid=4, raw_kind=None, out=None, out_big_endian=False, counts=None):
    """Calibration and bad-pixel repair of one frame in ONE launch (``apgpu_calibrate_repair``): what
    ``calibrate`` followed by ``fix_badpix`` gives, bit for bit, without the intermediate image.

    ``raw``: float32, uint16, or int16/uint16 holding the raw FITS data unit (``raw_kind="u16_fits"``).
    ``mask``: uint8 CUDA image (non-zero = bad) or None.  Returns ``(out, counts)``; ``out_big_endian``
    writes the float32 result byte-swapped, ready to be the data unit of a BITPIX=-32 FITS file."""
    torch = _native.require_cuda()
    lib = _native.load()
    _check_image(torch, raw, "raw")
    _check_image(torch, bias, "bias", torch.float32)
    _check_image(torch, dark, "dark", torch.float32)
    if normflat is not None:
        _check_image(torch, normflat, "normflat", torch.float32)
    if mask is not None:
        _check_image(torch, mask, "mask", torch.uint8)
    for name, t in (("bias", bias), ("dark", dark), ("normflat", normflat), ("mask", mask)):
        if t is not None and tuple(t.shape) != tuple(raw.shape):
            raise RuntimeError(f"calibrate_repair: {name} shape {tuple(t.shape)} != raw shape {tuple(raw.shape)}")
    if raw_kind is None:
        raw_kind = "f32" if raw.dtype == torch.float32 else "u16"
    if raw_kind not in RAW_KINDS or (raw_kind == "f32") != (raw.dtype == torch.float32) or \
            (raw_kind != "f32" and raw.dtype not in (torch.uint16, torch.int16)):
        raise RuntimeError(f"calibrate_repair: raw dtype {raw.dtype} does not fit raw_kind {raw_kind}")
    h, w = raw.shape
    if out is None:
        out = torch.empty((h, w), dtype=torch.float32, device=raw.device)
    if mask is not None and counts is None:
        counts = torch.zeros(2, dtype=torch.int64, device=raw.device)
    has_ped = pedestal is not None and float(pedestal) != 0.0
    st = lib.apgpu_calibrate_repair(_ptr(raw), RAW_KINDS[raw_kind], float(pedestal or 0.0), int(has_ped),
                                    _ptr(bias), _ptr(dark), _ptr(normflat), float(exp_ratio),
                                    int(bool(dark_still_biased)), _ptr(mask), int(h), int(w), int(deltapix),
                                    int(min_valid), _ptr(out), int(bool(out_big_endian)), _ptr(counts), _stream(torch))
    _native.check(st, "apgpu_calibrate_repair")
    return out, counts


IMARITH_OPS = {"ADD": 0, "SUB": 1, "MUL": 2, "DIV": 3}


def imarith(a, op, b, out=None):
    """``np.<op>(a, b, out=float32)`` for a float32 CUDA image ``a`` and ``b`` a Python number, a float32 or a
    float64 CUDA image (core/ApImArith.py:321-333)."""
    torch = _native.require_cuda()
    lib = _native.load()
    _check_image(torch, a, "a", torch.float32)
    if op not in IMARITH_OPS:
        raise RuntimeError(f"imarith: bad operation {op}")
    if out is None:
        out = torch.empty_like(a)
    if isinstance(b, torch.Tensor):
        _check_image(torch, b, "b")
        if tuple(b.shape) != tuple(a.shape) or b.dtype not in (torch.float32, torch.float64):
            raise RuntimeError(f"imarith: second image {tuple(b.shape)} {b.dtype} does not fit {tuple(a.shape)} float32")
        st = lib.apgpu_imarith_f32(_ptr(a), _ptr(b), 1 if b.dtype == torch.float32 else 2, 0.0, IMARITH_OPS[op],
                                   _ptr(out), a.numel(), _stream(torch))
    else:
        st = lib.apgpu_imarith_f32(_ptr(a), ctypes.c_void_p(0), 0, float(b), IMARITH_OPS[op], _ptr(out), a.numel(),
                                   _stream(torch))
    _native.check(st, "apgpu_imarith_f32")
    return out


_MASK_DTYPES = None


def _mask_code(torch, dtype):
    global _MASK_DTYPES
    if _MASK_DTYPES is None:
        _MASK_DTYPES = {torch.uint8: 0, torch.bool: 0, torch.int16: 1, torch.uint16: 1,
                        torch.int32: 2, torch.float32: 3, torch.float64: 4}
    if dtype not in _MASK_DTYPES:
        raise RuntimeError(f"fix_badpix: mask dtype {dtype} not supported")
    return _MASK_DTYPES[dtype]


def fix_badpix(data, mask, deltapix=1, min_valid=4, image_rows=None, band_row0=0,
               row0=None, nrows=None, counts=None):
    """Median repair of masked pixels (core/ApFixBadPixels.py:380-419).

    ``data``/``mask`` hold rows ``[band_row0, band_row0+data.shape[0])`` of an
    image with ``image_rows`` rows (defaults: the whole image).  Returns
    ``(out, counts)`` where ``out`` has ``nrows`` rows starting at image row
    ``row0`` and ``counts`` is a 2-element int64 CUDA tensor
    ``[n_bad, n_fixed]`` (accumulated into when passed in)."""
    torch = _native.require_cuda()
    lib = _native.load()
    _check_image(torch, data, "data", torch.float32)
    _check_image(torch, mask, "mask")
    if tuple(data.shape) != tuple(mask.shape):
        raise RuntimeError(f"Error, the shape of the input data array ({tuple(data.shape)}) does not "
                           f"match that of the bad pixel mask array ({tuple(mask.shape)}).")
    band_rows, w = data.shape
    if image_rows is None:
        image_rows = band_row0 + band_rows
    if row0 is None:
        row0 = band_row0
    if nrows is None:
        nrows = band_row0 + band_rows - row0
    out = torch.empty((nrows, w), dtype=torch.float32, device=data.device)
    if counts is None:
        counts = torch.zeros(2, dtype=torch.int64, device=data.device)
    st = lib.apgpu_fix_badpix_f32(_ptr(data), _ptr(mask), _mask_code(torch, mask.dtype),
                                  int(image_rows), int(w), int(band_row0), int(band_rows),
                                  int(row0), int(nrows), int(deltapix), int(min_valid),
                                  _ptr(out), _ptr(counts), _stream(torch))
    _native.check(st, "apgpu_fix_badpix_f32")
    return out, counts


def sigma_clipped_stats(data, sigma=3.0, maxiters=5):
    """``astropy.stats.sigma_clipped_stats(data, sigma=sigma)`` over a float32 CUDA image.

    Returns ``(mean, median, std, count)`` of the surviving pixels as Python
    numbers (one small device-to-host read)."""
    torch = _native.require_cuda()
    lib = _native.load()
    _check_image(torch, data, "data", torch.float32)
    npix = data.numel()
    nbytes = int(lib.apgpu_image_stats_workspace_bytes(npix))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=data.device)
    out4 = torch.empty(4, dtype=torch.float64, device=data.device)
    _native.check(lib.apgpu_sigma_clipped_stats_f32(_ptr(data), npix, float(sigma), int(maxiters), _ptr(ws),
                                                    nbytes, _ptr(out4), _stream(torch)),
                  "apgpu_sigma_clipped_stats_f32")
    mean, med, std, cnt = out4.cpu().tolist()
    return mean, med, std, int(cnt)


def threshold_mask(data, lo, hi):
    """uint8 mask ``(data < lo) | (data > hi)`` and the number of set pixels."""
    torch = _native.require_cuda()
    lib = _native.load()
    _check_image(torch, data, "data", torch.float32)
    mask = torch.empty(data.shape, dtype=torch.uint8, device=data.device)
    nbad = torch.zeros(1, dtype=torch.int64, device=data.device)
    _native.check(lib.apgpu_threshold_mask_f32(_ptr(data), data.numel(), float(lo), float(hi), _ptr(mask),
                                               _ptr(nbad), _stream(torch)), "apgpu_threshold_mask_f32")
    return mask, int(nbad.item())
