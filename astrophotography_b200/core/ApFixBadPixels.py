"""ApFixBadPixels: median repair of pre-identified bad pixels, on the GPU.

Host-side mirror of the reference class ``AstroPhotography/core/ApFixBadPixels.py``
(ctor :28-53, ``fix_files`` :245-290, ``fix_bad_pixels`` :292-445): same
constructor, method names, argument meaning, returned ``(newdata, fixed_stats)``
and the same ``BPIX*`` keyword dictionary.  The per-bad-pixel Python loop
(:380-419) is replaced by one launch of ``apgpu_fix_badpix_f32``.
"""
from __future__ import annotations

import time
from pathlib import Path

import numpy as np

from .. import _native, kernels
from ._base import ApBase


class ApFixBadPixels(ApBase):
    """Fixes pre-identified bad pixels by replacing them with the median of the
    surrounding good pixels (window half-size ``deltapix``)."""

    MASK_GOOD = 0
    _name = "ApFixBadPixels"

    def __init__(self, loglevel):
        self._loglevel = loglevel
        self._initialize_logger(loglevel)
        self._min_valid = 4                    # reference :45
        self._replace_unfixable = False        # declared by the reference (:49-50), never used
        self._replace_unfixable_value = np.nan
        self._logger.debug(f"{self._name} instance constructed.")

    # -- file level ---------------------------------------------------------
    def fix_files(self, inpdata_file, badpixmask_file, outdata_file, deltapix=1):
        ext_num = 0
        deltapix = int(deltapix)
        self._logger.info(f"fix_files input data file={inpdata_file}, mask file={badpixmask_file},"
                          f" output file={outdata_file}, deltapix={deltapix}")
        # the reference's ApFixBadPixels._read_fits (:92-153) keeps the file's dtype
        idata, ihdr, ped = self._read_fits(inpdata_file, ext_num, to_float=False)
        if ped != 0:
            # (the reference's in-place ``ext_data += pedestal`` raises for integer images, :147;
            #  here an integer image with a PEDESTAL is promoted to float32 first)
            if not np.issubdtype(idata.dtype, np.floating):
                idata = idata.astype(np.float32)
            idata = idata + idata.dtype.type(ped)
        mskdata, _, _ = self._read_fits(badpixmask_file, ext_num, to_float=False)
        odata, odict = self.fix_bad_pixels(idata, mskdata, deltapix)
        odict["BPIXFILE"] = (Path(badpixmask_file).name, "Name of master bad pixel file used")
        self._write_image_like(inpdata_file, ext_num, outdata_file, odata, odict,
                               f"Applied {self._name}", drop_scaling=False, only_bpix=True)
        self._logger.info(f"Wrote bad pixel corrected file to {outdata_file}")

    def _stats_dict(self, npix, nbad, nfixed, deltapix):
        return _stats_dict_impl(self._min_valid, npix, nbad, nfixed, deltapix)

    # -- array level --------------------------------------------------------
    def fix_bad_pixels(self, data, badpixmask, deltapix=1):
        """Return ``(newdata, fixed_stats)``.

        ``data`` / ``badpixmask`` may be numpy arrays (result is a numpy array of
        the same dtype) or CUDA tensors (float32; result stays on the device)."""
        deltapix = int(deltapix)
        torch = _native.require_cuda()
        on_device = isinstance(data, torch.Tensor)
        shape, mshape = tuple(data.shape), tuple(badpixmask.shape)
        self._logger.info(f"fix_bad_pixels: data has {shape[0]} rows x {shape[1]} columns, dtype={data.dtype}")
        self._logger.info(f"fix_bad_pixels: mask has {mshape[0]} rows x {mshape[1]} columns, dtype={badpixmask.dtype}")
        if shape != mshape:
            msg = (f"Error, the shape of the input data array ({shape})"
                   f" does not match that of the bad pixel mask array ({mshape}).")
            self._logger.error(msg)
            raise RuntimeError(msg)
        t0 = time.perf_counter()
        if on_device:
            d_dev, m_dev, orig_dtype = data, _mask_to_device(torch, badpixmask, data.device), None
        else:
            data = np.asarray(data)
            orig_dtype = data.dtype
            if not np.issubdtype(orig_dtype, np.floating):
                self._logger.warning("Pixel medians may suffer from casting truncation because"
                                     f" the input data is not a floating point datatype ({orig_dtype}).")
            work = data if orig_dtype == np.float32 else data.astype(np.float32)
            if not _exact_in_f32(data):
                # np.median's even-count mean (a + b) / 2 is a float64 operation for float64 / 32- and 64-bit
                # integer data: float32 arithmetic could round differently, so those dtypes are refused
                # rather than silently repaired with other last bits
                msg = (f"fix_bad_pixels: dtype {orig_dtype} is not supported on the GPU path: only float32 and"
                       " 8/16-bit integer data are repaired bit for bit like the reference (cast the image to"
                       " float32 first if float32 medians are acceptable).")
                self._logger.error(msg)
                raise RuntimeError(msg)
            d_dev = torch.from_numpy(np.ascontiguousarray(work)).cuda()
            m_dev = _mask_to_device(torch, badpixmask, d_dev.device)
        out, counts = kernels.fix_badpix(d_dev.contiguous(), m_dev, deltapix, self._min_valid)
        nbad, nfixed = (int(v) for v in counts.cpu().numpy())
        run_time = time.perf_counter() - t0
        npix = shape[0] * shape[1]
        nnotfix = nbad - nfixed
        pctbad = 100.0 * nbad / npix
        self._logger.debug(f"Percentage of pixels considered bad: {pctbad:.3f} ({nbad:d}/{npix:d})")
        ms_per_pix = 1000 * run_time / nbad if nbad else float("inf")
        self._logger.info(f"Processed {nbad} pixels in {run_time:.3f} s, {ms_per_pix:.4f} ms per bad pixel.")
        if nnotfix > 0:
            self._logger.warning(f"Could not fix {nnotfix} pixels as they had less"
                                 f" than {self._min_valid} good neighbors when deltapix={deltapix} pixels.")
        fixed_stats = self._stats_dict(npix, nbad, nfixed, deltapix)
        if on_device:
            return out, fixed_stats
        newdata = out.cpu().numpy()
        if orig_dtype != np.float32:
            # numpy assigns the float median into an integer array by C truncation (:409)
            newdata = np.trunc(newdata).astype(orig_dtype) if not np.issubdtype(orig_dtype, np.floating) \
                else newdata.astype(orig_dtype)
        return newdata, fixed_stats


def _stats_dict_impl(min_valid, npix, nbad, nfixed, deltapix):
    """The ``fixed_stats`` dictionary of the reference (core/ApFixBadPixels.py:430-443)."""
    return {
        "numpix": (npix, "Total number of pixels in image"),
        "BPIXNBAD": (nbad, "Total number of bad pixels in bad pixel file"),
        "pctbad": (100.0 * nbad / npix, "Percentage of pixel defined bad"),
        "BPIX_MIN": (min_valid, "Minimum number of good neighors needed"),
        "BPIXDPIX": (deltapix, "Half height/width of collection region (pixels)"),
        "BPIXNREM": (nbad - nfixed, "Number of bad pixels not corrected"),
        "BPIXCORR": (nfixed > 0, "True if any bad pixels were corrected"),
        "BPIXNFIX": (nfixed, "Number of bad pixels corrected"),
    }


def _exact_in_f32(a):
    """float32 data, and integers whose pairwise sums are exact in float32 (8 / 16 bits)."""
    return a.dtype in (np.float32, np.uint8, np.int8, np.uint16, np.int16, np.bool_)


def _mask_to_device(torch, mask, device):
    if isinstance(mask, torch.Tensor):
        return mask.to(device).contiguous()
    mask = np.ascontiguousarray(mask)
    if mask.dtype == np.uint16:
        return torch.from_numpy(mask.view(np.int16)).to(device)      # only (!= 0) matters
    if mask.dtype in (np.int8,):
        return torch.from_numpy(mask.view(np.uint8)).to(device)
    if mask.dtype in (np.int64, np.uint32, np.uint64, np.float16):
        mask = (mask != 0).astype(np.uint8)
    return torch.from_numpy(mask).to(device)
