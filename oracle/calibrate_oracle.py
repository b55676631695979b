"""numpy restatement of the science-frame calibration (test infrastructure only).

Follows ``/root/reference/AstroPhotography/core/ApCalibrate.py``:

* ``read_convert``   -- ``_read_fits`` :303-326: non-float data -> float32, then
  ``+= PEDESTAL`` when the keyword is present and non-zero.
* ``normalise_flat`` -- ``_generate_flat`` :178-190: ``flat / np.nanmean(flat)``.
* ``calibrate``      -- ``calibrate`` :439-474: bias subtraction, optional
  dark-bias subtraction, exposure-ratio dark scaling, guarded flat division.
* ``exptime_ratio``  -- ``_find_exptime_ratio`` :128-164.

PINNED: ``oracle/make_golden.py`` runs the reference source verbatim
(``oracle/ref_exec.py``) on seeded inputs and stores its outputs in
``tests/golden/calibrate_*.npz``; ``tests/test_oracle_calibrate.py`` checks this
restatement bit-for-bit against those fixtures (and against the live reference
when ``/root/reference`` is present).

The operations are written exactly as the reference writes them so that
numpy's dtype promotion (NEP 50: the Python-float exposure ratio is 'weak'
and adopts the array dtype) and per-operation rounding are reproduced.
"""
from __future__ import annotations

import warnings

import numpy as np


def read_convert(data, pedestal=None):
    """ApCalibrate._read_fits :303-326 (dtype conversion and pedestal)."""
    data = np.asarray(data)
    if not np.issubdtype(data.dtype, np.floating):
        data = data.astype(np.float32)
    else:
        data = data.copy()
    if pedestal is not None:
        pedestal = float(pedestal)
        if pedestal != 0:
            data += pedestal
    return data


def exptime_ratio(img_hdr, dark_hdr):
    """ApCalibrate._find_exptime_ratio :128-164."""
    img_exp = dark_exp = None
    for kw in ("EXPOSURE", "EXPTIME"):
        if img_exp is None and kw in img_hdr:
            img_exp = float(img_hdr[kw])
        if dark_exp is None and kw in dark_hdr:
            dark_exp = float(dark_hdr[kw])
    if img_exp is None and dark_exp is None:
        raise RuntimeError("Could not determine exposure time for both image and dark.")
    if img_exp is None:
        raise RuntimeError("Could not determine exposure time for image (dark exposure found).")
    if dark_exp is None:
        raise RuntimeError("Could not determine exposure time for dark (img exposure found).")
    return img_exp / dark_exp


def flat_norm_factor(flat):
    """``np.nanmean(flat_data)`` (ApCalibrate.py:181)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        return np.nanmean(flat)


def normalise_flat(flat):
    """ApCalibrate._generate_flat :178-190 (MEAN_FULL)."""
    with np.errstate(all="ignore"):
        return flat / flat_norm_factor(flat)


def calibrate(raw, bias, dark, exp_ratio, norm_flat=None, dark_still_biased=False):
    """ApCalibrate.calibrate :439-474 on arrays already through ``read_convert``.

    ``exp_ratio`` must be a Python float (as the reference computes it,
    :161) so that it is a weak scalar in ``exp_ratio * dark``.
    """
    exp_ratio = float(exp_ratio)
    with np.errstate(all="ignore"):
        img_sub_b = raw - bias
        dark_sub_b = dark - bias if dark_still_biased else dark
        dark_scaled = exp_ratio * dark_sub_b
        img_sub_bd = img_sub_b - dark_scaled
        if norm_flat is not None:
            return np.where(norm_flat != 0, img_sub_bd / norm_flat, img_sub_bd)
        return img_sub_bd


def numpy_pairwise_sum_f32(a):
    """Pure restatement of numpy's float32 pairwise summation of a contiguous
    1-D array (``FLOAT_pairwise_sum``, numpy/_core/src/umath/loops_utils.h.src,
    PW_BLOCKSIZE=128, 8 accumulators), which is what ``np.sum`` /
    ``np.nanmean`` run on a C-contiguous float32 image.  Used to pin the
    summation tree the CUDA flat-norm kernel reproduces.
    """
    a = np.ascontiguousarray(a, dtype=np.float32).ravel()

    def pw(lo, n):
        if n < 8:
            res = np.float32(0.0)
            for i in range(n):
                res = np.float32(res + a[lo + i])
            return res
        if n <= 128:
            r = a[lo:lo + 8].copy()
            m = n - (n % 8)
            blk = a[lo + 8:lo + m].reshape(-1, 8)
            for row in blk:
                r = (r + row).astype(np.float32)
            res = np.float32(np.float32(np.float32(r[0] + r[1]) + np.float32(r[2] + r[3]))
                             + np.float32(np.float32(r[4] + r[5]) + np.float32(r[6] + r[7])))
            for i in range(m, n):
                res = np.float32(res + a[lo + i])
            return res
        n2 = n // 2
        n2 -= n2 % 8
        return np.float32(pw(lo, n2) + pw(lo + n2, n - n2))

    return np.float32(np.float32(0.0) + pw(0, a.size))
