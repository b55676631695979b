"""float64 numpy restatement of the master-frame combine (test infrastructure only).

PARITY UNPINNED.  The reference does its combine with one call into third-party
code, ``ccdproc.combine(files, method='average', sigma_clip=True,
sigma_clip_low_thresh=5, sigma_clip_high_thresh=5, sigma_clip_func=np.ma.median,
sigma_clip_dev_func=mad_std, mem_limit=5e8, unit='adu')``
(``/root/reference/AstroPhotography/scripts/ap_combine_darks.py:394-420``).
ccdproc (``requirements.txt:18`` ``ccdproc>=2.1.0``, effectively 2.4.x with
``astropy>=6.0`` ``requirements.txt:13``) and astropy are not vendored in the
reference, not installed here and not installable (no network), and the
reference has no test, fixture or golden vector for the combine
(``RELEASE_NOTES.md:167-168``).  This file therefore restates the *published*
algorithm of

* ``ccdproc.Combiner.__init__``      stack of N frames as a float64 (N,H,W) array,
* ``ccdproc.Combiner.sigma_clipping`` -> ``astropy.stats.sigma_clip(axis=0,
  sigma_lower, sigma_upper, maxiters, cenfunc, stdfunc)``: data cast to float64,
  non-finite values rejected up front, then ``while changed and it < maxiters``:
  ``c = cenfunc(kept)``, ``s = stdfunc(kept)`` (``std`` is the population
  standard deviation, ``mad_std = 1.482602218505602 * median|x - median(x)|``),
  ``lo = c - s*k_lo``, ``hi = c + s*k_hi``, reject ``x < lo`` or ``x > hi``,
* ``Combiner.average_combine``  ``nanmean`` over the kept values, mask = all N
  rejected, uncertainty = ``std(kept)/sqrt(n_kept)``,
* ``Combiner.median_combine``   ``nanmedian`` over the kept values,
  uncertainty = ``mad_std(kept)/sqrt(n_kept)``,

and is anchored on the hand-computable known-answer stacks in
``tests/golden/combine_kat.npz`` (minted by ``oracle/make_golden.py``).  If a real
ccdproc/astropy ever becomes importable on a box, ``tests/test_oracle_combine.py::
test_against_real_ccdproc`` cross-checks this restatement against it.

numpy facts relied on (checked in ``tests/test_oracle_combine.py``):
``np.sum(..., axis=0)`` over a C-contiguous (N,H,W) array accumulates the N
frames sequentially in frame order, so ``nanmean(axis=0)`` is the sequential
float64 sum of the kept values divided by their count.
"""
from __future__ import annotations

import warnings

import numpy as np

MAD_TO_STD = 1.482602218505602      # 1 / scipy.stats.norm.ppf(0.75); astropy.stats.mad_std

METHODS = ("median", "average", "min", "max")


def _nanmedian0(a):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        return np.nanmedian(a, axis=0)


def _nanmean0(a):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        return np.nanmean(a, axis=0)


def _nanstd0(a):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        return np.nanstd(a, axis=0)          # ddof=0: population std, as astropy


def _mad_std0(a):
    med = _nanmedian0(a)
    return MAD_TO_STD * _nanmedian0(np.abs(a - med))


ORDERS = ("bounds", "deviation")


def _rejected(data, c, s, k_lo, k_hi, order):
    """The clip decision in one of the two published operation orders.

    ``bounds``     astropy.stats.sigma_clip (what ccdproc >= 2.4 delegates to):
                   ``lo = c - s*k_lo; hi = c + s*k_hi; reject x < lo or x > hi``.
    ``deviation``  ccdproc <= 2.3 ``Combiner.sigma_clipping`` (and the callable-function path):
                   ``reject (x - c) < -k_lo*s or (x - c) > k_hi*s``.
    The two differ only where a sample sits within an ulp or two of a clip bound
    (``boundary_ties``); SURVEY.md section 7 item 1(b)."""
    with np.errstate(invalid="ignore"):
        if order == "bounds":
            lo = c - s * k_lo
            hi = c + s * k_hi
            return (data < lo) | (data > hi)
        if order == "deviation":
            d = data - c
            return (d < -k_lo * s) | (d > k_hi * s)
    raise ValueError(order)


def sigma_clip_stack(stack, k_lo=5.0, k_hi=5.0, maxiters=1,
                     cen="median", dev="mad_std", order="bounds"):
    """Per-pixel sigma clipping along axis 0 (astropy.stats.sigma_clip semantics).

    Returns the float64 (N,H,W) array with rejected entries set to NaN.
    ``maxiters=0`` disables clipping (only the float64 cast is applied; NaN
    inputs stay NaN, +-inf stay as they are); ``maxiters=None`` iterates to
    convergence.  ``(5, 5, 1, 'median', 'mad_std')`` is the ApMasterCal setting
    (ap_combine_darks.py:394-399,416-417).
    """
    data = np.array(stack, dtype=np.float64, copy=True)
    if maxiters == 0:
        return data
    data[~np.isfinite(data)] = np.nan       # sigma_clip masks non-finite input
    cenf = {"median": _nanmedian0, "mean": _nanmean0}[cen]
    devf = {"std": _nanstd0, "mad_std": _mad_std0}[dev]
    it = 0
    while maxiters is None or it < maxiters:
        it += 1
        c = cenf(data)
        s = devf(data)
        rej = _rejected(data, c, s, k_lo, k_hi, order)
        if not rej.any():
            break
        data[rej] = np.nan
    return data


def combine(stack, method="average", k_lo=5.0, k_hi=5.0, maxiters=1,
            cen="median", dev="mad_std", want_uncert=True, order="bounds"):
    """Restatement of ``ccdproc.combine`` on an in-memory (N,H,W) stack.

    Returns a dict: ``data`` (float64 H,W), ``nrej`` (int, number of the N
    samples not used: clipped or non-finite/NaN), ``allmasked`` (uint8, 1 where no
    sample survived), ``uncert`` (float64) and ``ncombine`` = N.
    """
    if method not in METHODS:
        raise ValueError(method)
    stack = np.asarray(stack)
    n = stack.shape[0]
    kept = sigma_clip_stack(stack, k_lo, k_hi, maxiters, cen, dev, order)
    nkept = np.sum(~np.isnan(kept), axis=0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        if method == "average":
            data = np.nanmean(kept, axis=0)
        elif method == "median":
            data = np.nanmedian(kept, axis=0)
        elif method == "min":
            data = np.nanmin(kept, axis=0)
        else:
            data = np.nanmax(kept, axis=0)
        out = {"data": data, "nrej": (n - nkept).astype(np.int64),
               "allmasked": (nkept == 0).astype(np.uint8), "ncombine": n}
        if want_uncert:
            if method == "median":
                u = _mad_std0(kept)
            else:
                u = np.nanstd(kept, axis=0)
            out["uncert"] = u / np.sqrt(nkept)
    return out


def order_census(stack, k_lo, k_hi, maxiters, cen, dev):
    """How the two operation orders relate on this stack: number of pixels, of pixels whose
    rejection count differs between the orders, of ``boundary_ties`` pixels, and whether every
    disagreement is a boundary tie (it must be)."""
    a = combine(stack, "average", k_lo, k_hi, maxiters, cen, dev, want_uncert=False, order="bounds")["nrej"]
    b = combine(stack, "average", k_lo, k_hi, maxiters, cen, dev, want_uncert=False, order="deviation")["nrej"]
    ties = boundary_ties(stack, k_lo, k_hi, maxiters, cen, dev)
    diff = a != b
    return {"pixels": int(a.size), "orders_differ": int(diff.sum()), "boundary_ties": int(ties.sum()),
            "differences_are_ties": bool(np.all(ties[diff]))}


def boundary_ties(stack, k_lo, k_hi, maxiters, cen, dev, rel=1e-12):
    """Pixels where some sample sits within ``rel`` (relative) of a clip bound
    in any iteration -- the 'documented kappa-boundary ties' at which a
    differently-ordered but equally valid evaluation may flip a rejection."""
    data = np.array(stack, dtype=np.float64, copy=True)
    tie = np.zeros(data.shape[1:], dtype=bool)
    if maxiters == 0:
        return tie
    data[~np.isfinite(data)] = np.nan
    cenf = {"median": _nanmedian0, "mean": _nanmean0}[cen]
    devf = {"std": _nanstd0, "mad_std": _mad_std0}[dev]
    it = 0
    while maxiters is None or it < maxiters:
        it += 1
        c = cenf(data)
        s = devf(data)
        lo = c - s * k_lo
        hi = c + s * k_hi
        scale = np.maximum(np.abs(lo), np.abs(hi)) * rel
        with np.errstate(invalid="ignore"):
            near = (np.abs(data - lo) <= scale) | (np.abs(data - hi) <= scale)
            tie |= np.any(near, axis=0)
            rej = (data < lo) | (data > hi)
        if not rej.any():
            break
        data[rej] = np.nan
    return tie


def sigma_clipped_stats_global(data, sigma=3.0, maxiters=5):
    """Restatement of ``astropy.stats.sigma_clipped_stats(data, sigma=...)`` with
    its defaults (cenfunc='median', stdfunc='std', maxiters=5) over the whole
    array -- the call at ``core/ApFindBadPixels.py:191``.  Returns
    ``(mean, median, std)`` of the surviving values (float64).
    """
    x = np.asarray(data, dtype=np.float64).ravel()
    x = x[np.isfinite(x)]
    it = 0
    while maxiters is None or it < maxiters:
        it += 1
        c = np.median(x)
        s = np.std(x)
        lo = c - s * sigma
        hi = c + s * sigma
        keep = (x >= lo) & (x <= hi)
        if keep.all():
            break
        x = x[keep]
    # numpy float64 scalars (not Python floats), as astropy returns: this keeps
    # the later ``data < lothresh`` comparison in float64 under NEP 50.
    return np.float64(np.mean(x)), np.float64(np.median(x)), np.float64(np.std(x))
