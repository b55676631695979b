"""numpy restatement of bad-pixel repair and mask rules (test infrastructure only).

* ``fix_bad_pixels_loop`` -- per-bad-pixel restatement of
  ``/root/reference/AstroPhotography/core/ApFixBadPixels.py:334-443``
  (window clipped to the image :383-386, donors taken from the ORIGINAL data
  and mask :391-392, ``ngood >= 4`` :397, ``np.median`` of the good donors :407).
  Small inputs only (pure-Python loop, like the reference).
* ``fix_bad_pixels_vec``  -- the same function vectorised over all bad pixels
  (gather (2dp+1)^2 windows, sort with invalid donors pushed to +inf) so that
  full-size frames finish in seconds; checked against the loop form and the
  verbatim reference in ``tests/test_oracle_badpix.py``.
* ``user_mask`` / ``auto_mask`` -- ``core/ApFindBadPixels.py:70-158`` (1-based
  inclusive YAML columns / rows / rectangles, out-of-range entries skipped,
  overlaps accumulate ``+= 2``) and ``:191-209`` (``AUTO_BAD=1`` outside
  ``median -+ sigma*std`` of the sigma-clipped dark).

PINNED: ``tests/golden/badpix_*.npz`` hold outputs of the reference source run
verbatim (``oracle/ref_exec.py``, minted by ``oracle/make_golden.py``).

np.median semantics reproduced: even donor count -> ``float32(a+b)/2`` (the
mean of the two middle float32 values is taken in float32); any NaN donor ->
NaN.
"""
from __future__ import annotations

import numpy as np

MIN_VALID = 4       # ApFixBadPixels.py:45
GOOD, AUTO_BAD, USER_BAD = 0, 1, 2   # ApFindBadPixels.py:26-28


def _stats(npix, nbad, nrem, deltapix, min_valid):
    nfixed = nbad - nrem
    pct = 100.0 * nbad / npix
    return {
        "numpix": (npix, "Total number of pixels in image"),
        "BPIXNBAD": (nbad, "Total number of bad pixels in bad pixel file"),
        "pctbad": (pct, "Percentage of pixel defined bad"),
        "BPIX_MIN": (min_valid, "Minimum number of good neighors needed"),
        "BPIXDPIX": (deltapix, "Half height/width of collection region (pixels)"),
        "BPIXNREM": (nrem, "Number of bad pixels not corrected"),
        "BPIXCORR": (nfixed > 0, "True if any bad pixels were corrected"),
        "BPIXNFIX": (nfixed, "Number of bad pixels corrected"),
    }


def fix_bad_pixels_loop(data, badpixmask, deltapix=1, min_valid=MIN_VALID):
    deltapix = int(deltapix)
    if data.shape != badpixmask.shape:
        raise RuntimeError("shape mismatch")
    newdata = data.copy()
    mask = badpixmask != GOOD
    nrows, ncols = mask.shape
    nrem = 0
    rr, cc = np.nonzero(mask)           # C order, like row_idxs[mask] (:378-379)
    with np.errstate(all="ignore"):
        for r, c in zip(rr, cc):
            rmin, rmax = max(0, r - deltapix), min(nrows, r + deltapix + 1)
            cmin, cmax = max(0, c - deltapix), min(ncols, c + deltapix + 1)
            d = data[rmin:rmax, cmin:cmax]
            m = mask[rmin:rmax, cmin:cmax]
            if m.size - m.sum() >= min_valid:
                newdata[r, c] = np.median(d[~m])
            else:
                nrem += 1
    return newdata, _stats(data.size, int(mask.sum()), nrem, deltapix, min_valid)


def fix_bad_pixels_vec(data, badpixmask, deltapix=1, min_valid=MIN_VALID):
    """Vectorised equivalent of ``fix_bad_pixels_loop`` (float32 or float64 data)."""
    deltapix = int(deltapix)
    if data.shape != badpixmask.shape:
        raise RuntimeError("shape mismatch")
    newdata = data.copy()
    mask = badpixmask != GOOD
    nrows, ncols = mask.shape
    rr, cc = np.nonzero(mask)
    nbad = rr.size
    if nbad == 0:
        return newdata, _stats(data.size, 0, 0, deltapix, min_valid)
    offs = np.arange(-deltapix, deltapix + 1)
    dr, dc = np.meshgrid(offs, offs, indexing="ij")
    wr = rr[:, None] + dr.ravel()[None, :]
    wc = cc[:, None] + dc.ravel()[None, :]
    inside = (wr >= 0) & (wr < nrows) & (wc >= 0) & (wc < ncols)
    wrc = np.clip(wr, 0, nrows - 1)
    wcc = np.clip(wc, 0, ncols - 1)
    good = inside & ~mask[wrc, wcc]
    vals = data[wrc, wcc]
    ngood = good.sum(axis=1)
    anynan = np.any(good & np.isnan(vals), axis=1)
    # Non-donors sort to the end.  Donors that are NaN are handled by anynan;
    # donors that are +inf are real values and sort among the padding, which
    # is harmless because the k-th order statistic of the padded row equals
    # that of the donors for k < ngood.
    key = np.where(good, vals, np.inf)
    key = np.where(np.isnan(key), np.inf, key)
    key.sort(axis=1)
    idx = np.arange(nbad)
    ng = np.maximum(ngood, 1)
    lo = key[idx, (ng - 1) // 2]
    hi = key[idx, ng // 2]
    with np.errstate(all="ignore"):
        med = np.where(ng % 2 == 1, lo, (lo + hi) / np.asarray(2, dtype=data.dtype))
    med = med.astype(data.dtype)
    med = np.where(anynan, np.asarray(np.nan, dtype=data.dtype), med)
    fix = ngood >= min_valid
    newdata[rr[fix], cc[fix]] = med[fix]
    return newdata, _stats(data.size, nbad, int(nbad - fix.sum()), deltapix, min_valid)


def user_mask(shape, bad_columns=None, bad_rows=None, bad_rectangles=None, mask=None):
    """ApFindBadPixels._add_bad_columns/_rows/_rectangles (:70-158).

    Returns ``(mask uint8, num_user_bad)`` with the reference's counting rule
    (every in-range entry counts its full area, overlaps counted twice)."""
    nrows, ncols = shape
    if mask is None:
        mask = np.zeros(shape, dtype=np.uint8)
    nuser = 0
    for col in (bad_columns or []):
        c1 = col - 1
        if 0 <= c1 < ncols:
            mask[:, c1:col] += USER_BAD
            nuser += nrows
    for row in (bad_rows or []):
        r1 = row - 1
        if 0 <= r1 < nrows:
            mask[r1:row, :] += USER_BAD
            nuser += ncols
    for rect in (bad_rectangles or []):
        if len(rect) != 4:
            continue
        r1, r2, c1, c2 = rect[0] - 1, rect[1], rect[2] - 1, rect[3]
        if r1 < 0 or r2 > nrows or c1 < 0 or c2 > ncols:
            continue
        mask[r1:r2, c1:c2] += USER_BAD
        nuser += (r2 - r1) * (c2 - c1)
    return mask, nuser


def auto_mask(dark, sigma):
    """ApFindBadPixels._generate_sigmaclip_mask (:171-217)."""
    from .combine_oracle import sigma_clipped_stats_global
    mean, med, std = sigma_clipped_stats_global(dark, sigma=sigma)
    lo = med - (sigma * std)
    hi = med + (sigma * std)
    with np.errstate(invalid="ignore"):
        m = np.logical_or(dark < lo, dark > hi).astype("uint8")
    return m, int(m.sum()), (mean, med, std)
