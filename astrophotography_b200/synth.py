"""Seeded synthetic calibration / science frames.

Shared by the tests, ``bench.py`` and ``__graft_entry__.smoke()`` so that the
CUDA path, the oracle and the CPU baseline all see identical inputs.  There is
no network and the reference ships no FITS fixtures, so every input on this
path is generated here from ``numpy.random.default_rng(seed)``:

* dark / bias frame ``k``: ``f32(1000 + rate*t + N(0, 12))`` + fixed hot pixels
  (0.05 % at +5000*U(0.5,1), same positions in every frame) + per-frame cosmic
  hits (0.01 % at +U(500,30000)); optionally rounded to integers (real frames
  are uint16, ``/root/reference/doc/fits_metadata.md:70-75``) to create ties;
* flat: ``30000 * vignette(r) * (1 + N(0, 0.01))`` with ~10 exact zeros and one NaN;
* science: sky 2000 + a few Gaussian stars + noise, uint16 or float32;
* mask: the ``etc/user_badpixels.yml`` rules scaled to the frame + 0.1 % random
  auto-bad pixels.

Seeds: frames ``1000+k``, flat ``7``, science ``11``, mask ``13``.
"""
from __future__ import annotations

import numpy as np

BIAS_LEVEL = 1000.0
READ_NOISE = 12.0
HOT_FRACTION = 5e-4
COSMIC_FRACTION = 1e-4


def hot_pixels(shape, seed=999):
    rng = np.random.default_rng(seed)
    n = int(round(HOT_FRACTION * shape[0] * shape[1]))
    idx = rng.choice(shape[0] * shape[1], size=n, replace=False)
    amp = (5000.0 * rng.uniform(0.5, 1.0, size=n)).astype(np.float32)
    return idx, amp


def dark_frame(k, shape, exptime=0.0, dark_rate=0.1, quantise=False, hot=None):
    """Frame ``k`` of a bias (``exptime=0``) or dark series, float32 (H, W)."""
    rng = np.random.default_rng(1000 + k)
    h, w = shape
    img = rng.normal(BIAS_LEVEL + dark_rate * exptime, READ_NOISE, size=(h, w)).astype(np.float32)
    flat = img.reshape(-1)
    if hot is None:
        hot = hot_pixels(shape)
    flat[hot[0]] += hot[1] * np.float32(max(exptime, 1.0) / 300.0 if exptime else 1.0)
    ncos = int(round(COSMIC_FRACTION * h * w))
    if ncos:
        cidx = rng.integers(0, h * w, size=ncos)
        flat[cidx] += rng.uniform(500.0, 30000.0, size=ncos).astype(np.float32)
    if quantise:
        np.rint(img, out=img)
        np.clip(img, 0, 65535, out=img)
    return img


def dark_stack(n, shape, exptime=0.0, quantise=False):
    hot = hot_pixels(shape)
    return np.stack([dark_frame(k, shape, exptime, quantise=quantise, hot=hot)
                     for k in range(n)])


def flat_frame(shape, seed=7, with_specials=True):
    rng = np.random.default_rng(seed)
    h, w = shape
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    r2 = ((yy - h / 2) ** 2 + (xx - w / 2) ** 2) / np.float32((h / 2) ** 2 + (w / 2) ** 2)
    vign = 1.0 - 0.3 * r2
    flat = (30000.0 * vign * (1.0 + rng.normal(0, 0.01, size=(h, w)))).astype(np.float32)
    if with_specials:
        f = flat.reshape(-1)
        z = rng.choice(h * w, size=min(11, h * w), replace=False)
        f[z[:-1]] = 0.0
        f[z[-1]] = np.nan
    return flat


def science_frame(shape, seed=11, as_uint16=True, nstars=40):
    rng = np.random.default_rng(seed)
    h, w = shape
    img = rng.normal(BIAS_LEVEL + 2000.0, 45.0, size=(h, w))
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(nstars):
        cy, cx = rng.uniform(0, h), rng.uniform(0, w)
        amp, sig = rng.uniform(500, 40000), rng.uniform(1.2, 3.0)
        y0, y1 = int(max(0, cy - 6 * sig)), int(min(h, cy + 6 * sig + 1))
        x0, x1 = int(max(0, cx - 6 * sig)), int(min(w, cx + 6 * sig + 1))
        if y1 > y0 and x1 > x0:
            img[y0:y1, x0:x1] += amp * np.exp(
                -((yy[y0:y1, x0:x1] - cy) ** 2 + (xx[y0:y1, x0:x1] - cx) ** 2) / (2 * sig * sig))
    if as_uint16:
        return np.clip(np.rint(img), 0, 65535).astype(np.uint16)
    return img.astype(np.float32)


# etc/user_badpixels.yml:33-51 of the reference, as data (1-based inclusive).
USER_BADPIX_EXAMPLE = {
    "bad_columns": [12, 13, 17],
    "bad_rectangles": [[1, 1, 1, 1], [5, 6, 7, 12], [200, 300, 400, 420]],
    "bad_rows": {},
}


def scaled_user_rules(shape):
    """The example rules clipped so that they fall inside ``shape``."""
    h, w = shape
    cols = [c for c in USER_BADPIX_EXAMPLE["bad_columns"] if c <= w]
    rects = [r for r in USER_BADPIX_EXAMPLE["bad_rectangles"] if r[1] <= h and r[3] <= w]
    if h < 300 or w < 420:        # keep one 'large unfixable interior' rectangle
        r0, c0 = max(1, h // 2), max(1, w // 2)
        rects.append([r0, min(h, r0 + max(6, h // 8)), c0, min(w, c0 + max(6, w // 16))])
    return cols, [], rects


def badpix_mask(shape, seed=13, auto_fraction=1e-3):
    """uint8 mask: user rules (+2 each) + random auto-bad pixels (+1)."""
    rng = np.random.default_rng(seed)
    h, w = shape
    mask = np.zeros(shape, dtype=np.uint8)
    cols, rows, rects = scaled_user_rules(shape)
    for c in cols:
        mask[:, c - 1:c] += 2
    for r in rows:
        mask[r - 1:r, :] += 2
    for r1, r2, c1, c2 in rects:
        mask[r1 - 1:r2, c1 - 1:c2] += 2
    nauto = int(round(auto_fraction * h * w))
    if nauto:
        idx = rng.choice(h * w, size=nauto, replace=False)
        mask.reshape(-1)[idx] += 1
    return mask
