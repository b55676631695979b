"""ApFindBadPixels: bad-pixel mask from a master dark/bias plus user-defined regions.

Host-side mirror of ``AstroPhotography/core/ApFindBadPixels.py`` of the
reference (ctor :30-68, ``add_user_badpix`` :414-438, ``get_mask`` :440,
``write_mask`` :445-473).  The whole-image sigma-clipped statistics
(``astropy.stats.sigma_clipped_stats``, :191) and the threshold mask (:194-209)
run on the GPU (``apgpu_sigma_clipped_stats_f32`` / ``apgpu_threshold_mask_f32``);
the user rules are a handful of integer slice increments and stay on the host.

Mask values as in the reference: 0 good, +1 algorithmically bad, +2 for every
user column / row / rectangle that covers the pixel (overlaps accumulate).
"""
from __future__ import annotations

from datetime import datetime, timezone
from pathlib import Path

import numpy as np
import yaml

from .. import _native, fitsio, kernels
from ..version import __version__
from ._base import ApBase


class ApFindBadPixels(ApBase):
    GOOD = 0
    AUTO_BAD = 1
    USER_BAD = 2
    _name = "ApFindBadPixels"

    def __init__(self, darkfile, sigma, loglevel):
        self._loglevel = loglevel
        self._initialize_logger(loglevel)
        self._imfile = darkfile
        self._imextnum = 0
        self._userfile = None
        self._nbad_auto = 0
        self._nbad_user = 0
        self._sigma = float(sigma)        # the reference's CLI passes a str here (ap_find_badpix.py:52-58)
        self._imdata, self._imhdr, _ = self._read_fits(darkfile, 0)
        self._generate_sigmaclip_mask(self._imdata, self._sigma)

    # -- automatic mask -------------------------------------------------------
    def _generate_sigmaclip_mask(self, data, sigma):
        torch = _native.require_cuda()
        npix = data.size
        self._logger.debug(f"Generating a bad pixel mask using sigma={sigma} clipping on the input image data values.")
        if data.dtype != np.float32:
            data = data.astype(np.float32)
        dev = torch.from_numpy(np.ascontiguousarray(data)).cuda()
        mean, med, std, _ = kernels.sigma_clipped_stats(dev, sigma=sigma)
        self._logger.debug(f"Sigma-clipped mean={mean:.2f}, median={med:.2f}, and madstddev={std:.2f} values (ADU).")
        lothresh = med - (sigma * std)
        hithresh = med + (sigma * std)
        self._logger.info(f"Good pixels have values between {lothresh:.2f} and {hithresh:.2f} ADU.")
        mask, nbad = kernels.threshold_mask(dev, lothresh, hithresh)
        self._badpixmask = mask.cpu().numpy()
        self._clipped_stats = (mean, med, std)
        pct_bad = 100 * (nbad / npix)
        self._logger.info(f"Out of {npix} pixels, {nbad} are bad ({pct_bad:.4f}%).")
        self._nbad_auto = nbad

    # -- user-defined regions ---------------------------------------------------
    def _add_bad_columns(self, bad_col_list):
        nrows, ncols = self._badpixmask.shape
        self._logger.info(f"Adding {len(bad_col_list)} bad columns to {nrows} row, {ncols} column mask.")
        num_user_bad = 0
        for col in bad_col_list:                 # 1-based, as reported by ds9
            c0 = col - 1
            if c0 < 0 or c0 >= ncols:
                self._logger.warning(f"Warning, column {col} (1-based) outside image.")
                continue
            self._badpixmask[:, c0:col] += ApFindBadPixels.USER_BAD
            num_user_bad += nrows
        return num_user_bad

    def _add_bad_rows(self, bad_row_list):
        nrows, ncols = self._badpixmask.shape
        self._logger.info(f"Adding {len(bad_row_list)} bad rows to {nrows} row, {ncols} column mask.")
        num_user_bad = 0
        for row in bad_row_list:
            r0 = row - 1
            if r0 < 0 or r0 >= nrows:
                self._logger.warning(f"Warning, row {row} (1-based) outside image.")
                continue
            self._badpixmask[r0:row, :] += ApFindBadPixels.USER_BAD
            num_user_bad += ncols
        return num_user_bad

    def _add_bad_rectangles(self, bad_rectangle_list):
        nrows, ncols = self._badpixmask.shape
        self._logger.info(f"Adding {len(bad_rectangle_list)} bad rectangles to {nrows} row, {ncols} column mask.")
        num_user_bad = 0
        for rect in bad_rectangle_list:
            if len(rect) != 4:
                self._logger.warning(f"Error, expecting 4-element list, got {rect}. Skipping.")
                continue
            # inclusive 1-based [row_start, row_end, col_start, col_end] -> half-open 0-based
            r0, r1, c0, c1 = rect[0] - 1, rect[1], rect[2] - 1, rect[3]
            if r0 < 0 or r1 > nrows:
                self._logger.warning(f"Warning, row range {r0}:{r1} (0-based) outside image.")
            elif c0 < 0 or c1 > ncols:
                self._logger.warning(f"Warning, column range {c0}:{c1} (0-based) outside image.")
            else:
                self._badpixmask[r0:r1, c0:c1] += ApFindBadPixels.USER_BAD
                num_user_bad += (r1 - r0) * (c1 - c0)
        return num_user_bad

    def _read_user_badpix(self, user_badpix_file):
        path = self._check_file_exists(user_badpix_file)
        with open(path) as f:
            yobj = yaml.safe_load(f.read()) or {}
        out = []
        for key in ("bad_columns", "bad_rows", "bad_rectangles"):
            # The reference raises TypeError when a key is absent (len(None), :362-367);
            # here an absent key simply means "none", like the documented ``{}``.
            val = yobj.get(key) or None
            self._logger.debug(f"{key}: {0 if val is None else len(val)} entries in the user-defined badpixel file.")
            out.append(val)
        return tuple(out)

    def add_user_badpix(self, user_badpix_file):
        user_badpix_file = Path(user_badpix_file).expanduser()
        self._logger.info(f"Processing user-defined bad pixels from {user_badpix_file}")
        badcols, badrows, badrect = self._read_user_badpix(user_badpix_file)
        self._userfile = user_badpix_file
        num_user_bad = 0
        if badcols is not None:
            num_user_bad += self._add_bad_columns(badcols)
        if badrows is not None:
            num_user_bad += self._add_bad_rows(badrows)
        if badrect is not None:
            num_user_bad += self._add_bad_rectangles(badrect)
        self._nbad_user = num_user_bad
        self._logger.debug(f"Total number of user-defined bad pixels applied to mask: {num_user_bad}")

    # -- results ------------------------------------------------------------------
    def get_mask(self):
        """The bad pixel mask as a uint8 numpy array."""
        return self._badpixmask

    def write_mask(self, mask_file_name):
        hdr = fitsio.new_header()
        hdr["IMAGETYP"] = ("BADPIX", "Type of file")
        hdr["CREATOR"] = (self._name, "Software that generated this file.")
        hdr["DATE"] = (datetime.now(timezone.utc).isoformat(timespec="seconds"), "UTC creation time.")
        hdr["DATAFILE"] = (str(self._imfile), "Data file used to identify bad pixels.")
        if self._userfile is not None:
            hdr["USERFILE"] = (self._userfile.name, "User-defined bad pixel file.")
        hdr["NBADAUTO"] = (int(self._nbad_auto), "Number of algorithm-detected bad pixels.")
        hdr["NBADUSER"] = (int(self._nbad_user), "Number of user-defined bad pixels.")
        for kw in ("TELESCOP", "INSTRUME", "SET-TEMP", "CCD-TEMP", "XPIXSZ", "YPIXSZ", "XBINNING",
                   "YBINNING", "XORGSUBF", "YORGSUBF", "SITELAT", "SITELONG"):
            if kw in self._imhdr:
                hdr[kw] = (self._imhdr[kw], self._imhdr.comments[kw])
        tnow = datetime.now().isoformat(timespec="milliseconds")
        hdr["HISTORY"] = f"Processed by {self._name} {__version__} at {tnow}"
        fitsio.write_image(mask_file_name, self._badpixmask, hdr, overwrite=True)
        self._logger.info(f"Wrote bad pixel mask to {mask_file_name}")
