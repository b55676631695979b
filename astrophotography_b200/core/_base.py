"""Shared plumbing for the Ap* host classes: logger set-up, FITS read/convert,
FITS write with keyword dictionaries.

The reference gives every class its own copy of these helpers
(``core/ApCalibrate.py:230-328,348-404``, ``core/ApFixBadPixels.py:62-243``,
``core/ApFindBadPixels.py:236-323``); here they live once.  Behaviour kept:
logger name = class name, format ``asctime | name | level | message``,
``ValueError('Invalid log level: ...')`` on a bad level, ``RuntimeError`` with
the reference's wording for a missing file, integer data -> float32 and
``PEDESTAL`` added on read, ``PEDESTAL`` removed from written headers.
Not kept: a new StreamHandler per construction (the reference prints every line
k times after k constructions, ``ApCalibrate.py:243-253``) and the eager
full-frame ``nanmedian`` on every read (``:310-313``), which is a hidden
full-frame sort -- statistics are computed only when DEBUG logging is on.
"""
from __future__ import annotations

import logging
from datetime import datetime
from pathlib import Path

import numpy as np

from .. import fitsio
from ..version import __version__

LOG_FORMAT = "%(asctime)s | %(name)s | %(levelname)s | %(message)s"


def make_logger(name: str, loglevel: str) -> logging.Logger:
    logger = logging.getLogger(name)
    numeric_level = getattr(logging, str(loglevel).upper(), None)
    if not isinstance(numeric_level, int):
        raise ValueError("Invalid log level: {}".format(loglevel))
    logger.setLevel(numeric_level)
    if not logger.handlers:
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter(LOG_FORMAT))
        logger.addHandler(handler)
    for h in logger.handlers:
        h.setLevel(numeric_level)
    logger.propagate = False
    return logger


class ApBase:
    """Logger + FITS helpers shared by ApCalibrate / ApFixBadPixels / ApFindBadPixels / ApMasterCal."""

    _name = "ApBase"

    def _initialize_logger(self, loglevel):
        self._logger = make_logger(self._name, loglevel)

    def _check_file_exists(self, filename):
        path = Path(filename).expanduser()
        if not path.exists():
            err_msg = f"Cannot find {filename}. Not a valid path or file."
            self._logger.error(err_msg)
            raise RuntimeError(err_msg)
        return path

    def _read_fits(self, image_filename, image_extension=0, to_float=True, apply_pedestal=True):
        """Read one image HDU.  Returns ``(data, header, pedestal)``.

        With ``to_float`` non-floating data becomes float32 and a non-zero
        ``PEDESTAL`` is added (reference ``_read_fits``).  With
        ``to_float=False`` integer data is returned untouched together with
        the pedestal still to be applied, so that the conversion can be fused
        into a kernel."""
        path = self._check_file_exists(image_filename)
        self._logger.info("Loading extension {} of FITS file {}".format(image_extension, image_filename))
        data, hdr = fitsio.read_image(path, image_extension)
        ndim = hdr["NAXIS"]
        info = "{}-D BITPIX={} image with {} columns, {} rows".format(
            ndim, hdr["BITPIX"], hdr["NAXIS1"], hdr.get("NAXIS2", 1))
        for kw in ("BSCALE", "BZERO"):
            if kw in hdr:
                info += f", {kw}={hdr[kw]}"
        self._logger.debug(info)
        if ndim == 3:
            msg = "Error, 3-D handling has not been implemented yet."
            self._logger.error(msg)
            raise SystemExit(1)          # the reference calls sys.exit(1) here
        pedestal = 0.0
        if "PEDESTAL" in hdr:
            pedestal = float(hdr["PEDESTAL"])
        if to_float:
            if not np.issubdtype(data.dtype, np.floating):
                orig = data.dtype
                data = data.astype(np.float32)
                self._logger.debug(f"  Converted data type from {orig} to float32")
            elif not data.flags.writeable or not data.flags.c_contiguous:
                data = np.ascontiguousarray(data).copy()
            if apply_pedestal and pedestal != 0:
                self._logger.debug(f"Removing a PEDESTAL value of {pedestal} ADU.")
                data += pedestal
                pedestal = 0.0
        if self._logger.isEnabledFor(logging.DEBUG):
            with np.errstate(all="ignore"):
                self._logger.debug("Raw data statistics are min={:.2f}, max={:.2f}, median={:.2f}".format(
                    float(np.nanmin(data)), float(np.nanmax(data)), float(np.nanmedian(data))))
        return data, hdr, pedestal

    def _header_like(self, hdr, odict, history, drop_scaling=True, only_bpix=False):
        """A copy of ``hdr`` without PEDESTAL (and the scaling keywords), plus ``odict`` and a HISTORY line."""
        hdr = hdr.copy()
        drop = ["PEDESTAL"] + (["BSCALE", "BZERO"] if drop_scaling else [])
        for kw in drop:
            if kw in hdr:
                del hdr[kw]
        for kw, val in odict.items():
            if only_bpix and "BPIX" not in kw:
                continue
            hdr[kw] = _plain(val)
        tnow = datetime.now().isoformat(timespec="milliseconds")
        hdr["HISTORY"] = f"{history} {__version__} at {tnow}"
        return hdr

    def _write_image_like(self, inpdata_file, ext_num, outdata_file, odata, odict,
                          history, drop_scaling=True, only_bpix=False):
        """Write ``odata`` with the header of ``inpdata_file`` plus ``odict``."""
        self._logger.debug(f"FITS header keywords added to output: {odict}")
        path = self._check_file_exists(inpdata_file)
        hdr = self._header_like(fitsio.read_header(path, ext_num), odict, history, drop_scaling, only_bpix)
        fitsio.write_image(outdata_file, odata, hdr, overwrite=True)


def _plain(val):
    """numpy scalars -> Python scalars inside (value, comment) pairs."""
    if isinstance(val, tuple):
        return (_plain(val[0]),) + tuple(val[1:])
    if isinstance(val, np.generic):
        return val.item()
    return val
