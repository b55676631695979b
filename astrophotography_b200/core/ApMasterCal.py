"""ApMasterCal: combine raw bias / dark / flat frames into a master frame on the GPU.

Host-side mirror of the reference class in ``AstroPhotography/scripts/ap_combine_darks.py``
(:100-441): same constructor ``(rootdir, exclude_pattern, telescop, temptol,
loglevel)``, same consistency checks on ``telescop, imagetyp, naxis1, naxis2,
exptime, set-temp`` (:207-212, RuntimeError), same CCD-TEMP tolerance filter
(:270-284), same output keywords (``IMAGETYP 'MASTER BIAS|DARK|FLAT'``,
``TELESCOP, CREATOR, SET-TEMP, CCD-TEMP, DATE, IFILEnnn`` :339-352; ``UT, TIME-OBS,
SWOWNER, SWCREATE, SBSTDVER`` removed :426-430) and a three-HDU output like the
``CCDData`` the reference writes (primary data, ``MASK`` uint8, ``UNCERT``).

The one call the reference makes for the arithmetic -- ``ccdproc.combine(files,
method='average', sigma_clip=True, low=5, high=5, func=np.ma.median,
dev_func=mad_std, mem_limit=5e8)`` (:411-420) -- becomes a double-buffered
host->GPU->host row-band pipeline around ``apgpu_stack_reduce_f32``
(``pipeline.HostStackCombiner``): every frame is read once (ccdproc re-reads every
file for every memory chunk), the defaults are the reference's (:394-399), and the
additive keyword arguments expose the other combine settings.

ccdproc's ``ImageFileCollection`` is replaced by a small header scan
(``_FileCollection``); FITS access goes through ``fitsio`` (astropy when present).
"""
from __future__ import annotations

import fnmatch
import os
from datetime import datetime, timezone
from pathlib import Path

import numpy as np

from .. import fitsio, pipeline
from ._base import ApBase

_FITS_SUFFIXES = (".fit", ".fits", ".fts")


class _FileCollection:
    """Header summary of the FITS files of one directory (what the reference gets
    from ``ccdproc.ImageFileCollection(dir, keywords, glob_exclude, filenames)``)."""

    def __init__(self, location, keywords, glob_exclude=None, filenames=None):
        self.location = Path(location)
        self.keywords = list(keywords)
        if filenames is not None:
            names = list(filenames)
        else:
            names = sorted(f for f in os.listdir(self.location)
                           if f.lower().endswith(_FITS_SUFFIXES)
                           and not (glob_exclude and fnmatch.fnmatch(f, glob_exclude)))
        self.summary = {kw: [] for kw in self.keywords}
        for name in names:
            hdr = fitsio.read_header(self.location / name, 0)
            for kw in self.keywords:
                if kw == "file":
                    self.summary[kw].append(name)
                else:
                    key = kw.upper()
                    self.summary[kw].append(hdr[key] if key in hdr else "")

    def __len__(self):
        return len(self.summary["file"])

    def values(self, keyword, unique=False):
        vals = list(self.summary[keyword])
        if not unique:
            return vals
        seen = []
        for v in vals:
            if v not in seen:
                seen.append(v)
        return seen

    def files_filtered(self, include_path=False):
        return [str(self.location / f) if include_path else f for f in self.summary["file"]]


class ApMasterCal(ApBase):
    """Combines a series of darks, biases or flats into a master frame.

    Like the reference it refuses directories holding mixed frame types, sizes,
    exposure times or set temperatures (RuntimeError)."""

    _name = "ApMasterCal"

    def __init__(self, rootdir, exclude_pattern, telescop, temptol, loglevel, *,
                 method="average", sigma_clip=True, sigma_clip_low_thresh=5.0, sigma_clip_high_thresh=5.0,
                 maxiters=1, cenfunc="median", devfunc="mad_std", out_dtype="float64"):
        self._loglevel = loglevel
        self._initialize_logger(loglevel)
        self._rootdir = rootdir
        self._telescop = telescop
        self._temptol = float(temptol)       # the reference's CLI hands a str through (:90-95)
        self._combine = dict(method=method, k_lo=float(sigma_clip_low_thresh), k_hi=float(sigma_clip_high_thresh),
                             maxiters=int(maxiters) if sigma_clip else 0, cen=cenfunc, dev=devfunc)
        self._out_f64 = np.dtype(out_dtype) == np.float64
        self._set_temperature = ""
        self._summary_kw = ["file", "date-obs", "telescop", "imagetyp", "filter", "exptime",
                            "set-temp", "ccd-temp", "naxis1", "naxis2"]
        self._data_dir = Path(rootdir)
        if not self._data_dir.is_dir():
            msg = f"Cannot find {rootdir}. Not a valid path or file."
            self._logger.error(msg)
            raise RuntimeError(msg)
        self._file_list = None
        self._files = self._create_file_collection(self._data_dir, exclude_pattern, self._file_list)
        self._file_list = self._check_files(True)
        self._files = self._create_file_collection(self._data_dir, exclude_pattern, self._file_list)
        self._logger.debug("ApMasterCal constructor completed.")

    def _create_file_collection(self, data_dir, exclude_pattern, file_list):
        if file_list is not None:
            self._logger.info(f"Looking for FITS files in {data_dir}, including only files in the list: {file_list}")
        else:
            self._logger.info(f'Looking for FITS files in {data_dir}, excluding files matching the pattern "{exclude_pattern}"')
        coll = _FileCollection(data_dir, self._summary_kw, exclude_pattern, file_list)
        self._logger.info(f"Found {len(coll)} FITS files matching the constraints.")
        return coll

    def _check_files(self, list_all=None):
        raw_file_list = self._files.values("file")
        if not raw_file_list:
            msg = f"Error, no FITS files found in {self._data_dir}."
            self._logger.error(msg)
            raise RuntimeError(msg)
        uniq = {kw: self._files.values(kw, unique=True) for kw in self._summary_kw}
        if list_all:
            for kw, vals in uniq.items():
                if kw != "file":
                    self._logger.debug(f"For keyword {kw} there are {len(vals)} values: {vals}")
        for kw in ("telescop", "imagetyp", "naxis1", "naxis2", "exptime", "set-temp"):
            if len(uniq[kw]) > 1:
                msg = (f"Error, there are {len(uniq[kw])} unique values of {kw} in the files being"
                       f" processed: {uniq[kw]}")
                self._logger.error(msg)
                raise RuntimeError(msg)
        self._imgtype = str(uniq["imagetyp"][0])
        self._exptime = uniq["exptime"][0]
        telescop = str(uniq["telescop"][0])
        if not telescop.strip():
            self._logger.warning(f"TELESCOP keyword empty or missing in input files. Using {self._telescop} instead.")
        else:
            self._telescop = telescop.strip()
        set_temperature = None
        val = uniq["set-temp"][0]
        if isinstance(val, str):
            if not val.strip():
                self._logger.warning("No numeric value found for SET-TEMP. Will use median of CCD-TEMP instead.")
            else:
                try:
                    set_temperature = float(val)
                except ValueError:
                    self._logger.error(f"Error, could not convert SET-TEMP value of {val} to a float."
                                       " Will use median of CCD-TEMP instead.")
        else:
            set_temperature = float(val)
        temps = self._files.values("ccd-temp")
        if all(isinstance(t, str) and not t.strip() for t in temps):
            self._logger.warning("No files contain CCD-TEMP metadata. Continuing assuming all files"
                                 " obtained at the same temperature.")
            if set_temperature is not None:
                self._set_temperature = set_temperature
            return raw_file_list
        temps = [float(t) for t in temps]
        if set_temperature is None:
            set_temperature = float(np.median(temps))
            self._logger.debug(f"Using median of CCD-TEMP value: {set_temperature} degrees C.")
        temp_min, temp_max = set_temperature - self._temptol, set_temperature + self._temptol
        self._logger.info(f"Selecting only files with CCD-TEMP between {temp_min:.2f} and {temp_max:.2f} degrees C.")
        self._set_temperature = set_temperature
        good = []
        for fname, temp in zip(raw_file_list, temps):
            if temp_min <= temp <= temp_max:
                good.append(fname)
            else:
                # (the reference crashes here: self.logger / '{temp:2.f}', ap_combine_darks.py:284)
                self._logger.warning(f"Excluding {fname} as CCD-TEMP={temp:.2f} outside allowed range.")
        self._logger.info(f"Updated file list contains {len(good)} files ({len(raw_file_list)} before filtering).")
        if not good:
            msg = "Error, no files left after CCD-TEMP filtering."
            self._logger.error(msg)
            raise RuntimeError(msg)
        return good

    def _generate_final_keywords(self):
        creation = datetime.now(timezone.utc).isoformat(timespec="seconds")
        raw_imgtype = self._imgtype.lower()
        imgtype = raw_imgtype
        for key in ("bias", "dark", "flat"):
            if key in raw_imgtype:
                imgtype = f"MASTER {key.upper()}"
                break
        else:
            self._logger.warning(f"Unexpected input image type: {raw_imgtype}")
        kw = {"IMAGETYP": (imgtype, "Type of file"),
              "TELESCOP": (self._telescop, "Telescope used."),
              "CREATOR": ("ApMasterCal", "Software that generated this file."),
              "SET-TEMP": (self._set_temperature, "[Celsius] Desired CCD temperature"),
              "CCD-TEMP": (self._set_temperature, "[Celsius] Desired CCD temperature"),
              "DATE": (creation, "Date/time file was created.")}
        for idx, fname in enumerate(self._files.values("file")):
            kw[f"IFILE{idx:03d}"] = fname
        return kw

    def _load_frames(self):
        """The frames as float32 host arrays (array-level users; ``make_master`` streams the files instead)."""
        frames = []
        for path in self._files.files_filtered(include_path=True):
            data, _hdr = fitsio.read_image(path, 0)
            if data.ndim != 2:
                msg = f"Error, {path} is not a 2-D image."
                self._logger.error(msg)
                raise RuntimeError(msg)
            frames.append(np.ascontiguousarray(data, dtype=np.float32))
        return frames

    def combine_frames(self, frames):
        """Combine in-memory (H,W) frames (float32, or uint16 raw frames) with this object's settings.
        Returns the dict of host arrays ``data, nrej, uncert, allmasked``.  The double-buffered combiner
        (device buffers, pinned result planes) is kept and reused while the stack geometry stays the same."""
        n = len(frames)
        h, w = frames[0].shape
        dtype = np.dtype(frames[0].dtype)
        key = (n, h, w, dtype.str)
        if getattr(self, "_combiner_key", None) != key:
            self._combiner = pipeline.HostStackCombiner(n, h, w, out_f64=self._out_f64, want_nrej=True, want_uncert=True,
                                                        want_allmasked=True, dtype=dtype, **self._combine)
            self._combiner_key = key
        res = self._combiner.combine(frames)
        return {k: np.array(v, copy=True) for k, v in res.items()}

    def combine_files(self, gpus=1):
        """Combine the files of this object's collection: every file is read once, straight into page-locked
        memory, while the previous frames upload (raw BITPIX=16 frames travel undecoded, 2 bytes per pixel);
        ``gpus > 1`` shards the rows over that many GPUs of this box (one process each)."""
        paths = self._files.files_filtered(include_path=True)
        opts = dict(out_f64=self._out_f64, want_nrej=True, want_uncert=True, want_allmasked=True, **self._combine)
        try:
            if gpus and int(gpus) > 1:
                return pipeline.combine_files_sharded(paths, int(gpus), **opts)
            src = pipeline.FileFrames(paths, fitsio)
            self._logger.debug(f"Streaming {src.n} frames of {src.w}x{src.h} as {src.dtype} ({src.u16_format}).")
            res = pipeline.combine_files(src, **opts)
        except RuntimeError as exc:
            if "not a 2-D image" in str(exc) or "has shape" in str(exc):
                self._logger.error(str(exc))
            raise
        return {k: np.array(v, copy=True) for k, v in res.items()}

    def make_master(self, output_master_file, gpus=1):
        """Combine the raw calibration files into a master file and write it."""
        kw_dict = self._generate_final_keywords()
        cal_type = kw_dict["IMAGETYP"][0]
        nfiles = len(self._files)
        c = self._combine
        self._logger.debug(f"About to combine {nfiles} {cal_type} files, method={c['method']} sigma_clip={c['maxiters'] != 0}"
                           f" sig_clip_lothresh={c['k_lo']} sig_clip_hithresh={c['k_hi']}.")
        res = self.combine_files(gpus)
        first = self._files.files_filtered(include_path=True)[0]
        hdr = fitsio.read_header(first, 0).copy()
        # PEDESTAL stays, as in the reference: ccdproc.combine never applies it to the frames and keeps the
        # first file's header, and ApCalibrate._read_fits (core/ApCalibrate.py:318-326) then removes the
        # pedestal from the master exactly as it does from a raw frame
        for kw in ("UT", "TIME-OBS", "SWOWNER", "SWCREATE", "SBSTDVER", "BZERO", "BSCALE"):
            if kw in hdr:
                del hdr[kw]
        hdr["NCOMBINE"] = (nfiles, "Number of frames combined")
        hdr["BUNIT"] = ("adu", "Pixel value units.")
        hdr["COMBINED"] = (True, "Produced by a stack combine")
        for kw, val in kw_dict.items():
            hdr[kw] = val
        fitsio.write_image(output_master_file, res["data"], hdr,
                           extensions=[("MASK", res["allmasked"]), ("UNCERT", res["uncert"])], overwrite=True)
        self._last_result = res
        self._logger.info(f"Wrote combined calibration file: {output_master_file}")
