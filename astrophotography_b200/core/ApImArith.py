"""ApImArith: ADD / SUB / MUL / DIV of a FITS image with a constant or another image, on the GPU.

Host-side mirror of the reference class ``AstroPhotography/core/ApImArith.py`` (ctor :26-41,
``process_files`` :255-346): same constructor, same ``process_files(inp_img, operation, value, out_img,
units)``, same errors (``ValueError`` for an unknown operation :188-191 or a value that is neither a number nor
a FITS file :288-303), same output (a copy of the input file with the primary data replaced, ``PEDESTAL``
removed, ``BUNIT`` set when ``units`` is given, two HISTORY lines :240-241).  The four numpy ufunc calls
(:321-333) become one launch of ``apgpu_imarith_f32``.

The reference keeps the input's dtype for the result (``np.zeros(data1.shape, dtype=data1.dtype)`` :320); the
GPU path covers what the pipeline feeds it -- float32 images (calibrated frames, ``calibrate_all.sh:435-464``
subtracts the sky background with ``ap_imarith.py SUB``) with a scalar, a float32 or a float64 second image --
and raises ``RuntimeError`` for any other input dtype instead of silently computing something else.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from .. import _native, kernels
from ._base import ApBase


class ApImArith(ApBase):
    """A utility class with the basic functionality of the HEASOFT fimarith / fcarith tools."""

    _name = "ApImArith"

    def __init__(self, loglevel):
        self._allowed_ops = ["ADD", "SUB", "MUL", "DIV"]
        self._loglevel = loglevel
        self._initialize_logger(loglevel)
        self._logger.debug(f"{self._name} instance constructed.")

    @staticmethod
    def _is_float(value_str):
        try:
            float(value_str)
            return True
        except (ValueError, TypeError):
            return False

    def _sanitize_operation(self, operation):
        cleaned = str(operation).strip().upper()
        if cleaned not in self._allowed_ops:
            raise ValueError(f"Error, input operation {cleaned} is not one of the allowed operations: {self._allowed_ops}")
        return cleaned

    # -- array level (additive) ----------------------------------------------------------------------------
    def apply(self, data1, operation, data2):
        """``np.<operation>(data1, data2, out=float32 array)`` for a float32 image (numpy array or CUDA tensor)
        and a number / float32 / float64 image; the result has the type of ``data1``."""
        torch = _native.require_cuda()
        operation = self._sanitize_operation(operation)
        on_device = isinstance(data1, torch.Tensor)
        if not on_device:
            data1 = np.asarray(data1)
            if data1.dtype != np.float32:
                msg = (f"{self._name}: input image dtype {data1.dtype} is not supported on the GPU path"
                       " (float32 images only).")
                self._logger.error(msg)
                raise RuntimeError(msg)
        a = data1 if on_device else torch.from_numpy(np.ascontiguousarray(data1)).cuda()
        if isinstance(data2, (int, float, np.floating, np.integer)):
            b = float(data2)
        elif isinstance(data2, torch.Tensor):
            b = data2
        else:
            arr = np.asarray(data2)
            if arr.dtype not in (np.float32, np.float64):
                arr = arr.astype(np.float64)            # numpy would promote an integer image against float32 to float64
            b = torch.from_numpy(np.ascontiguousarray(arr)).to(a.device)
        out = kernels.imarith(a, operation, b)
        return out if on_device else out.cpu().numpy()

    # -- file level (reference signature) ------------------------------------------------------------------
    def process_files(self, inp_img, operation, value, out_img, units):
        operation = self._sanitize_operation(operation)
        data1, hdr1, _ = self._read_fits(inp_img, 0, to_float=False)
        if "PEDESTAL" in hdr1 and float(hdr1["PEDESTAL"]) != 0:
            # reference _read_fits :139-147: ext_data += pedestal (in the image's own dtype)
            data1 = data1 + data1.dtype.type(float(hdr1["PEDESTAL"]))
        if self._is_float(value):
            data2 = float(value)
            second_data = f"scalar {data2}"
            value_str = f"{data2}"
            self._logger.debug(f"Input second item can be cast to float: {data2:.3f}")
        else:
            if not Path(value).exists():
                raise ValueError(f"Error, {value} is not a scalar or a valid file path.")
            self._logger.debug(f"Input second item is an existing path: {value}")
            try:
                data2, hdr2, _ = self._read_fits(value, 0, to_float=False)
                if "PEDESTAL" in hdr2 and float(hdr2["PEDESTAL"]) != 0:
                    data2 = data2 + data2.dtype.type(float(hdr2["PEDESTAL"]))
                second_data = "array"
                if data1.shape != data2.shape:
                    raise RuntimeError("Error, the dimension of the second data array does not match the first."
                                       f" First image shape: {data1.shape}, second image shape: {data2.shape}")
            except Exception as exc:                    # noqa: BLE001  (the reference's bare except, :300-302)
                raise ValueError(f"Error, {value} is not a valid FITS file.") from exc
            value_str = Path(value).name
        result = self.apply(data1, operation, data2)
        verb = {"ADD": f"Added {second_data} to input image", "SUB": f"Subtracted {second_data} from input image",
                "MUL": f"Multiplied input image by {second_data}", "DIV": f"Divided input image by {second_data}"}[operation]
        self._logger.info(verb)
        ohistory_str = f"{self._name} input1 operation input2 are: {Path(inp_img).name} {operation} {value_str}"
        odict = {} if units is None else {"BUNIT": (units, "Pixel value units")}
        from .. import fitsio
        path = self._check_file_exists(inp_img)
        hdr = self._header_like(fitsio.read_header(path, 0), odict, f"Applied {self._name}", drop_scaling=False)
        hdr["HISTORY"] = ohistory_str
        fitsio.write_image(out_img, result, hdr, overwrite=True)
        self._logger.info(f"Wrote modified data to {out_img}")
        self._logger.debug("File processing completed.")
