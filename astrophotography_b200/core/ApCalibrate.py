"""ApCalibrate: bias / scaled-dark / flat calibration and bad-pixel repair on the GPU.

Host-side mirror of ``AstroPhotography/core/ApCalibrate.py`` of the reference:
same constructor (:48-53), same ``calibrate(raw_image, cal_image, delta_pix,
norm_flat, fixcosmic)`` (:406-410), same output keywords (``BIASCORR, BIASFILE,
DARKCORR, DARKFILE, BUNIT, FLATCORR, FLATFILE, BPIXFILE, BPIX*`` :454-486), same
errors.  The masters are read once, converted to float32, uploaded and kept
resident in HBM for the life of the object; the normalised flat is produced on
the device (``apgpu_flat_norm_f32`` reproduces ``np.nanmean`` bit-for-bit).  Each
``calibrate`` call is then H2D of the raw frame (uint16 frames travel as 2
bytes/pixel, conversion and PEDESTAL fused into the kernel), one fused
``apgpu_calibrate_*`` launch, one ``apgpu_fix_badpix_f32`` launch, D2H.

Additive API (not in the reference): ``calibrate_array`` for in-memory frames --
the reference's own TODO (``ApCalibrate.py:3``).

Documented deviations: float64 masters (what ``ccdproc.combine`` writes) are
cast to float32 on load, so the output is float32 where numpy's promotion
would have made it float64; ``fixcosmic=True`` (L.A.Cosmic via
ccdproc/astroscrappy, ``:491-497``) is outside the hot path and raises
unless ``ccdproc`` is importable.
"""
from __future__ import annotations

import time
from pathlib import Path

import numpy as np

from .. import _native, kernels
from ._base import ApBase
from .ApFixBadPixels import ApFixBadPixels, _mask_to_device


class ApCalibrate(ApBase):
    """Astronomical CCD image calibrator: bias subtraction, dark subtraction,
    flat fielding and bad pixel correction, for one telescope/detector
    (and, through the flat, one filter)."""

    MEAN_FULL = 0      #: normalise the flat by the mean of the entire flat
    MEDIAN_FULL = 1    #: reserved by the reference, not implemented there either

    _name = "ApCalibrate"

    def __init__(self, master_bias_file, master_dark_file, master_flat_file,
                 master_badpix_file, loglevel, dark_still_biased=None):
        self._master_bias_file = master_bias_file
        self._master_dark_file = master_dark_file
        self._master_flat_file = master_flat_file
        self._master_badpix_file = master_badpix_file
        self._loglevel = loglevel
        self._initialize_logger(loglevel)
        self._master_bias = Path(master_bias_file)
        self._master_dark = Path(master_dark_file)
        self._dark_still_biased = bool(dark_still_biased) if dark_still_biased is not None else False
        self._torch = _native.require_cuda()
        self._device = self._torch.device("cuda", self._torch.cuda.current_device())

        ext_num = 0
        bias, self._bias_hdr, _ = self._read_fits(self._master_bias, ext_num)
        dark, self._dark_hdr, _ = self._read_fits(self._master_dark, ext_num)
        self._bias_dev = self._upload_f32(bias, "master bias")
        self._dark_dev = self._upload_f32(dark, "master dark")
        self._check_same_shape(self._dark_dev, "master dark")

        self._norm_flat_dev = None
        if master_flat_file is not None:
            self._master_flat = Path(master_flat_file)
            self._flat_method = ApCalibrate.MEAN_FULL
            self._logger.info(f"Reading master flat field {self._master_flat.name}")
            flat, _, _ = self._read_fits(self._master_flat, ext_num)
            self._norm_flat_dev = self._generate_flat(self._upload_f32(flat, "master flat"), self._flat_method)
            self._check_same_shape(self._norm_flat_dev, "master flat")

        self._bpix = None
        self._mask_dev = None
        if master_badpix_file is not None:
            self._bpix = ApFixBadPixels(loglevel)
            self._master_bpix = Path(master_badpix_file)
            mask, self._mskhdr, _ = self._read_fits(self._master_bpix, ext_num, to_float=False)
            self._mask_dev = _mask_to_device(self._torch, mask, self._device)
            self._check_same_shape(self._mask_dev, "master bad pixel mask")

    # -- helpers ------------------------------------------------------------
    def _upload_f32(self, arr, what):
        if arr.dtype != np.float32:
            self._logger.info(f"Casting {what} from {arr.dtype} to float32 for the GPU path.")
            arr = arr.astype(np.float32)
        return self._torch.from_numpy(np.ascontiguousarray(arr)).to(self._device)

    def _check_same_shape(self, t, what):
        if tuple(t.shape) != tuple(self._bias_dev.shape):
            msg = (f"Error, the shape of the {what} ({tuple(t.shape)}) does not match that of"
                   f" the master bias ({tuple(self._bias_dev.shape)}).")
            self._logger.error(msg)
            raise RuntimeError(msg)

    def _find_exptime_ratio(self, img_hdr, dark_hdr):
        """Image-to-dark exposure ratio from EXPOSURE or EXPTIME (seconds)."""
        found = {}
        for label, hdr in (("image", img_hdr), ("dark", dark_hdr)):
            for kw in ("EXPOSURE", "EXPTIME"):
                if kw in hdr:
                    found[label] = float(hdr[kw])
                    self._logger.debug(f"{label.capitalize()} exposure time [seconds]: {found[label]:.2f}")
                    break
        msg = None
        if "image" not in found and "dark" not in found:
            msg = "Could not determine exposure time for both image and dark."
        elif "image" not in found:
            msg = "Could not determine exposure time for image (dark exposure found)."
        elif "dark" not in found:
            msg = "Could not determine exposure time for dark (img exposure found)."
        if msg is not None:
            self._logger.error(msg)
            raise RuntimeError(msg)
        exp_ratio = found["image"] / found["dark"]
        self._logger.info(f"Image to dark exposure time ratio: {exp_ratio:.3f}")
        return exp_ratio

    def _generate_flat(self, flat_dev, flat_method):
        if flat_method != ApCalibrate.MEAN_FULL:
            msg = f"Error, flat field normalization method {flat_method} has not been implemented yet."
            self._logger.error(msg)
            raise RuntimeError(msg)
        self._logger.debug("Using mean value of input field image to normalize by.")
        norm_flat, norm = kernels.flat_normalise(flat_dev)
        self._flat_norm_factor = float(norm.item())
        self._logger.info(f"Flat field normalization factor: {self._flat_norm_factor:.2f}")
        return norm_flat

    def _get_gain(self, hdr):
        gain = None
        for kw in ("GAIN", "EGAIN"):
            if kw in hdr:
                gain = float(hdr[kw])
        if gain is None:
            gain = 1.0
            self._logger.warning(f"Could not find gain value in header. Assuming gain={gain:.3f} e/ADU.")
        return gain

    @property
    def norm_flat(self):
        """The normalised flat as a numpy array (or None)."""
        return None if self._norm_flat_dev is None else self._norm_flat_dev.cpu().numpy()

    # -- array level (additive) --------------------------------------------
    def calibrate_array(self, raw_data, raw_hdr, delta_pix=2, pedestal=0.0, return_device=False):
        """Calibrate one in-memory frame.

        ``raw_data``: numpy uint16 / float32 (H,W) array, or a CUDA tensor of
        those dtypes.  ``pedestal`` is applied to uint16 input inside the
        kernel.  Returns ``(calibrated, odict)``."""
        torch = self._torch
        if isinstance(raw_data, torch.Tensor):
            raw_dev = raw_data
        else:
            raw_data = np.ascontiguousarray(raw_data)
            if raw_data.dtype == np.uint16:
                raw_dev = torch.from_numpy(raw_data.view(np.int16)).to(self._device, non_blocking=True).view(torch.uint16)
            else:
                if raw_data.dtype != np.float32:
                    raw_data = raw_data.astype(np.float32)
                    if pedestal:
                        raw_data += np.float32(pedestal)
                        pedestal = 0.0
                raw_dev = torch.from_numpy(raw_data).to(self._device, non_blocking=True)
        if tuple(raw_dev.shape) != tuple(self._bias_dev.shape):
            msg = (f"Error, the shape of the raw image ({tuple(raw_dev.shape)}) does not match that of"
                   f" the master bias ({tuple(self._bias_dev.shape)}).")
            self._logger.error(msg)
            raise RuntimeError(msg)
        if self._dark_still_biased:
            self._logger.info("Subtracting bias from dark")
        else:
            self._logger.debug("Dark assumed to already be bias-subtracted.")
        exp_ratio = self._find_exptime_ratio(raw_hdr, self._dark_hdr)
        if raw_dev.dtype == torch.float32 and pedestal:
            raise RuntimeError("calibrate_array: apply PEDESTAL to float32 frames before the call")
        img = kernels.calibrate(raw_dev, self._bias_dev, self._dark_dev, self._norm_flat_dev, exp_ratio,
                                self._dark_still_biased, pedestal=pedestal if raw_dev.dtype == torch.uint16 else None)
        odict = {"BIASCORR": (True, "True if bias subtracted."),
                 "BIASFILE": (self._master_bias.name, "Master bias file used."),
                 "DARKCORR": (True, "True if scaled dark subtracted."),
                 "DARKFILE": (self._master_dark.name, "Master dark file used."),
                 "BUNIT": ("adu", "Pixel value units.")}
        if self._norm_flat_dev is not None:
            odict["FLATCORR"] = (True, "True if flat field applied.")
            odict["FLATFILE"] = (self._master_flat.name, "Master flat file used.")
        else:
            self._logger.info("No flat field correction applied.")
        if self._bpix is not None:
            img, bpix_odict = self._bpix.fix_bad_pixels(img, self._mask_dev, delta_pix)
            odict["BPIXFILE"] = (self._master_bpix.name, "Name of master bad pixel file used")
            for key, val in bpix_odict.items():
                if "BPIX" in key:
                    odict[key] = val
        else:
            self._logger.info("No bad pixel correction applied.")
        if return_device:
            return img, odict
        return img.cpu().numpy(), odict

    def _base_odict(self):
        odict = {"BIASCORR": (True, "True if bias subtracted."),
                 "BIASFILE": (self._master_bias.name, "Master bias file used."),
                 "DARKCORR": (True, "True if scaled dark subtracted."),
                 "DARKFILE": (self._master_dark.name, "Master dark file used."),
                 "BUNIT": ("adu", "Pixel value units.")}
        if self._norm_flat_dev is not None:
            odict["FLATCORR"] = (True, "True if flat field applied.")
            odict["FLATFILE"] = (self._master_flat.name, "Master flat file used.")
        return odict

    # -- batch driver (additive; replaces scripts/calibrate_all.sh:353-480) --------------------------------
    def calibrate_many(self, raw_images, cal_images, delta_pix=2, ring=3):
        """Calibrate many frames with ONE set of device-resident masters.

        The reference's batch driver starts one Python process per frame (``calibrate_all.sh:406-411``), each
        re-reading the four masters.  Here frame ``k+1`` is read from its file into page-locked memory (raw
        BITPIX=16 data units undecoded: 2 bytes per pixel over PCIe) and uploaded on a copy stream while frame
        ``k`` runs through ONE fused calibrate + repair launch and frame ``k-1`` returns on a third stream --
        already in FITS byte order, so writing the output is the header plus one ``write`` of the pinned buffer.
        Output files equal those of ``calibrate(raw, cal, delta_pix, None, False)`` pixel for pixel.
        Returns the list of per-frame keyword dictionaries."""
        from .. import fitsio, pipeline
        torch = self._torch
        raw_images = [Path(p) for p in raw_images]
        cal_images = [Path(p) for p in cal_images]
        if len(raw_images) != len(cal_images):
            raise RuntimeError("calibrate_many: one output name per raw frame is needed")
        shape = tuple(self._bias_dev.shape)
        h, w = shape
        mask_u8 = None
        if self._bpix is not None:
            mask_u8 = self._mask_dev if self._mask_dev.dtype == torch.uint8 else (self._mask_dev != 0).to(torch.uint8)
        if self._dark_still_biased:
            self._logger.info("Subtracting bias from dark")
        copy_s, comp_s, back_s = (torch.cuda.Stream(device=self._device) for _ in range(3))
        slots = []
        for _ in range(ring):
            u16_np, u16_t = pipeline.pinned_empty(shape, np.uint16)
            f32_np, f32_t = pipeline.pinned_empty(shape, np.float32)
            out_np, out_t = pipeline.pinned_empty(shape, np.float32)
            slots.append(dict(u16_np=u16_np, u16_t=u16_t.view(torch.int16), f32_np=f32_np, f32_t=f32_t, out_np=out_np, out_t=out_t,
                              raw16=torch.empty(shape, dtype=torch.int16, device=self._device),
                              raw32=torch.empty(shape, dtype=torch.float32, device=self._device),
                              out=torch.empty(shape, dtype=torch.float32, device=self._device),
                              counts=torch.zeros(2, dtype=torch.int64, device=self._device),
                              counts_host=torch.zeros(2, dtype=torch.int64).pin_memory(),
                              done=None, pending=None, consumed=None))
        results = [None] * len(raw_images)

        def finish(slot):
            if slot["pending"] is None:
                return
            idx, hdr, odict, t0 = slot["pending"]
            slot["done"].synchronize()
            if mask_u8 is not None:
                nbad, nfixed = (int(v) for v in slot["counts_host"].tolist())
                odict["BPIXFILE"] = (self._master_bpix.name, "Name of master bad pixel file used")
                for key, val in self._bpix._stats_dict(h * w, nbad, nfixed, int(delta_pix)).items():
                    if "BPIX" in key:
                        odict[key] = val
            out_hdr = self._header_like(hdr, odict, f"Processed by {self._name}")
            fitsio.write_data_unit(cal_images[idx], slot["out_np"], -32, shape, out_hdr)
            results[idx] = odict
            self._logger.info(f"Calibrated {raw_images[idx].name} in {time.perf_counter() - t0:.3f} seconds.")
            slot["pending"] = None

        for idx, raw_path in enumerate(raw_images):
            slot = slots[idx % ring]
            finish(slot)                                   # the frame that used this slot before is written out
            t0 = time.perf_counter()
            self._check_file_exists(raw_path)
            lay = fitsio.ImageLayout(raw_path)
            if tuple(lay.shape) != shape:
                msg = (f"Error, the shape of the raw image ({tuple(lay.shape)}) does not match that of"
                       f" the master bias ({shape}).")
                self._logger.error(msg)
                raise RuntimeError(msg)
            hdr = lay.header
            pedestal = float(hdr["PEDESTAL"]) if "PEDESTAL" in hdr else 0.0
            exp_ratio = self._find_exptime_ratio(hdr, self._dark_hdr)
            if slot["consumed"] is not None:
                slot["consumed"].synchronize()             # the staging buffers' last upload is done
            if lay.raw_u16:
                lay.read_rows_raw(0, h, slot["u16_np"])
                kind, host_t, dev_raw, ped = "u16_fits", slot["u16_t"], slot["raw16"], pedestal
            else:
                data, _ = fitsio.read_image(raw_path, 0)
                np.copyto(slot["f32_np"], data, casting="unsafe")        # reference: astype(float32)
                if pedestal != 0:
                    slot["f32_np"] += np.float32(pedestal)
                kind, host_t, dev_raw, ped = "f32", slot["f32_t"], slot["raw32"], 0.0
            with torch.cuda.stream(copy_s):
                if slot["done"] is not None:
                    copy_s.wait_event(slot["done"])        # the kernel that last read this device buffer is done
                dev_raw.copy_(host_t, non_blocking=True)
                up = torch.cuda.Event()
                up.record(copy_s)
                slot["consumed"] = up
            with torch.cuda.stream(comp_s):
                comp_s.wait_event(up)
                slot["counts"].zero_()
                kernels.calibrate_repair(dev_raw, self._bias_dev, self._dark_dev, self._norm_flat_dev, exp_ratio,
                                         self._dark_still_biased, pedestal=ped, mask=mask_u8, deltapix=int(delta_pix),
                                         min_valid=self._bpix._min_valid if self._bpix is not None else 4,
                                         raw_kind=kind, out=slot["out"], out_big_endian=True, counts=slot["counts"])
                cev = torch.cuda.Event()
                cev.record(comp_s)
            with torch.cuda.stream(back_s):
                back_s.wait_event(cev)
                slot["out_t"].copy_(slot["out"], non_blocking=True)
                slot["counts_host"].copy_(slot["counts"], non_blocking=True)
                done = torch.cuda.Event()
                done.record(back_s)
                slot["done"] = done
            slot["pending"] = (idx, hdr, self._base_odict(), t0)
        for k in range(ring):
            finish(slots[(len(raw_images) + k) % ring])
        return results

    # -- file level (reference signature) ------------------------------------
    def calibrate(self, raw_image, cal_image, delta_pix, norm_flat, fixcosmic):
        perf_time_start = time.perf_counter()
        raw_image = Path(raw_image)
        ext_num = 0
        if not fixcosmic:
            # the common case runs through the batch driver's frame path (raw FITS data unit straight to the GPU,
            # one fused calibrate + repair call, FITS byte order produced on the device): same pixels, same keywords
            self.calibrate_many([raw_image], [cal_image], delta_pix, ring=1)
            if norm_flat is not None and self._norm_flat_dev is not None:
                self._logger.debug(f"Writing normalized flat field to {norm_flat}")
                self._write_image_like(raw_image, ext_num, norm_flat, self.norm_flat, {}, f"Processed by {self._name}")
            self._logger.info(f"Wrote bias/dark/flat corrected file to {cal_image}")
            return
        raw_data, raw_hdr, pedestal = self._read_fits(raw_image, ext_num, to_float=False)
        if raw_data.dtype != np.uint16:
            # anything but the common unsigned 16-bit frame takes the reference's
            # host conversion: -> float32, += PEDESTAL
            if not np.issubdtype(raw_data.dtype, np.floating):
                raw_data = raw_data.astype(np.float32)
            elif raw_data.dtype != np.float32:
                self._logger.info(f"Casting raw image from {raw_data.dtype} to float32 for the GPU path.")
                raw_data = raw_data.astype(np.float32)
            else:
                raw_data = np.array(raw_data, dtype=np.float32, copy=True)
            if pedestal != 0:
                self._logger.debug(f"Removing a PEDESTAL value of {pedestal} ADU.")
                raw_data += pedestal
            pedestal = 0.0
        img_bdf, odict = self.calibrate_array(raw_data, raw_hdr, delta_pix, pedestal)
        if norm_flat is not None and self._norm_flat_dev is not None:
            self._logger.debug(f"Writing normalized flat field to {norm_flat}")
            self._write_image_like(raw_image, ext_num, norm_flat, self.norm_flat, {},
                                   f"Processed by {self._name}")
        if fixcosmic:
            self._logger.info("Correcting cosmic rays...")
            img_bdf, cr_kw = self._fix_cosmic(img_bdf, self._get_gain(raw_hdr))
            odict.update(cr_kw)
        run_time_secs = time.perf_counter() - perf_time_start
        self._logger.info(f"Writing calibrated image to {cal_image}")
        self._write_image_like(raw_image, ext_num, cal_image, img_bdf, odict, f"Processed by {self._name}")
        self._logger.info(f"Wrote bias/dark/flat corrected file to {cal_image}")
        self._logger.info(f"Calibrated {raw_image.name} in {run_time_secs:.3f} seconds.")

    def _fix_cosmic(self, img, gain):
        try:
            import ccdproc      # noqa: F401  (un-vendored third party; not part of the GPU hot path)
        except Exception as exc:  # noqa: BLE001
            msg = ("fixcosmic=True needs ccdproc/astroscrappy (L.A.Cosmic); that step is outside"
                   " the GPU hot path and those packages are not installed.")
            self._logger.error(msg)
            raise RuntimeError(msg) from exc
        clean, crmask = ccdproc.cosmicray_lacosmic(img, gain=gain, gain_apply=False)   # pragma: no cover
        return np.asarray(clean, dtype=np.float32), {"CRCORR": (True, "True if cosmic rays removed")}  # pragma: no cover
