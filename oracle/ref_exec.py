"""Execute the reference's own numpy code VERBATIM (test infrastructure only).

The reference (``/root/reference``, AstroPhotography 0.5.1) cannot be imported
as a package here: ``astropy``, ``ccdproc`` and friends are absent.  But the
arithmetic of

* ``AstroPhotography/core/ApCalibrate.py``      (``calibrate`` :406-509),
* ``AstroPhotography/core/ApFixBadPixels.py``   (``fix_bad_pixels`` :292-445),
* ``AstroPhotography/core/ApFindBadPixels.py``  (mask rules :70-217, :414-438)

is plain numpy.  This module loads those three source files *from where they
lie* with ``importlib`` under a private stub package, behind

* a fake ``astropy.io.fits`` backed by an in-memory ``{path: [HDU, ...]}``
  store (``open`` / ``writeto`` / ``PrimaryHDU`` / ``HDUList``),
* a stub ``astropy.stats.sigma_clipped_stats`` (numpy restatement, see
  ``combine_oracle.sigma_clipped_stats_global``),
* a stub ``ApFixCosmicRays`` (never called: ``fixcosmic=False`` on this path).

No reference source is copied into this repository.  ``/root/reference`` does
not exist on the GPU box, so nothing here is reachable from ``-m gpu`` tests,
``smoke()`` or ``bench.py``; it is used by ``oracle/make_golden.py`` (run in the
authoring container) to mint ``tests/golden/*.npz`` and by CPU tests that skip
when the reference is absent.
"""
from __future__ import annotations

import contextlib
import copy
import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("AP_REFERENCE_ROOT", "/root/reference")
_PKG = "_apref_verbatim"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(
        REFERENCE_ROOT, "AstroPhotography", "core", "ApCalibrate.py"))


# --------------------------------------------------------------------------
# fake astropy.io.fits
# --------------------------------------------------------------------------
class _Comments:
    def __init__(self, hdr):
        self._hdr = hdr

    def __getitem__(self, kw):
        return self._hdr._comments.get(kw, "")


class FakeHeader:
    """Just enough of astropy.io.fits.Header for the three reference classes."""

    def __init__(self, cards=None):
        self._d = {}
        self._comments = {}
        self.history = []
        for k, v in (cards or {}).items():
            self[k] = v

    def __contains__(self, kw):
        return kw in self._d

    def __getitem__(self, kw):
        return self._d[kw]

    def __setitem__(self, kw, val):
        if kw == "HISTORY":
            self.history.append(val)
            return
        if isinstance(val, tuple):
            self._d[kw] = val[0]
            if len(val) > 1:
                self._comments[kw] = val[1]
        else:
            self._d[kw] = val

    def __delitem__(self, kw):
        del self._d[kw]
        self._comments.pop(kw, None)

    def keys(self):
        return self._d.keys()

    def items(self):
        return self._d.items()

    @property
    def comments(self):
        return _Comments(self)

    def copy(self):
        return copy.deepcopy(self)


class FakeHDU:
    def __init__(self, data=None, header=None):
        self.data = data
        self.header = header if header is not None else FakeHeader()


class FakeHDUList(list):
    def writeto(self, name, output_verify=None, overwrite=False):
        FITS_STORE[str(name)] = FakeHDUList(
            FakeHDU(None if h.data is None else np.array(h.data, copy=True),
                    h.header.copy()) for h in self)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


FITS_STORE: dict[str, FakeHDUList] = {}


def _fake_open(name, uint=True, do_not_scale_image_data=False, **kw):
    src = FITS_STORE[str(name)]
    # A fresh copy per open(), like re-reading a file: the reference mutates
    # the array it has just read (``ext_data += pedestal``, ApCalibrate.py:322).
    return FakeHDUList(
        FakeHDU(None if h.data is None else np.array(h.data, copy=True),
                h.header.copy()) for h in src)


def put_image(name, data, header=None):
    """Register an in-memory 'FITS file' (primary HDU only)."""
    hdr = FakeHeader(header or {})
    data = np.asarray(data)
    hdr._d.setdefault("NAXIS", data.ndim)
    hdr._d.setdefault("NAXIS1", data.shape[-1])
    hdr._d.setdefault("NAXIS2", data.shape[-2])
    bitpix = {np.dtype("uint8"): 8, np.dtype("int16"): 16, np.dtype("uint16"): 16,
              np.dtype("int32"): 32, np.dtype("float32"): -32,
              np.dtype("float64"): -64}.get(data.dtype, -32)
    hdr._d.setdefault("BITPIX", bitpix)
    FITS_STORE[str(name)] = FakeHDUList([FakeHDU(data, hdr)])
    return str(name)


def get_image(name):
    hdu = FITS_STORE[str(name)][0]
    return hdu.data, hdu.header


# --------------------------------------------------------------------------
# loader
# --------------------------------------------------------------------------
_loaded = {}


@contextlib.contextmanager
def _stubbed_modules():
    """Temporarily install the fake third-party modules in sys.modules."""
    from . import combine_oracle

    fits_mod = types.ModuleType("astropy.io.fits")
    fits_mod.open = _fake_open
    fits_mod.PrimaryHDU = lambda data=None, header=None: FakeHDU(data, header)
    fits_mod.HDUList = FakeHDUList
    io_mod = types.ModuleType("astropy.io")
    io_mod.fits = fits_mod
    stats_mod = types.ModuleType("astropy.stats")
    stats_mod.sigma_clipped_stats = combine_oracle.sigma_clipped_stats_global
    astropy_mod = types.ModuleType("astropy")
    astropy_mod.io = io_mod
    astropy_mod.stats = stats_mod
    fakes = {"astropy": astropy_mod, "astropy.io": io_mod,
             "astropy.io.fits": fits_mod, "astropy.stats": stats_mod}
    saved = {k: sys.modules.get(k) for k in fakes}
    sys.modules.update(fakes)
    try:
        yield
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _load():
    if _loaded:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    core_dir = os.path.join(REFERENCE_ROOT, "AstroPhotography", "core")

    top = types.ModuleType(_PKG)
    top.__path__ = []
    top.__version__ = "0.5.1"          # AstroPhotography/__version__.py:19
    core = types.ModuleType(_PKG + ".core")
    core.__path__ = []
    top.core = core
    sys.modules[_PKG] = top
    sys.modules[_PKG + ".core"] = core

    # ApFixCosmicRays imports ccdproc at module import; it is never called on
    # this path (fixcosmic=False), so a stub class stands in for it.
    crmod = types.ModuleType(_PKG + ".core.ApFixCosmicRays")

    class ApFixCosmicRays:                       # noqa: D401 - stub
        def __init__(self, loglevel):
            pass

        def process(self, *a, **k):
            raise RuntimeError("cosmic-ray removal is outside the oracle's scope")
    crmod.ApFixCosmicRays = ApFixCosmicRays
    sys.modules[_PKG + ".core.ApFixCosmicRays"] = crmod

    def load(modname):
        path = os.path.join(core_dir, modname + ".py")
        spec = importlib.util.spec_from_file_location(
            f"{_PKG}.core.{modname}", path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        return mod

    with _stubbed_modules():
        fixmod = load("ApFixBadPixels")
        # core/__init__.py:19 re-exports the class under the module's name,
        # which is what ``from . import ApFixBadPixels`` in ApCalibrate.py:23 sees.
        core.ApFixBadPixels = fixmod.ApFixBadPixels
        findmod = load("ApFindBadPixels")
        core.ApFindBadPixels = findmod.ApFindBadPixels
        calmod = load("ApCalibrate")

    for cls in (fixmod.ApFixBadPixels, calmod.ApCalibrate):
        cls._check_file_exists = lambda self, filename: None
    findmod.ApFindBadPixels._check_file_exists = (
        lambda self, filename: filename if str(filename) in FITS_STORE
        else __import__("pathlib").Path(filename).expanduser())
    # ApFindBadPixels._read_fits refers to ``sys`` without importing it
    # (core/ApFindBadPixels.py:302) -- only on the 3-D error path.
    with _stubbed_modules():
        arithmod = load("ApImArith")
    arithmod.ApImArith._check_file_exists = lambda self, filename: None
    _loaded.update(ApFixBadPixels=fixmod.ApFixBadPixels,
                   ApFindBadPixels=findmod.ApFindBadPixels,
                   ApCalibrate=calmod.ApCalibrate,
                   ApImArith=arithmod.ApImArith)
    return _loaded


def ref_fix_bad_pixels(data, badpixmask, deltapix=1, loglevel="ERROR"):
    """Reference ``ApFixBadPixels.fix_bad_pixels`` executed verbatim."""
    cls = _load()["ApFixBadPixels"]
    with np.errstate(all="ignore"):
        newdata, stats = cls(loglevel).fix_bad_pixels(data, badpixmask, deltapix)
    return newdata, stats


def ref_calibrate(raw, raw_hdr, bias, dark, dark_hdr, flat=None, mask=None,
                  delta_pix=2, dark_still_biased=False, bias_hdr=None,
                  loglevel="ERROR"):
    """Reference ``ApCalibrate(...).calibrate(...)`` executed verbatim.

    Arrays are registered as in-memory FITS files; returns ``(calibrated
    ndarray, output header dict, normalised flat or None)``.
    """
    cls = _load()["ApCalibrate"]
    FITS_STORE.clear()
    put_image("bias.fits", bias, bias_hdr)
    put_image("dark.fits", dark, dark_hdr)
    put_image("raw.fits", raw, raw_hdr)
    flat_name = put_image("flat.fits", flat) if flat is not None else None
    mask_name = put_image("mask.fits", mask) if mask is not None else None
    with np.errstate(all="ignore"):
        cal = cls("bias.fits", "dark.fits", flat_name, mask_name, loglevel,
                  dark_still_biased)
        cal.calibrate("raw.fits", "out.fits", delta_pix, None, False)
    out, hdr = get_image("out.fits")
    normflat = getattr(cal, "_norm_flat", None)
    return out, dict(hdr.items()), normflat


def ref_find_bad_pixels(dark, sigma, user_yaml_path=None, dark_hdr=None,
                        loglevel="ERROR"):
    """Reference ``ApFindBadPixels`` (+ ``add_user_badpix``) executed verbatim."""
    cls = _load()["ApFindBadPixels"]
    FITS_STORE.clear()
    put_image("mdark.fits", dark, dark_hdr)
    with _stubbed_modules(), np.errstate(all="ignore"):
        obj = cls("mdark.fits", sigma, loglevel)
        if user_yaml_path is not None:
            obj.add_user_badpix(user_yaml_path)
    return obj.get_mask().copy(), obj._nbad_auto, obj._nbad_user


def ref_imarith(data1, operation, value, hdr1=None, units=None, loglevel="ERROR"):
    """Reference ``ApImArith.process_files`` (core/ApImArith.py:255-346) executed verbatim.  ``value`` is a
    number (passed as its string, like the CLI does) or a second image (registered as an in-memory FITS file
    behind a real temporary path, because the reference tests ``Path(value).exists()``).  Returns
    ``(result ndarray, output header dict, history list)``."""
    import tempfile
    cls = _load()["ApImArith"]
    FITS_STORE.clear()
    put_image("in.fits", data1, hdr1)
    tmp = None
    if isinstance(value, np.ndarray):
        tmp = tempfile.NamedTemporaryFile(suffix=".fits")
        put_image(tmp.name, value)
        value = tmp.name
    else:
        value = repr(float(value))
    try:
        with np.errstate(all="ignore"):
            cls(loglevel).process_files("in.fits", operation, value, "out.fits", units)
    finally:
        if tmp is not None:
            tmp.close()
    out, hdr = get_image("out.fits")
    return out, dict(hdr.items()), list(hdr.history)
