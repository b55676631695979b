"""CPU tests of the host side: FITS seam, C-ABI exports, no-GPU behaviour, ApMasterCal
file checks, CLI surface, row-band partitioning."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

from astrophotography_b200 import _native, fitsio, pipeline
from astrophotography_b200.core.ApMasterCal import ApMasterCal


# ---------------------------------------------------------------- C ABI
def _header_functions():
    text = open(os.path.join(ROOT, "include", "apgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(apgpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads on a CPU-only box and exports everything include/apgpu.h declares."""
    lib = _native.load()
    names = _header_functions()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), f"{name} declared in apgpu.h but not exported"
        assert name in _native.SIGNATURES, f"{name} has no ctypes signature in _native.py"
    assert lib.apgpu_abi_version() == 2
    assert lib.apgpu_last_error() is not None
    assert int(lib.apgpu_flat_norm_workspace_bytes(4096 * 4096)) > 64
    assert int(lib.apgpu_image_stats_workspace_bytes(100)) > 8192


def test_kernel_selection_is_host_logic():
    from astrophotography_b200 import kernels
    assert kernels.stack_kernel_name(100).startswith("sorted_medmad1<100>")
    assert kernels.stack_kernel_name(100, "average", 3, 3, 5, "mean", "std") == "meanclip<100>"
    assert kernels.stack_kernel_name(100, "average", 3, 3, 5, "mean", "std", prefer="shared") == "meanclip_smem"
    assert kernels.stack_kernel_name(300, "average", 3, 3, 5, "mean", "std") == "meanclip_coop<4>"
    assert kernels.stack_kernel_name(400, "average", 3, 3, 5, "mean", "std") == "meanclip_coop<8>"
    assert kernels.stack_kernel_name(150, "average", 3, 3, 5, "mean", "std") == "meanclip_coop<2>"
    assert kernels.stack_kernel_name(300, "average", 3, 3, 5, "mean", "std", prefer="shared") == "meanclip_smem"
    assert kernels.stack_kernel_name(200, "average", 3, 3, 5, "mean", "std") == "meanclip_coop<4>"
    assert kernels.stack_kernel_name(1000, "average", 3, 3, 5, "mean", "std") == "meanclip_split<8>"
    assert kernels.stack_kernel_name(30, "median", maxiters=0) == "sorted_median<32>"
    assert kernels.stack_kernel_name(30, "median", maxiters=0, want_uncert=True) == "sorted_median_mad<32>"
    assert kernels.stack_kernel_name(300, "median", maxiters=0) == "median_coop<8>"
    assert kernels.stack_kernel_name(256, "median", maxiters=0) == "median_coop<4>"
    assert kernels.stack_kernel_name(300, "median", maxiters=0, want_uncert=True) == "median_mad_coop<8>"
    assert kernels.stack_kernel_name(250) == "medmad1_coop<4>" and kernels.stack_kernel_name(513).startswith("generic")
    assert kernels.stack_kernel_name(300, "average", 3, 3, 5, "median", "std").startswith("generic")
    assert kernels.stack_kernel_name(600, "median", maxiters=0).startswith("generic")
    assert kernels.stack_kernel_name(30, force_generic=True).startswith("generic")
    assert kernels.stack_kernel_name(600, "average", 3, 3, 5, "mean", "std") == "meanclip_split<8>"
    assert kernels.stack_kernel_name(600, "average", 3, 3, 5, "mean", "std", prefer="registers").startswith("generic")


def test_argument_errors_surface_without_gpu():
    """Validation happens before any CUDA call: status + apgpu_last_error()."""
    lib = _native.load()
    st = lib.apgpu_calibrate_f32(None, None, None, None, 1.0, 0, None, 16, None)
    assert st == 1 and b"null" in lib.apgpu_last_error()
    with pytest.raises(RuntimeError, match="null"):
        _native.check(st, "calibrate")
    st = lib.apgpu_fix_badpix_f32(ctypes.c_void_p(8), ctypes.c_void_p(8), 0, 10, 10, 0, 10, 0, 10, 9, 4,
                                  ctypes.c_void_p(8), ctypes.c_void_p(8), None)
    assert st == 1 and b"deltapix" in lib.apgpu_last_error()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from astrophotography_b200 import kernels
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        kernels.stack_reduce(np.zeros((3, 4, 4), np.float32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        kernels.calibrate(None, None, None)
    import astrophotography_b200 as ap
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ap.ApFixBadPixels("ERROR").fix_bad_pixels(np.zeros((4, 4), np.float32), np.zeros((4, 4), np.uint8))


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "astrophotography_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


# ---------------------------------------------------------------- FITS seam
@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.uint16, np.int32, np.float32, np.float64])
def test_fits_roundtrip(tmp_path, dtype):
    rng = np.random.default_rng(1)
    data = (rng.random((13, 29)) * 60000).astype(dtype)
    hdr = fitsio.new_header({"EXPTIME": (300.0, "seconds"), "IMAGETYP": "Dark Frame", "PEDESTAL": -100,
                             "SET-TEMP": -20.0, "FLAG": True, "OBSERVER": "O'Neil"})
    hdr["HISTORY"] = "made by test"
    path = tmp_path / "a.fits"
    fitsio.write_image(path, data, hdr, extensions=[("MASK", (data > 100).astype(np.uint8)), ("UNCERT", data.astype(np.float64))])
    assert os.path.getsize(path) % 2880 == 0
    back, h2 = fitsio.read_image(path, 0)
    assert back.dtype == np.dtype(dtype) and np.array_equal(back, data)
    assert h2["EXPTIME"] == 300.0 and h2["IMAGETYP"] == "Dark Frame" and h2["PEDESTAL"] == -100
    assert h2["FLAG"] is True and h2["OBSERVER"] == "O'Neil" and h2["NAXIS1"] == 29 and h2["NAXIS2"] == 13
    assert h2.comments["EXPTIME"] == "seconds"
    m, hm = fitsio.read_image(path, 1)
    u, _ = fitsio.read_image(path, 2)
    assert m.dtype == np.uint8 and hm["EXTNAME"] == "MASK" and u.dtype == np.float64 and np.array_equal(u, data.astype(np.float64))
    assert fitsio.read_header(path, 0)["SET-TEMP"] == -20.0


def test_fits_long_strings_comments_and_nonfinite(tmp_path):
    """Cards the reference's outputs really carry: DATAFILE / BIASFILE / IFILEnnn hold long paths, COMMENT
    cards must survive a header copy, NaN / inf values must not produce invalid tokens."""
    long_path = "/data/iTelescope/T05/2024-01-30/calibration/" + "very_long_directory_name/" * 3 + "master_dark_-20C_900s.fits"
    quoted = "it's " * 20
    hdr = fitsio.new_header({"DATAFILE": (long_path, "Data file used to generate mask"), "OBSERVER": quoted,
                             "SHORT": ("abc", "a comment that is far too long to fit on the card " * 3),
                             "BADVAL": float("nan"), "INFVAL": float("inf"), "AFTER": 7})
    hdr["COMMENT"] = "FITS (Flexible Image Transport System) format"
    hdr["HISTORY"] = "x" * 150
    path = tmp_path / "long.fits"
    fitsio.write_image(path, np.zeros((3, 4), np.float32), hdr)
    raw = open(path, "rb").read()
    assert len(raw) % 2880 == 0
    head = raw[: raw.index(b"END" + b" " * 77) + 80].decode("ascii")
    assert len(head) % 80 == 0
    for i in range(0, len(head), 80):                       # every string card is closed
        card = head[i:i + 80]
        if card[:8].strip() in ("DATAFILE", "OBSERVER", "CONTINUE", "SHORT"):
            val = card[10:] if card[8:10] == "= " else card[8:]
            assert fitsio._split_value_comment(val)[0].strip().endswith("'"), card
    _, h2 = fitsio.read_image(path, 0)
    assert h2["DATAFILE"] == long_path and h2["OBSERVER"] == quoted.rstrip() and h2["SHORT"] == "abc" and h2["AFTER"] == 7
    assert h2["BADVAL"] == "NAN" and h2["INFVAL"] == "INF"
    assert h2.comment_cards == ["FITS (Flexible Image Transport System) format"]
    assert "".join(fitsio.header_history(h2)) == "x" * 150
    c = h2.copy()
    assert c.comment_cards == h2.comment_cards


def _fake_astropy_modules():
    """A stand-in for ``astropy.io.fits`` with the properties that matter at the seam: big-endian data arrays,
    a Header without ``.history`` whose ``['HISTORY']`` is a list, HDU objects and ``HDUList.writeto``."""
    import types

    class FakeHeader(dict):
        def __init__(self):
            super().__init__()
            self._com, self._hist = {}, []
        def __setitem__(self, k, v):
            k = k.upper()
            if k == "HISTORY":
                self._hist.append(str(v))
                return
            if isinstance(v, tuple):
                self._com[k] = v[1] if len(v) > 1 else ""
                v = v[0]
            super().__setitem__(k, v)
        def __getitem__(self, k):
            return list(self._hist) if k.upper() == "HISTORY" else super().__getitem__(k.upper())
        def __contains__(self, k):
            return (k.upper() == "HISTORY" and bool(self._hist)) or super().__contains__(k.upper())
        def __delitem__(self, k):
            super().__delitem__(k.upper())
        def get(self, k, d=None):
            return self[k] if k in self else d
        @property
        def comments(self):
            return {k: self._com.get(k, "") for k in self.keys()}
        def copy(self):
            h = FakeHeader()
            for k in self.keys():
                h[k] = (dict.__getitem__(self, k), self._com.get(k, ""))
            h._hist = list(self._hist)
            return h

    def to_fake(h):
        f = FakeHeader()
        for k in h.keys():
            f[k] = (h[k], h.comments[k])
        for ln in h.history:
            f["HISTORY"] = ln
        return f

    def from_fake(f):
        h = fitsio.Header()
        for k in f.keys():
            h[k] = (dict.__getitem__(f, k), f._com.get(k, ""))
        for ln in f._hist:
            h["HISTORY"] = ln
        return h

    class HDU:
        def __init__(self, data=None, header=None, name=None):
            self.data, self.header, self.name = data, header if header is not None else FakeHeader(), name
    class HDUList(list):
        def writeto(self, path, output_verify="exception", overwrite=False):
            blob = b""
            for i, hdu in enumerate(self):
                blob += fitsio._encode_hdu(hdu.data, from_fake(hdu.header), i == 0, extname=hdu.name)
            with open(path, "wb") as f:
                f.write(blob)
        def __enter__(self):
            return self
        def __exit__(self, *a):
            return False

    def fits_open(path, uint=False, do_not_scale_image_data=False):
        out = HDUList()
        ext = 0
        while True:
            try:
                data, hdr = fitsio._mini_read(str(path), ext)
            except (OSError, KeyError):
                break
            if data is not None and data.dtype.itemsize > 1:
                data = data.astype(data.dtype.newbyteorder(">"))     # astropy hands out big-endian arrays
            out.append(HDU(data, to_fake(hdr)))
            ext += 1
        return out

    fits = types.ModuleType("astropy.io.fits")
    fits.open, fits.Header, fits.PrimaryHDU, fits.ImageHDU, fits.HDUList = fits_open, FakeHeader, HDU, HDU, HDUList
    fits.getheader = lambda path, ext=0: fits_open(path)[ext].header
    io = types.ModuleType("astropy.io")
    io.fits = fits
    top = types.ModuleType("astropy")
    top.io = io
    return {"astropy": top, "astropy.io": io, "astropy.io.fits": fits}


def test_astropy_branch_of_the_seam_with_a_shim(tmp_path, monkeypatch):
    """The ``HAVE_ASTROPY`` branches of fitsio (never reachable offline otherwise): native byte order out of
    ``read_image``, HISTORY through ``header_history``, keyword round trip, extensions."""
    import importlib
    for name, mod in _fake_astropy_modules().items():
        monkeypatch.setitem(sys.modules, name, mod)
    try:
        importlib.reload(fitsio)
        assert fitsio.HAVE_ASTROPY
        data = (np.arange(12 * 20).reshape(12, 20) * 3).astype(np.uint16)
        hdr = fitsio.new_header({"EXPTIME": (300.0, "seconds"), "PEDESTAL": -100})
        hdr["HISTORY"] = "made by the shim test"
        path = tmp_path / "shim.fits"
        fitsio.write_image(path, data, hdr, extensions=[("MASK", (data > 100).astype(np.int16)), ("UNCERT", data.astype(np.float64))])
        back, h2 = fitsio.read_image(path, 0)
        assert back.dtype == np.uint16 and back.dtype.isnative and np.array_equal(back, data)
        assert h2["EXPTIME"] == 300.0 and h2["PEDESTAL"] == -100 and not hasattr(h2, "history")
        assert fitsio.header_history(h2) == ["made by the shim test"]
        m, _ = fitsio.read_image(path, 1)
        u, _ = fitsio.read_image(path, 2)
        assert m.dtype == np.int16 and m.dtype.isnative and u.dtype == np.float64 and u.dtype.isnative
        import torch
        torch.from_numpy(m)                      # what _mask_to_device does: must not raise on byte order
        torch.from_numpy(u)
        assert fitsio.read_header(path, 0)["EXPTIME"] == 300.0
        hc = h2.copy()
        del hc["PEDESTAL"]
        hc["HISTORY"] = "second"
        assert "PEDESTAL" not in hc and "PEDESTAL" in h2 and len(fitsio.header_history(hc)) == 2
    finally:
        monkeypatch.undo()
        importlib.reload(fitsio)
    assert not fitsio.HAVE_ASTROPY or "astropy" in sys.modules


def test_header_mapping():
    h = fitsio.Header({"A": 1})
    h["b"] = (2, "two")
    assert "B" in h and h["b"] == 2 and h.comments["B"] == "two"
    h["B"] = 3                                  # keeps the comment
    assert h.comments["B"] == "two" and dict(h.items())["B"] == 3
    del h["A"]
    assert "A" not in h and len(h) == 1
    c = h.copy()
    c["B"] = 9
    assert h["B"] == 3


# ---------------------------------------------------------------- sharding
@pytest.mark.parametrize("nrows,world", [(6388, 1), (6388, 2), (6388, 8), (7, 8), (100, 3)])
def test_row_band_partition(nrows, world):
    covered = []
    for r in range(world):
        r0, r1, h0, h1 = pipeline.row_band(nrows, world, r, halo=2)
        assert 0 <= h0 <= r0 <= r1 <= h1 <= nrows
        assert h0 == max(0, r0 - 2) and h1 == min(nrows, r1 + 2)
        covered += list(range(r0, r1))
    assert covered == list(range(nrows))
    sizes = [pipeline.row_band(nrows, world, r)[1] - pipeline.row_band(nrows, world, r)[0] for r in range(world)]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        pipeline.row_band(10, 2, 2)


# ---------------------------------------------------------------- ApMasterCal host logic
def _write_frames(d, n=3, imagetyp="Dark Frame", exptime=300.0, settemp=-20.0, temps=None, shape=(8, 10), telescop="T5"):
    rng = np.random.default_rng(0)
    for k in range(n):
        hdr = fitsio.new_header({"IMAGETYP": imagetyp, "EXPTIME": exptime, "SET-TEMP": settemp,
                                 "CCD-TEMP": (temps[k] if temps else settemp), "TELESCOP": telescop,
                                 "DATE-OBS": f"2020-01-0{k + 1}", "UT": "x", "SWOWNER": "me"})
        fitsio.write_image(os.path.join(d, f"dark{k}.fits"), rng.integers(900, 1100, shape).astype(np.uint16), hdr)


def test_mastercal_file_checks(tmp_path):
    d = tmp_path / "ok"
    d.mkdir()
    _write_frames(str(d), 4, temps=[-20.0, -20.3, -19.8, -25.0])
    fitsio.write_image(d / "master_old.fits", np.zeros((8, 10), np.float32), fitsio.new_header({"IMAGETYP": "x"}))
    mc = ApMasterCal(str(d), "master*", "UNKNOWN", 0.5, "ERROR")
    assert mc._files.values("file") == ["dark0.fits", "dark1.fits", "dark2.fits"]     # -25 C excluded, master* skipped
    kw = mc._generate_final_keywords()
    assert kw["IMAGETYP"][0] == "MASTER DARK" and kw["TELESCOP"][0] == "T5" and kw["SET-TEMP"][0] == -20.0
    assert kw["IFILE002"] == "dark2.fits" and kw["CREATOR"][0] == "ApMasterCal"

    bad = tmp_path / "mixed"
    bad.mkdir()
    _write_frames(str(bad), 2)
    fitsio.write_image(bad / "bias9.fits", np.zeros((8, 10), np.uint16),
                       fitsio.new_header({"IMAGETYP": "Bias Frame", "EXPTIME": 0.0, "SET-TEMP": -20.0, "CCD-TEMP": -20.0,
                                          "TELESCOP": "T5"}))
    with pytest.raises(RuntimeError, match="unique values of imagetyp"):
        ApMasterCal(str(bad), "master*", "UNKNOWN", 0.5, "ERROR")
    with pytest.raises(RuntimeError):
        ApMasterCal(str(tmp_path / "missing"), "master*", "UNKNOWN", 0.5, "ERROR")
    with pytest.raises(ValueError, match="Invalid log level"):
        ApMasterCal(str(d), "master*", "UNKNOWN", 0.5, "LOUD")


# ---------------------------------------------------------------- CLI surface
@pytest.mark.parametrize("script,needle", [
    ("ap_calibrate", "--dark_still_biased"), ("ap_fix_badpix", "--deltapix"), ("ap_calibrate_all", "--gpus"), ("ap_imarith", "--units"),
    ("ap_find_badpix", "--user_badpix"), ("ap_combine_darks", "--temptol")])
def test_cli_help(script, needle):
    r = subprocess.run([sys.executable, "-m", f"astrophotography_b200.scripts.{script}", "--help"],
                       capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0 and needle in r.stdout and f"usage: {script}" in r.stdout


def test_cli_defaults_match_reference():
    from astrophotography_b200.scripts import ap_calibrate, ap_combine_darks, ap_find_badpix, ap_fix_badpix
    a = ap_calibrate.command_line_opts(["r.fits", "b.fits", "d.fits", "o.fits"])
    assert a.deltapix == 2 and a.master_flat is None and a.fixcosmic is False and a.loglevel == "INFO"
    assert ap_fix_badpix.command_line_opts(["i", "m", "o"]).deltapix == 2
    assert ap_find_badpix.command_line_opts(["d", "o"]).sigma == 4.0
    c = ap_combine_darks.command_line_opts(["dir", "out.fits"])
    assert (c.temptol, c.telescop, c.exclude_pattern) == (0.5, "UNKNOWN", "master*")
    assert (c.method, c.kappa_low, c.kappa_high, c.maxiters, c.cenfunc, c.devfunc) == ("average", 5.0, 5.0, 1, "median", "mad_std")
