"""GPU tests of the Ap* class / ap_* CLI surface, file in -> file out, against the
verbatim-reference goldens and the numpy oracle."""
import json
import os

import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu


def _fits(path, data, **cards):
    from astrophotography_b200 import fitsio
    fitsio.write_image(path, data, fitsio.new_header(cards))
    return str(path)


def test_apcalibrate_files_match_reference_goldens(cuda, golden_dir, tmp_path):
    import astrophotography_b200 as ap
    from astrophotography_b200 import fitsio
    z = np.load(os.path.join(golden_dir, "calibrate_small.npz"))
    meta = json.loads(str(z["meta_json"]))
    bias = _fits(tmp_path / "mbias.fits", z["bias"])
    flat = _fits(tmp_path / "mflat.fits", z["flat"])
    mask = _fits(tmp_path / "mask.fits", z["mask"])
    for (name, rawk, darkk, ped, iexp, dexp, useflat, usemask, dp, sb, nbad, nfix, nrem) in meta:
        dark = _fits(tmp_path / f"mdark_{name}.fits", z[darkk], EXPTIME=dexp)
        cards = {"EXPTIME": iexp, "OBJECT": "M31"}
        if ped != 0:
            cards["PEDESTAL"] = ped
        raw = _fits(tmp_path / f"raw_{name}.fits", z[rawk], **cards)
        out = str(tmp_path / f"cal_{name}.fits")
        cal = ap.ApCalibrate(bias, dark, flat if useflat else None, mask if usemask else None, "ERROR", bool(sb))
        cal.calibrate(raw, out, dp, None, False)
        data, hdr = fitsio.read_image(out, 0)
        assert data.dtype == np.float32 and bits_equal(data, z[f"out_{name}"]), name
        assert hdr["BIASCORR"] is True and hdr["DARKCORR"] is True and hdr["BUNIT"] == "adu"
        assert hdr["BIASFILE"] == "mbias.fits" and hdr["DARKFILE"] == f"mdark_{name}.fits"
        assert "PEDESTAL" not in hdr and "BZERO" not in hdr and hdr["OBJECT"] == "M31"
        assert ("FLATCORR" in hdr) == bool(useflat)
        if usemask:
            assert (hdr["BPIXNBAD"], hdr["BPIXNFIX"], hdr["BPIXNREM"]) == (nbad, nfix, nrem)
            assert hdr["BPIXFILE"] == "mask.fits" and hdr["BPIXDPIX"] == dp and hdr["BPIX_MIN"] == 4
        assert any("Processed by ApCalibrate" in ln for ln in fitsio.header_history(hdr))


def test_apcalibrate_errors(cuda, tmp_path):
    import astrophotography_b200 as ap
    rng = np.random.default_rng(0)
    b = _fits(tmp_path / "b.fits", rng.normal(1000, 5, (16, 16)).astype(np.float32))
    d = _fits(tmp_path / "d.fits", rng.normal(10, 5, (16, 16)).astype(np.float32))          # no exposure keyword
    r = _fits(tmp_path / "r.fits", rng.integers(0, 4000, (16, 16)).astype(np.uint16), EXPTIME=10.0)
    with pytest.raises(RuntimeError, match="Cannot find"):
        ap.ApCalibrate(str(tmp_path / "nope.fits"), d, None, None, "ERROR")
    cal = ap.ApCalibrate(b, d, None, None, "ERROR")
    with pytest.raises(RuntimeError, match="exposure time for dark"):
        cal.calibrate(r, str(tmp_path / "o.fits"), 2, None, False)
    d2 = _fits(tmp_path / "d2.fits", rng.normal(10, 5, (16, 16)).astype(np.float32), EXPOSURE=20.0)
    with pytest.raises(RuntimeError, match="ccdproc"):
        ap.ApCalibrate(b, d2, None, None, "ERROR").calibrate(r, str(tmp_path / "o.fits"), 2, None, True)
    small = _fits(tmp_path / "small.fits", np.zeros((8, 16), np.float32), EXPOSURE=20.0)
    with pytest.raises(RuntimeError, match="does not match"):
        ap.ApCalibrate(b, small, None, None, "ERROR")
    with pytest.raises(ValueError, match="Invalid log level"):
        ap.ApCalibrate(b, d2, None, None, "NOISY")


def test_apfixbadpixels_arrays_and_files(cuda, golden_dir, tmp_path):
    import astrophotography_b200 as ap
    from astrophotography_b200 import fitsio
    z = np.load(os.path.join(golden_dir, "badpix_cases.npz"))
    fixer = ap.ApFixBadPixels("ERROR")
    for name in z["names"]:
        dp, nbad, nfix, nrem = [int(v) for v in z[f"stat_{name}"]]
        out, st = fixer.fix_bad_pixels(z[f"data_{name}"], z[f"mask_{name}"], dp)
        assert bits_equal(out, z[f"out_{name}"]), name
        assert (st["BPIXNBAD"][0], st["BPIXNFIX"][0], st["BPIXNREM"][0]) == (nbad, nfix, nrem)
        assert st["BPIXCORR"][0] == (nfix > 0) and st["numpix"][0] == out.size and st["BPIXDPIX"][0] == dp
    name = "yml_dp2"
    inp = _fits(tmp_path / "in.fits", z[f"data_{name}"], EXPTIME=1.0)
    msk = _fits(tmp_path / "msk.fits", z[f"mask_{name}"])
    fixer.fix_files(inp, msk, str(tmp_path / "fixed.fits"), 2)
    data, hdr = fitsio.read_image(tmp_path / "fixed.fits", 0)
    assert bits_equal(data, z[f"out_{name}"]) and hdr["BPIXFILE"] == "msk.fits" and hdr["BPIXCORR"] is True
    # uint16 image: medians are truncated into the integer array, like numpy's assignment
    from oracle import badpix_oracle as bo
    rng = np.random.default_rng(2)
    u = rng.integers(0, 4000, (30, 40)).astype(np.uint16)
    m = (rng.random((30, 40)) < 0.1).astype(np.uint8)
    got, _ = fixer.fix_bad_pixels(u, m, 1)
    expf, _ = bo.fix_bad_pixels_vec(u.astype(np.float32), m, 1)
    assert got.dtype == np.uint16 and np.array_equal(got, np.trunc(expf).astype(np.uint16))
    with pytest.raises(RuntimeError, match="does not match"):
        fixer.fix_bad_pixels(np.zeros((4, 4), np.float32), np.zeros((4, 5), np.uint8))


def test_apfindbadpixels_matches_reference_goldens(cuda, golden_dir, tmp_path):
    import astrophotography_b200 as ap
    from astrophotography_b200 import fitsio, synth
    import yaml
    z = np.load(os.path.join(golden_dir, "findbadpix.npz"))
    yml = tmp_path / "user.yml"
    yml.write_text(yaml.safe_dump(synth.USER_BADPIX_EXAMPLE))
    dark = _fits(tmp_path / "mdark.fits", z["dark"], TELESCOP="T5", **{"SET-TEMP": -20.0})
    f = ap.ApFindBadPixels(dark, 4.0, "ERROR")
    assert np.array_equal(f.get_mask(), z["mask_auto"]) and f._nbad_auto == int(z["counts"][0])
    f.add_user_badpix(str(yml))
    assert np.array_equal(f.get_mask(), z["mask"]) and f._nbad_user == int(z["counts"][1])
    f.write_mask(str(tmp_path / "badpix.fits"))
    m, hdr = fitsio.read_image(tmp_path / "badpix.fits", 0)
    assert m.dtype == np.uint8 and np.array_equal(m, z["mask"])
    assert hdr["IMAGETYP"] == "BADPIX" and hdr["NBADAUTO"] == int(z["counts"][0]) and hdr["NBADUSER"] == int(z["counts"][1])
    assert hdr["TELESCOP"] == "T5" and hdr["USERFILE"] == "user.yml"
    small = _fits(tmp_path / "small.fits", z["dark_small"])
    g = ap.ApFindBadPixels(small, 4.0, "ERROR")
    g.add_user_badpix(str(yml))
    assert np.array_equal(g.get_mask(), z["mask_small"])
    (tmp_path / "empty.yml").write_text("bad_columns: {}\n")
    g2 = ap.ApFindBadPixels(small, 4.0, "ERROR")
    g2.add_user_badpix(str(tmp_path / "empty.yml"))          # absent keys mean "none" (the reference raises TypeError)
    assert g2._nbad_user == 0


def test_sigma_clipped_stats_kernel(cuda):
    torch = cuda
    from astrophotography_b200 import kernels
    from oracle import combine_oracle as C
    rng = np.random.default_rng(12)
    for n, sig in [(1, 3.0), (2, 3.0), (1000, 3.0), (123457, 4.0), (1 << 20, 2.5)]:
        a = rng.normal(1000, 12, n).astype(np.float32)
        if n > 100:
            a[rng.integers(0, n, n // 200)] += 5000
            a[7] = np.nan
            a[9] = np.inf
        mean, med, std, cnt = kernels.sigma_clipped_stats(torch.from_numpy(a).cuda().view(1, n), sig)
        emean, emed, estd = C.sigma_clipped_stats_global(a, sig)
        assert med == emed, (n, med, emed)                                   # exact selection
        assert abs(mean - emean) <= 1e-12 * abs(emean) and abs(std - estd) <= 1e-10 * max(estd, 1e-30) + 1e-300, n


def test_apmastercal_make_master(cuda, tmp_path):
    from astrophotography_b200 import ApMasterCal, fitsio, synth
    from oracle import combine_oracle as C
    shape, n = (40, 64), 9
    st = synth.dark_stack(n, shape, exptime=300.0, quantise=True)
    d = tmp_path / "darks"
    d.mkdir()
    for k in range(n):
        hdr = fitsio.new_header({"IMAGETYP": "Dark Frame", "EXPTIME": 300.0, "SET-TEMP": -20.0, "CCD-TEMP": -20.1,
                                 "TELESCOP": "", "UT": "12:00", "SWOWNER": "x"})
        fitsio.write_image(d / f"d{k:02d}.fits", st[k].astype(np.uint16), hdr)
    mc = ApMasterCal(str(d), "master*", "iTelescope 5", 0.5, "ERROR")
    out = str(d / "master_dark.fits")
    mc.make_master(out)
    data, hdr = fitsio.read_image(out, 0)
    mask, mh = fitsio.read_image(out, 1)
    unc, uh = fitsio.read_image(out, 2)
    exp = C.combine(st, "average", 5.0, 5.0, 1, "median", "mad_std")
    assert data.dtype == np.float64 and np.allclose(data, exp["data"], rtol=1e-13, atol=0)
    assert np.array_equal(mask, exp["allmasked"]) and np.allclose(unc, exp["uncert"], rtol=1e-10, atol=1e-12)
    assert mh["EXTNAME"] == "MASK" and uh["EXTNAME"] == "UNCERT"
    assert hdr["IMAGETYP"] == "MASTER DARK" and hdr["TELESCOP"] == "iTelescope 5" and hdr["NCOMBINE"] == n
    assert hdr["CREATOR"] == "ApMasterCal" and hdr["IFILE000"] == "d00.fits" and "UT" not in hdr and "SWOWNER" not in hdr
    assert np.array_equal(mc._last_result["nrej"].astype(np.int64), exp["nrej"])
    # a second run must skip the master it just wrote (exclude pattern "master*")
    mc2 = ApMasterCal(str(d), "master*", "iTelescope 5", 0.5, "ERROR", method="median", sigma_clip=False, out_dtype="float32")
    assert len(mc2._files) == n
    mc2.make_master(str(tmp_path / "median.fits"))
    med, _ = fitsio.read_image(tmp_path / "median.fits", 0)
    assert med.dtype == np.float32 and np.array_equal(med, np.median(st.astype(np.float64), axis=0).astype(np.float32))


def _write_dark_files(d, n, shape, as_u16=True):
    from astrophotography_b200 import fitsio, synth
    st = synth.dark_stack(n, shape, exptime=300.0, quantise=True)
    d.mkdir()
    for k in range(n):
        hdr = fitsio.new_header({"IMAGETYP": "Dark Frame", "EXPTIME": 300.0, "SET-TEMP": -20.0, "CCD-TEMP": -20.0, "TELESCOP": "T5"})
        fitsio.write_image(d / f"d{k:03d}.fits", st[k].astype(np.uint16) if as_u16 else st[k], hdr)
    return st


@pytest.mark.parametrize("as_u16", [True, False])
def test_make_master_streams_files_and_shards_rows(cuda, tmp_path, as_u16):
    """make_master reads every file once into page-locked memory while the previous frames upload (raw BITPIX=16
    data units travel undecoded); ``gpus=2`` shards the rows over two worker processes (both on cuda:0 here) that
    write into one shared host array.  Same master either way, equal to the oracle."""
    from astrophotography_b200 import ApMasterCal, fitsio, pipeline
    from oracle import combine_oracle as C
    n, shape = 12, (45, 72)
    st = _write_dark_files(tmp_path / "darks", n, shape, as_u16)
    exp = C.combine(st, "average", 5.0, 5.0, 1, "median", "mad_std")
    mc = ApMasterCal(str(tmp_path / "darks"), "master*", "T5", 0.5, "ERROR")
    src = pipeline.FileFrames(mc._files.files_filtered(include_path=True), fitsio)
    assert src.raw_u16 == as_u16 and src.u16_format == ("fits" if as_u16 else "native")
    one = mc.combine_files()
    opts = dict(out_f64=True, want_nrej=True, want_uncert=True, want_allmasked=True, **mc._combine)
    two = pipeline.combine_files_sharded(src.paths, 2, devices=[0, 0], **opts)
    for res in (one, two):
        assert np.array_equal(res["nrej"].astype(np.int64), exp["nrej"])
        assert np.allclose(res["data"], exp["data"], rtol=1e-13, atol=0)
        assert np.allclose(res["uncert"], exp["uncert"], rtol=1e-10, atol=1e-12)
    assert np.array_equal(one["data"], two["data"])
    if torch_device_count() >= 2:
        mc.make_master(str(tmp_path / "m2.fits"), gpus=2)
        data, _ = fitsio.read_image(tmp_path / "m2.fits", 0)
        assert np.array_equal(data, one["data"])


def torch_device_count():
    import torch
    return torch.cuda.device_count()


def test_make_master_overlaps_file_reads_with_uploads(cuda, tmp_path):
    """30 x 4096^2 raw uint16 files (BASELINE config 2's master): streaming the files through combine_files must
    cost no more than 1.3 x (reading them + combining pre-pinned arrays); the numbers are printed."""
    import time
    import torch
    from astrophotography_b200 import fitsio, pipeline
    n, shape = 30, (4096, 4096)
    d = tmp_path / "big"
    d.mkdir()
    rng = np.random.default_rng(1)
    base = rng.normal(1000, 12, shape).astype(np.float32)
    for k in range(n):
        fr = np.clip(np.rint(base + rng.normal(0, 12, shape).astype(np.float32)), 0, 65535).astype(np.uint16)
        fitsio.write_image(d / f"b{k:02d}.fits", fr, fitsio.new_header({"IMAGETYP": "Bias Frame"}))
    paths = sorted(str(p) for p in d.iterdir())
    src = pipeline.FileFrames(paths, fitsio)
    assert src.raw_u16
    params = dict(method="average", k_lo=5.0, k_hi=5.0, maxiters=1, cen="median", dev="mad_std")
    pipeline.combine_files(src, **params)                                  # warm-up (page cache, CUDA context)
    t0 = time.perf_counter()
    res = pipeline.combine_files(src, **params)
    t_files = time.perf_counter() - t0
    pinned = [pipeline.pinned_empty(shape, np.uint16) for _ in range(n)]
    t0 = time.perf_counter()
    for i in range(n):
        src.read_band(i, 0, shape[0], pinned[i][0])
    t_read = time.perf_counter() - t0
    comb = pipeline.HostStackCombiner(n, shape[0], shape[1], dtype=np.uint16, u16_format="fits", **params)
    comb.combine([p[0] for p in pinned])
    t0 = time.perf_counter()
    ref = comb.combine([p[0] for p in pinned])
    torch.cuda.synchronize()
    t_comb = time.perf_counter() - t0
    print(f"combine_files {t_files * 1e3:.1f} ms; read only {t_read * 1e3:.1f} ms; pre-pinned HostStackCombiner {t_comb * 1e3:.1f} ms")
    assert np.array_equal(res["data"], ref["data"]) and np.array_equal(res["nrej"], ref["nrej"])
    assert t_files <= 1.3 * (t_read + t_comb)


def test_fix_bad_pixels_refuses_dtypes_it_cannot_repair_exactly(cuda):
    """float64 / 32-bit integer images: np.median's even-count mean is a float64 operation there, so the float32
    kernel refuses them instead of returning other last bits; 16-bit integers are exact and truncate like numpy."""
    import astrophotography_b200 as ap
    from oracle import badpix_oracle as bo
    rng = np.random.default_rng(2)
    mask = (rng.random((24, 30)) < 0.05).astype(np.uint8)
    fixer = ap.ApFixBadPixels("CRITICAL")
    for dt in (np.float64, np.int32, np.int64):
        with pytest.raises(RuntimeError, match="not supported on the GPU path"):
            fixer.fix_bad_pixels(rng.normal(1000, 30, (24, 30)).astype(dt), mask, 2)
    img = rng.integers(0, 60000, (24, 30)).astype(np.uint16)
    got, _ = fixer.fix_bad_pixels(img, mask, 2)
    # numpy assigns np.median's float64 result into the uint16 array by truncation (core/ApFixBadPixels.py:409);
    # for 16-bit values the float32 median is the same number
    expf, _ = bo.fix_bad_pixels_vec(img.astype(np.float32), mask, 2)
    assert got.dtype == np.uint16 and np.array_equal(got, np.trunc(expf).astype(np.uint16))


def test_calibrate_many_equals_per_frame_calibrate(cuda, tmp_path):
    """The batch driver (masters resident, three streams, fused calibrate + repair, FITS byte order produced on
    the GPU) writes the same pixels and keywords as one ApCalibrate.calibrate call per frame; the ap_calibrate_all
    CLI skips existing outputs like calibrate_all.sh."""
    import astrophotography_b200 as ap
    from astrophotography_b200 import fitsio, synth
    from astrophotography_b200.scripts import ap_calibrate_all
    shape = (48, 72)
    rng = np.random.default_rng(5)
    bias = _fits(tmp_path / "mbias.fits", rng.normal(1000, 5, shape).astype(np.float32))
    dark = _fits(tmp_path / "mdark.fits", rng.normal(1040, 6, shape).astype(np.float32), EXPTIME=900.0)
    flat = _fits(tmp_path / "mflat.fits", synth.flat_frame(shape))
    mask = _fits(tmp_path / "mask.fits", synth.badpix_mask(shape, auto_fraction=1e-2))
    rawdir = tmp_path / "raw"
    rawdir.mkdir()
    raws = []
    for k in range(7):
        fr = synth.science_frame(shape, seed=20 + k, as_uint16=(k != 3), nstars=3)
        cards = {"EXPTIME": 300.0 + 10 * k, "OBJECT": f"T{k}"}
        if k % 2:
            cards["PEDESTAL"] = -100
        raws.append(_fits(rawdir / f"raw{k}.fits", fr, **cards))
    cal = ap.ApCalibrate(bias, dark, flat, mask, "ERROR", True)
    outs = [str(tmp_path / f"many{k}.fits") for k in range(7)]
    odicts = cal.calibrate_many(raws, outs, 2)
    assert len(odicts) == 7 and all(o["BIASCORR"][0] for o in odicts)
    for k in range(7):
        one = str(tmp_path / f"one{k}.fits")
        cal.calibrate(raws[k], one, 2, None, False)
        a, ha = fitsio.read_image(outs[k], 0)
        b, hb = fitsio.read_image(one, 0)
        assert a.dtype == np.float32 and bits_equal(a, b), k
        # ... and the two-launch array path (apgpu_calibrate_* then apgpu_fix_badpix_f32) gives the same pixels
        rawd, rawh = fitsio.read_image(raws[k], 0)
        ped = float(rawh["PEDESTAL"]) if "PEDESTAL" in rawh else 0.0
        if rawd.dtype != np.uint16:
            rawd = rawd.astype(np.float32) + np.float32(ped)
            ped = 0.0
        two, _ = cal.calibrate_array(rawd, rawh, 2, ped)
        assert bits_equal(a, two), k
        for kw in ("BIASCORR", "BIASFILE", "DARKCORR", "DARKFILE", "BUNIT", "FLATCORR", "FLATFILE", "BPIXFILE", "BPIXNBAD",
                   "BPIXNFIX", "BPIXNREM", "BPIXDPIX", "BPIX_MIN", "BPIXCORR", "OBJECT", "EXPTIME"):
            assert ha[kw] == hb[kw], (k, kw)
        assert "PEDESTAL" not in ha and "BZERO" not in ha
        assert any("Processed by ApCalibrate" in ln for ln in fitsio.header_history(ha))
    outdir = tmp_path / "out"
    argv = [str(rawdir), bias, dark, str(outdir), "--master_flat", flat, "--master_badpix", mask, "--dark_still_biased", "-l", "ERROR"]
    assert ap_calibrate_all.main(argv) == 0
    for k in range(7):
        a, _ = fitsio.read_image(outdir / f"cal-raw{k}.fits", 0)
        b, _ = fitsio.read_image(outs[k], 0)
        assert bits_equal(a, b)
    stamp = os.path.getmtime(outdir / "cal-raw0.fits")
    assert ap_calibrate_all.main(argv) == 0 and os.path.getmtime(outdir / "cal-raw0.fits") == stamp      # skipped


def test_float64_masters_are_cast_and_stay_within_1e6_of_the_float64_arithmetic(cuda, tmp_path):
    """Documented deviation: masters written as float64 (what ccdproc.combine and ApMasterCal's default write)
    make numpy promote the reference's whole calibration to float64; here they are cast to float32 on load and
    the output is float32.  The result stays within float32 rounding of the float64 arithmetic: every operation
    contributes at most one ulp of its operands (raw, bias, scaled dark), divided by the flat."""
    import astrophotography_b200 as ap
    from astrophotography_b200 import fitsio, synth
    from oracle import calibrate_oracle as co
    shape = (40, 56)
    rng = np.random.default_rng(9)
    bias64 = rng.normal(1000, 5, shape)
    dark64 = rng.normal(1040, 6, shape)
    flat64 = synth.flat_frame(shape).astype(np.float64)
    raw = synth.science_frame(shape, nstars=3)
    b = _fits(tmp_path / "b64.fits", bias64)
    d = _fits(tmp_path / "d64.fits", dark64, EXPTIME=900.0)
    f = _fits(tmp_path / "f64.fits", flat64)
    r = _fits(tmp_path / "raw.fits", raw, EXPTIME=300.0)
    ap.ApCalibrate(b, d, f, None, "ERROR", True).calibrate(r, str(tmp_path / "cal.fits"), 2, None, False)
    got, _ = fitsio.read_image(tmp_path / "cal.fits", 0)
    assert got.dtype == np.float32
    nf64 = co.normalise_flat(flat64)
    exp = co.calibrate(co.read_convert(raw), bias64, dark64, 300.0 / 900.0, nf64, True)
    assert exp.dtype == np.float64                                     # what the reference would have written
    ok = np.isfinite(exp) & np.isfinite(got) & (nf64 != 0)
    scale = (np.abs(raw.astype(np.float64)) + 2 * np.abs(bias64) + np.abs(dark64)) / np.abs(np.where(ok, nf64, 1.0))
    err = np.abs(got.astype(np.float64) - exp)
    assert (err[ok] <= 4 * 2.0 ** -24 * scale[ok] + 1e-6 * np.abs(exp[ok])).all()
    assert np.array_equal(np.isnan(got), np.isnan(exp))


def test_master_keeps_pedestal_and_apcalibrate_removes_it(cuda, tmp_path):
    """MaximDL-style frames carry PEDESTAL=-100.  Like ccdproc.combine the master keeps the keyword (the
    frames are combined as stored) and ApCalibrate._read_fits removes the pedestal from the master bias /
    dark exactly as from the raw frame (reference core/ApCalibrate.py:318-326)."""
    import astrophotography_b200 as ap
    from astrophotography_b200 import ApMasterCal, fitsio, synth
    from oracle import calibrate_oracle as co, combine_oracle as C
    shape, n, ped = (32, 48), 7, -100
    masters = {}
    for kind, exptime, typ in (("bias", 0.0, "Bias Frame"), ("dark", 600.0, "Dark Frame")):
        dd = tmp_path / kind
        dd.mkdir()
        st = synth.dark_stack(n, shape, exptime=exptime, quantise=True)
        for k in range(n):
            fitsio.write_image(dd / f"{kind}{k}.fits", st[k].astype(np.uint16),
                               fitsio.new_header({"IMAGETYP": typ, "EXPTIME": exptime, "SET-TEMP": -20.0, "CCD-TEMP": -20.0,
                                                  "TELESCOP": "T5", "PEDESTAL": ped}))
        out = str(tmp_path / f"master_{kind}.fits")
        ApMasterCal(str(dd), "master*", "T5", 0.5, "ERROR", out_dtype="float32").make_master(out)
        data, hdr = fitsio.read_image(out, 0)
        assert hdr["PEDESTAL"] == ped                                   # kept, not applied
        exp = C.combine(st, "average", 5, 5, 1, "median", "mad_std")["data"].astype(np.float32)
        assert np.array_equal(data, exp)
        masters[kind] = (out, exp)
    raw = synth.science_frame(shape, nstars=3)
    rawf = str(tmp_path / "raw.fits")
    fitsio.write_image(rawf, raw, fitsio.new_header({"EXPTIME": 300.0, "PEDESTAL": ped}))
    cal = ap.ApCalibrate(masters["bias"][0], masters["dark"][0], None, None, "ERROR", True)
    cal.calibrate(rawf, str(tmp_path / "cal.fits"), 2, None, False)
    got, h = fitsio.read_image(tmp_path / "cal.fits", 0)
    bias = masters["bias"][1] + np.float32(ped)
    dark = masters["dark"][1] + np.float32(ped)
    exp = co.calibrate(co.read_convert(raw, float(ped)), bias, dark, 300.0 / 600.0, None, True)
    assert bits_equal(got, exp) and "PEDESTAL" not in h


def test_cli_mains_end_to_end(cuda, golden_dir, tmp_path):
    """ap_combine_darks -> ap_find_badpix -> ap_calibrate -> ap_fix_badpix on synthetic files."""
    from astrophotography_b200 import fitsio, synth
    from astrophotography_b200.scripts import ap_calibrate, ap_combine_darks, ap_find_badpix, ap_fix_badpix
    from oracle import badpix_oracle as bo, calibrate_oracle as co, combine_oracle as C
    shape, n = (48, 72), 8
    dirs = {}
    for kind, exptime, typ in (("bias", 0.0, "Bias Frame"), ("dark", 900.0, "Dark Frame")):
        dd = tmp_path / kind
        dd.mkdir()
        st = synth.dark_stack(n, shape, exptime=exptime, quantise=True)
        for k in range(n):
            fitsio.write_image(dd / f"{kind}{k}.fits", st[k].astype(np.uint16),
                               fitsio.new_header({"IMAGETYP": typ, "EXPTIME": exptime, "SET-TEMP": -20.0, "CCD-TEMP": -20.0,
                                                  "TELESCOP": "T5"}))
        dirs[kind] = (dd, st)
        assert ap_combine_darks.main([str(dd), str(tmp_path / f"master_{kind}.fits"), "-l", "ERROR", "--out_dtype", "float32"]) == 0
    assert ap_combine_darks.main([str(tmp_path / "nodir"), str(tmp_path / "x.fits"), "-l", "CRITICAL"]) == 1
    mb, _ = fitsio.read_image(tmp_path / "master_bias.fits", 0)
    md, mdh = fitsio.read_image(tmp_path / "master_dark.fits", 0)
    expb = C.combine(dirs["bias"][1], "average", 5, 5, 1, "median", "mad_std")["data"].astype(np.float32)
    assert np.array_equal(mb, expb) and mdh["EXPTIME"] == 900.0
    assert ap_find_badpix.main([str(tmp_path / "master_dark.fits"), str(tmp_path / "badpix.fits"), "--sigma", "4", "-l", "ERROR"]) == 0
    mask, _ = fitsio.read_image(tmp_path / "badpix.fits", 0)
    assert np.array_equal(mask, bo.auto_mask(md, 4.0)[0])
    raw = synth.science_frame(shape, nstars=4)
    fitsio.write_image(tmp_path / "raw.fits", raw, fitsio.new_header({"EXPTIME": 300.0, "PEDESTAL": -100}))
    flat = synth.flat_frame(shape)
    fitsio.write_image(tmp_path / "flat.fits", flat, fitsio.new_header({"IMAGETYP": "Flat"}))
    assert ap_calibrate.main([str(tmp_path / "raw.fits"), str(tmp_path / "master_bias.fits"), str(tmp_path / "master_dark.fits"),
                              str(tmp_path / "cal.fits"), "--master_flat", str(tmp_path / "flat.fits"),
                              "--master_badpix", str(tmp_path / "badpix.fits"), "--dark_still_biased",
                              "--normflat", str(tmp_path / "nf.fits"), "-l", "ERROR"]) == 0
    cal, ch = fitsio.read_image(tmp_path / "cal.fits", 0)
    nf = co.normalise_flat(flat)
    e = co.calibrate(co.read_convert(raw, -100), mb, md, 300.0 / 900.0, nf, True)
    e, st = bo.fix_bad_pixels_vec(e, mask.astype(np.float32), 2)
    assert bits_equal(cal, e) and ch["BPIXNFIX"] == st["BPIXNFIX"][0]
    assert bits_equal(fitsio.read_image(tmp_path / "nf.fits", 0)[0], nf)
    assert ap_fix_badpix.main([str(tmp_path / "raw.fits"), str(tmp_path / "badpix.fits"), str(tmp_path / "rawfix.fits"),
                               "--deltapix", "1", "-l", "ERROR"]) == 0
    rf, _ = fitsio.read_image(tmp_path / "rawfix.fits", 0)
    assert rf.shape == shape


def test_host_stack_combiner_bands(cuda):
    """The double-buffered row-band pipeline equals the single-shot device result."""
    torch = cuda
    from astrophotography_b200 import kernels, pipeline
    rng = np.random.default_rng(3)
    n, h, w = 12, 37, 50
    frames = [rng.normal(1000, 12, (h, w)).astype(np.float32) for _ in range(n)]
    frames[2][3, 3] = 9e4
    comb = pipeline.HostStackCombiner(n, h, w, band_bytes=n * w * 4 * 5, want_uncert=True, want_allmasked=True)   # 5-row bands
    assert comb.nbands == 8
    res = comb.combine(frames)
    one = kernels.stack_reduce(torch.from_numpy(np.stack(frames)).cuda(), want_uncert=True, want_allmasked=True)
    for k in ("data", "nrej", "uncert", "allmasked"):
        assert np.array_equal(res[k], one[k].cpu().numpy(), equal_nan=True), k
    res2 = comb.combine(frames)                   # reusable
    assert np.array_equal(res2["data"], res["data"])
    with pytest.raises(RuntimeError):
        comb.combine(frames[:-1])
