"""CPU tests: the numpy oracle against (a) the committed goldens minted from the
reference source executed verbatim, (b) the live reference when /root/reference is
present, (c) hand-derived known answers for the (unpinned) combine stage."""
import json
import os

import numpy as np
import pytest

from conftest import bits_equal
from oracle import badpix_oracle as bo
from oracle import calibrate_oracle as co
from oracle import combine_oracle as C
from oracle import ref_exec


# ---------------------------------------------------------------- calibrate
def test_calibrate_oracle_matches_reference_goldens(golden_dir):
    z = np.load(os.path.join(golden_dir, "calibrate_small.npz"))
    meta = json.loads(str(z["meta_json"]))
    assert len(meta) == 5
    for (name, rawk, darkk, ped, iexp, dexp, useflat, usemask, dp, sb, nbad, nfix, nrem) in meta:
        raw = co.read_convert(z[rawk], ped)
        nf = None
        if useflat:
            nf = co.normalise_flat(z["flat"])
            assert bits_equal(nf, z[f"normflat_{name}"]), name
        cal = co.calibrate(raw, z["bias"], z[darkk], co.exptime_ratio({"EXPTIME": iexp}, {"EXPOSURE": dexp}), nf, bool(sb))
        if usemask:
            for fix in (bo.fix_bad_pixels_loop, bo.fix_bad_pixels_vec):
                out, st = fix(cal, z["mask"].astype(np.float32), dp)
                assert bits_equal(out, z[f"out_{name}"]), (name, fix.__name__)
                assert (st["BPIXNBAD"][0], st["BPIXNFIX"][0], st["BPIXNREM"][0]) == (nbad, nfix, nrem)
        else:
            assert bits_equal(cal, z[f"out_{name}"]), name


def test_exptime_ratio_errors():
    with pytest.raises(RuntimeError, match="both image and dark"):
        co.exptime_ratio({}, {})
    with pytest.raises(RuntimeError, match="for image"):
        co.exptime_ratio({}, {"EXPTIME": 1})
    with pytest.raises(RuntimeError, match="for dark"):
        co.exptime_ratio({"EXPOSURE": 1}, {})
    assert co.exptime_ratio({"EXPOSURE": 300, "EXPTIME": 1}, {"EXPTIME": 900}) == 300 / 900


def test_numpy_pairwise_restatement_is_numpy():
    """The summation tree the CUDA flat-norm kernel reproduces IS numpy's (np.sum, np.nanmean)."""
    for n in [1, 7, 8, 9, 127, 128, 129, 1000, 4097, 12288, 70001]:
        a = np.random.default_rng(n).normal(30000, 300, n).astype(np.float32)
        assert co.numpy_pairwise_sum_f32(a) == np.sum(a), n
    a = np.random.default_rng(1).normal(30000, 3000, 9999).astype(np.float32)
    a[[5, 77, 4000]] = np.nan
    z = np.where(np.isnan(a), np.float32(0), a)
    tot = co.numpy_pairwise_sum_f32(z)
    assert np.float32(np.float64(tot) / np.float64(np.sum(~np.isnan(a)))) == np.nanmean(a)


# ---------------------------------------------------------------- bad pixels
def test_badpix_oracle_matches_reference_goldens(golden_dir):
    z = np.load(os.path.join(golden_dir, "badpix_cases.npz"))
    assert len(z["names"]) == 9
    for name in z["names"]:
        data, mask, exp = z[f"data_{name}"], z[f"mask_{name}"], z[f"out_{name}"]
        dp, nbad, nfix, nrem = [int(v) for v in z[f"stat_{name}"]]
        for fix in (bo.fix_bad_pixels_loop, bo.fix_bad_pixels_vec):
            out, st = fix(data, mask, dp)
            assert bits_equal(out, exp), (name, fix.__name__)
            assert (st["BPIXNBAD"][0], st["BPIXNFIX"][0], st["BPIXNREM"][0]) == (nbad, nfix, nrem), name
            assert st["BPIXCORR"][0] == (nfix > 0) and st["BPIX_MIN"][0] == 4 and st["BPIXDPIX"][0] == dp


def test_badpix_hand_cases():
    d = np.arange(25, dtype=np.float32).reshape(5, 5)
    m = np.zeros((5, 5), np.uint8)
    m[0, 0] = 1                              # corner: donors 1, 5, 6 -> only 3 good: not fixable at dp=1
    m[2, 2] = 2                              # centre: 8 donors -> median of {6,7,8,11,13,16,17,18} = 12
    out, st = bo.fix_bad_pixels_loop(d, m, 1)
    assert out[0, 0] == 0 and out[2, 2] == 12.0
    assert (st["BPIXNBAD"][0], st["BPIXNFIX"][0], st["BPIXNREM"][0]) == (2, 1, 1)
    out2, _ = bo.fix_bad_pixels_loop(d, m, 2)   # dp=2: corner window 3x3 minus itself = 8 donors, (2,2) is bad -> 7
    assert out2[0, 0] == np.median([1, 2, 5, 6, 7, 10, 11]).astype(np.float32)


def test_mask_rules_match_reference_goldens(golden_dir):
    z = np.load(os.path.join(golden_dir, "findbadpix.npz"))
    from astrophotography_b200 import synth
    rules = synth.USER_BADPIX_EXAMPLE
    auto, nauto, _ = bo.auto_mask(z["dark"], 4.0)
    assert np.array_equal(auto, z["mask_auto"]) and nauto == int(z["counts"][0])
    full, nuser = bo.user_mask(auto.shape, rules["bad_columns"], [], rules["bad_rectangles"], mask=auto.copy())
    assert np.array_equal(full, z["mask"]) and nuser == int(z["counts"][1])
    assert full.max() >= 3 and full[0, 0] == 2 + auto[0, 0]          # overlaps accumulate; 1-based corner rule
    auto_s, nauto_s, _ = bo.auto_mask(z["dark_small"], 4.0)
    small, nuser_s = bo.user_mask(auto_s.shape, rules["bad_columns"], [], rules["bad_rectangles"], mask=auto_s.copy())
    assert np.array_equal(small, z["mask_small"])                       # out-of-range rules skipped
    assert (nauto_s, nuser_s) == tuple(int(v) for v in z["counts_small"])


# ---------------------------------------------------------------- live reference
needs_ref = pytest.mark.skipif(not ref_exec.reference_available(), reason="/root/reference not present")


@needs_ref
def test_oracle_vs_live_reference_calibrate_and_repair():
    from astrophotography_b200 import synth
    shape = (60, 90)
    rng = np.random.default_rng(4)
    raw = synth.science_frame(shape, seed=21, nstars=4)
    bias = rng.normal(1000, 10, shape).astype(np.float32)
    dark = rng.normal(1100, 15, shape).astype(np.float32)
    flat = synth.flat_frame(shape, seed=8)
    mask = synth.badpix_mask(shape, seed=14, auto_fraction=0.01)
    out, hdr, nf = ref_exec.ref_calibrate(raw, {"EXPTIME": 200.0, "PEDESTAL": -64}, bias, dark, {"EXPTIME": 600.0},
                                          flat, mask, 2, True)
    nf_o = co.normalise_flat(flat)
    cal = co.calibrate(co.read_convert(raw, -64), bias, dark, 200.0 / 600.0, nf_o, True)
    fixed, st = bo.fix_bad_pixels_vec(cal, mask.astype(np.float32), 2)
    assert bits_equal(nf_o, nf) and bits_equal(fixed, out)
    assert hdr["BPIXNFIX"] == st["BPIXNFIX"][0] and hdr["BIASCORR"] is True and hdr["BUNIT"] == "adu"


@needs_ref
def test_oracle_vs_live_reference_fix_dtypes():
    rng = np.random.default_rng(6)
    for dtype in (np.float32, np.float64, np.uint16):
        data = rng.integers(0, 4000, (40, 50)).astype(dtype)
        mask = (rng.random((40, 50)) < 0.08).astype(np.uint8)
        ref, rst = ref_exec.ref_fix_bad_pixels(data, mask, 1)
        if np.issubdtype(dtype, np.floating):
            out, st = bo.fix_bad_pixels_vec(data, mask, 1)
            assert bits_equal(out, ref)


def test_imarith_goldens_are_the_live_reference(golden_dir):
    """tests/golden/imarith.npz was minted by running the reference's ApImArith.process_files verbatim; when the
    reference tree is present the same call must reproduce it, and plain numpy must agree (the arithmetic is
    four ufunc calls, core/ApImArith.py:321-333)."""
    import json
    z = np.load(os.path.join(golden_dir, "imarith.npz"))
    a = z["a"]
    for op, name, dtype, _b, _n in json.loads(str(z["meta_json"]))[:-1]:
        val = {"scalar": 3.3, "scalar0": 0.0, "b32": z["b32"], "b64": z["b64"]}[name]
        fn = {"ADD": np.add, "SUB": np.subtract, "MUL": np.multiply, "DIV": np.divide}[op]
        res = np.zeros(a.shape, dtype=a.dtype)
        with np.errstate(all="ignore"):
            fn(a, val, out=res)
        assert bits_equal(res, z[f"out_{op}_{name}"]), (op, name)
        if ref_exec.reference_available():
            out, _, _ = ref_exec.ref_imarith(a, op, val)
            assert bits_equal(out, z[f"out_{op}_{name}"])


# ---------------------------------------------------------------- combine (unpinned: KATs + numpy facts)
def test_combine_oracle_hand_kats(golden_dir):
    k = np.load(os.path.join(golden_dir, "combine_kat.npz"))
    a = C.combine(k["A_stack"], "average", 5, 5, 1, "median", "mad_std")
    assert np.allclose(a["data"][0], k["A_mean"], rtol=1e-14, atol=0, equal_nan=True)
    assert np.array_equal(a["nrej"][0], k["A_nrej"]) and np.array_equal(a["allmasked"][0], k["A_allmasked"])
    assert np.allclose(a["uncert"][0], k["A_uncert"], rtol=1e-14, atol=0, equal_nan=True)
    for m, key in (("median", "B_median"), ("min", "B_min"), ("max", "B_max"), ("average", "B_mean")):
        b = C.combine(k["B_stack"], m, maxiters=0)
        assert np.allclose(b["data"][0], k[key], rtol=1e-14, atol=0, equal_nan=True), m
        assert np.array_equal(b["nrej"][0], k["B_nrej"])
    c = C.combine(k["C_stack"], "average", 1.5, 1.5, 5, "mean", "std")
    assert np.allclose(c["data"][0], k["C_mean"]) and np.array_equal(c["nrej"][0], k["C_nrej"])
    d = C.combine(k["D_stack"], "average", 0.2, 3.0, 1, "median", "std")
    assert np.allclose(d["data"][0], k["D_mean"]) and np.array_equal(d["nrej"][0], k["D_nrej"])


def test_numpy_axis0_sum_is_sequential():
    """Relied on by the generic CUDA kernel: nanmean(axis=0) accumulates frames in order."""
    rng = np.random.default_rng(3)
    st = rng.normal(1000, 300, (37, 5, 11)).astype(np.float64)
    seq = np.zeros(st.shape[1:])
    for f in st:
        seq = seq + f
    assert np.array_equal(np.sum(st, axis=0), seq)
    assert np.array_equal(np.nanmean(st, axis=0), seq / 37)


def test_sigma_clip_properties():
    rng = np.random.default_rng(5)
    st = rng.normal(100, 5, (40, 6, 7)).astype(np.float32)
    st[3, 2, 2] = 1e4
    kept = C.sigma_clip_stack(st, 3, 3, 5, "median", "std")
    assert np.isnan(kept[3, 2, 2])
    again = C.sigma_clip_stack(np.where(np.isnan(kept), np.float32(np.nan), st), 3, 3, None, "median", "std")
    assert np.array_equal(np.isnan(again), np.isnan(kept))          # idempotent once converged
    none = C.combine(st, "average", maxiters=0)
    assert np.array_equal(none["data"], np.mean(st.astype(np.float64), axis=0)) and none["nrej"].sum() == 0
    mean, med, std = C.sigma_clipped_stats_global(st[0], 3.0)
    assert abs(med - 100) < 2 and 3 < std < 7


def _scalar_sigma_clip_pixel(x, k_lo, k_hi, maxiters, cen, dev):
    """astropy.stats.sigma_clip on one pixel's N samples, written out sample by sample with plain Python floats
    (a second, structurally different restatement: no masked arrays, no axis arithmetic)."""
    import math
    kept = [float(v) for v in x if math.isfinite(float(v))]           # sigma_clip masks non-finite input first

    def median(v):
        w = sorted(v)
        m = len(w)
        return w[m // 2] if m % 2 else (w[m // 2 - 1] + w[m // 2]) / 2.0

    it = 0
    while kept and (maxiters is None or it < maxiters):
        it += 1
        mean = math.fsum(kept) / len(kept)
        c = mean if cen == "mean" else median(kept)
        if dev == "std":
            s = math.sqrt(math.fsum((v - mean) ** 2 for v in kept) / len(kept))
        else:
            med = median(kept)
            s = 1.482602218505602 * median([abs(v - med) for v in kept])
        lo, hi = c - k_lo * s, c + k_hi * s
        new = [v for v in kept if not (v < lo or v > hi)]
        changed = len(new) != len(kept)
        kept = new
        if not changed:
            break
    return kept


@pytest.mark.parametrize("case", [(5.0, 5.0, 1, "median", "mad_std"), (3.0, 3.0, 5, "mean", "std"),
                                  (2.0, 3.5, None, "mean", "std"), (3.0, 3.0, 3, "median", "std"),
                                  (1.5, 4.0, 2, "mean", "mad_std")], ids=str)
def test_vectorised_oracle_equals_a_scalar_restatement(case):
    """The numpy oracle (masked, vectorised over the image) against the per-pixel scalar restatement above, on
    stacks with outliers, NaN / inf samples, constant and two-valued pixels: same survivors, same mean."""
    k_lo, k_hi, maxiters, cen, dev = case
    rng = np.random.default_rng(17)
    n, h, w = 23, 6, 9
    st = rng.normal(1000.0, 12.0, size=(n, h, w)).astype(np.float32)
    hits = rng.random((n, h, w)) < 0.03
    st[hits] += rng.uniform(100, 30000, size=int(hits.sum())).astype(np.float32)
    st[0, 0, 0] = np.nan
    st[5, 0, 1] = np.inf
    st[:, 0, 2] = np.nan
    st[:, 0, 3] = 7.0
    st[: n // 2, 0, 4] = 5.0
    st[n // 2:, 0, 4] = 6.0
    got = C.combine(st, "average", k_lo, k_hi, maxiters, cen, dev)
    for i in range(h):
        for j in range(w):
            kept = _scalar_sigma_clip_pixel(st[:, i, j], k_lo, k_hi, maxiters, cen, dev)
            assert got["nrej"][i, j] == n - len(kept), (i, j)
            if kept:
                assert abs(got["data"][i, j] - sum(kept) / len(kept)) <= 1e-12 * abs(got["data"][i, j]), (i, j)
            else:
                assert np.isnan(got["data"][i, j]) and got["allmasked"][i, j] == 1


def test_two_operation_orders_differ_only_at_ties(golden_dir):
    """``lo = c - k*s; x < lo`` (astropy, ccdproc >= 2.4) against
    ``x - c < -k*s`` (ccdproc <= 2.3).  The committed census says how often they disagree per
    configuration; every disagreement is a documented kappa-boundary tie."""
    import json
    from oracle import make_golden as G
    with open(os.path.join(golden_dir, "combine_order_census.json")) as f:
        census = json.load(f)
    assert set(census) == set(G.ORDER_CASES)
    assert census["tenths6_k1_mean_std"]["orders_differ"] > 0        # the orders are really different
    for name in ("dark30q_apmastercal", "tenths6_k1_mean_std", "tenths8_k1_medmad", "dark30q_astropy_default"):
        _gen, _n, _shape, _seed, k_lo, k_hi, maxiters, cen, dev = G.ORDER_CASES[name]
        got = C.order_census(G.order_case_stack(name), k_lo, k_hi, maxiters, cen, dev)
        assert got == census[name], name
        assert got["differences_are_ties"]


def test_against_real_ccdproc_if_present():
    ccdproc = pytest.importorskip("ccdproc")
    from astropy.nddata import CCDData
    from astropy.stats import mad_std
    rng = np.random.default_rng(9)
    st = rng.normal(1000, 12, (12, 16, 16)).astype(np.float32)
    st[2, 3, 3] += 5000
    ccds = [CCDData(f, unit="adu") for f in st]
    ref = ccdproc.combine(ccds, method="average", sigma_clip=True, sigma_clip_low_thresh=5, sigma_clip_high_thresh=5,
                          sigma_clip_func=np.ma.median, sigma_clip_dev_func=mad_std)
    mine = C.combine(st, "average", 5, 5, 1, "median", "mad_std")
    assert np.allclose(np.asarray(ref.data), mine["data"], rtol=1e-12)
