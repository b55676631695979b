"""Mint ``tests/golden/*.npz`` (run in the authoring container only).

    python -m oracle.make_golden

* ``calibrate_*.npz`` and ``badpix_*.npz`` and ``findbadpix.npz`` hold inputs
  and the outputs of the REFERENCE SOURCE EXECUTED VERBATIM from
  ``/root/reference`` (``oracle/ref_exec.py``): they pin the numpy restatements
  in ``oracle/calibrate_oracle.py`` / ``oracle/badpix_oracle.py`` and are what
  the ``-m gpu`` parity tests compare the CUDA path against on the GPU box,
  where ``/root/reference`` does not exist.
* ``combine_kat.npz`` holds small hand-computable stacks with their expected
  results worked out BY HAND below (not by running the oracle): the combine
  stage's arithmetic lives in un-vendored ccdproc/astropy, the reference has no
  vectors for it, so these known-answer cases are the anchor
  ("parity unpinned", see ``oracle/combine_oracle.py``).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from astrophotography_b200 import synth          # noqa: E402
from oracle import ref_exec                       # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def mint_calibrate():
    shape = (72, 112)
    raw_u16 = synth.science_frame(shape, seed=11, as_uint16=True, nstars=6)
    raw_f32 = synth.science_frame(shape, seed=12, as_uint16=False, nstars=6)
    bias = synth.dark_stack(5, shape).mean(0).astype(np.float32)
    dark = synth.dark_stack(5, shape, exptime=900.0).mean(0).astype(np.float32)
    dark_nb = (dark - bias).astype(np.float32)      # an already bias-subtracted dark
    flat = synth.flat_frame(shape, seed=7)
    mask = synth.badpix_mask(shape, seed=13, auto_fraction=4e-3)
    cases = {}
    variants = [
        # name, raw, raw_hdr, dark, dark_hdr, flat?, mask?, dp, still_biased
        ("u16_full_dp2_biased", raw_u16, {"EXPTIME": 300.0, "PEDESTAL": -100}, dark, {"EXPTIME": 900.0}, True, True, 2, True),
        ("u16_full_dp1", raw_u16, {"EXPOSURE": 300.0}, dark_nb, {"EXPOSURE": 900.0}, True, True, 1, False),
        ("f32_noflat_nomask", raw_f32, {"EXPTIME": 120.0, "PEDESTAL": 0}, dark_nb, {"EXPTIME": 900.0}, False, False, 2, False),
        ("f32_flat_nomask_biased", raw_f32, {"EXPTIME": 60.0, "PEDESTAL": 12.5}, dark, {"EXPTIME": 70.0}, True, False, 2, True),
        ("u16_noflat_mask_dp2", raw_u16, {"EXPTIME": 300.0}, dark_nb, {"EXPTIME": 300.0}, False, True, 2, False),
    ]
    store = dict(raw_u16=raw_u16, raw_f32=raw_f32, bias=bias, dark=dark, dark_nb=dark_nb,
                 flat=flat, mask=mask)
    meta = []
    for (name, raw, rhdr, drk, dhdr, useflat, usemask, dp, sb) in variants:
        out, hdr, nf = ref_exec.ref_calibrate(
            raw, dict(rhdr), bias, drk, dict(dhdr), flat if useflat else None,
            mask if usemask else None, dp, sb)
        store[f"out_{name}"] = out
        if useflat:
            store[f"normflat_{name}"] = nf
        img_exp = rhdr.get("EXPOSURE", rhdr.get("EXPTIME"))
        drk_exp = dhdr.get("EXPOSURE", dhdr.get("EXPTIME"))
        bp = [int(hdr.get(k, -1)) for k in ("BPIXNBAD", "BPIXNFIX", "BPIXNREM")]
        meta.append((name, "raw_u16" if raw is raw_u16 else "raw_f32",
                     "dark" if drk is dark else "dark_nb",
                     float(rhdr.get("PEDESTAL", 0)), float(img_exp), float(drk_exp),
                     int(useflat), int(usemask), dp, int(sb), *bp))
    store["meta_json"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(GOLD, "calibrate_small.npz"), **store)
    print("calibrate_small.npz", len(variants), "variants")


def mint_badpix():
    rng = np.random.default_rng(2024)
    store = {}
    names = []

    def add(name, data, mask, dp):
        out, st = ref_exec.ref_fix_bad_pixels(data, mask, dp)
        store[f"data_{name}"] = data
        store[f"mask_{name}"] = mask
        store[f"out_{name}"] = out
        store[f"stat_{name}"] = np.array(
            [dp, int(st["BPIXNBAD"][0]), int(st["BPIXNFIX"][0]), int(st["BPIXNREM"][0])])
        names.append(name)

    shape = (64, 96)
    base = rng.normal(2000, 50, size=shape).astype(np.float32)
    # (1) the etc/user_badpixels.yml shapes: corner pixel, cols 12/13/17, 2x6 block,
    #     a rectangle whose interior is unfixable at dp=2 -- uint8 mask with values 1,2,3,4.
    m = synth.badpix_mask(shape, seed=13, auto_fraction=5e-3)
    for dp in (1, 2, 3):
        add(f"yml_dp{dp}", base, m, dp)
    # (2) float32 mask (what ApCalibrate hands over, ApCalibrate.py:304-307) on a
    #     frame with NaN / +-inf donors and exact ties.
    d2 = np.rint(base).astype(np.float32)
    d2[5, 5] = np.nan
    d2[10, 40] = np.inf
    d2[11, 41] = -np.inf
    d2[30, 30] = np.inf
    d2[30, 31] = -np.inf
    m2 = np.zeros(shape, np.float32)
    for (r, c) in [(4, 4), (5, 6), (6, 5), (10, 41), (11, 40), (12, 42), (30, 32), (29, 30),
                   (0, 0), (0, 95), (63, 0), (63, 95), (0, 50), (63, 50), (20, 0), (20, 95)]:
        m2[r, c] = 1.0
    m2[40:48, 60:70] = 2.0                    # dense block: interior unfixable at dp 1,2
    m2[50, :] = 2.0                           # a full bad row
    for dp in (1, 2):
        add(f"special_dp{dp}", d2, m2, dp)
    # (3) no bad pixels at all, and everything bad.
    add("none_dp2", base, np.zeros(shape, np.uint8), 2)
    add("all_dp2", base, np.ones(shape, np.int16), 2)
    # (4) tiny images where the window is clipped on every side.
    tiny = rng.normal(100, 5, size=(3, 4)).astype(np.float32)
    tm = np.zeros((3, 4), np.uint8)
    tm[1, 1] = 1
    tm[0, 3] = 3
    add("tiny_dp2", tiny, tm, 2)
    add("tiny_dp1", tiny, tm, 1)
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "badpix_cases.npz"), **store)
    print("badpix_cases.npz", len(names), "cases")


def mint_findbadpix():
    shape = (320, 440)       # large enough for every rule in etc/user_badpixels.yml
    dark = synth.dark_stack(6, shape, exptime=900.0).mean(0).astype(np.float32)
    yml = os.path.join(ref_exec.REFERENCE_ROOT, "etc", "user_badpixels.yml")
    mask, nauto, nuser = ref_exec.ref_find_bad_pixels(dark, 4.0, yml)
    mask_auto, nauto2, _ = ref_exec.ref_find_bad_pixels(dark, 4.0, None)
    # a small frame on which several rules fall outside the image and are skipped
    small = (100, 15)
    dsmall = synth.dark_stack(4, small, exptime=900.0).mean(0).astype(np.float32)
    mask_small, nauto_s, nuser_s = ref_exec.ref_find_bad_pixels(dsmall, 4.0, yml)
    np.savez_compressed(os.path.join(GOLD, "findbadpix.npz"), dark=dark, mask=mask,
                        mask_auto=mask_auto, counts=np.array([nauto, nuser, nauto2]),
                        dark_small=dsmall, mask_small=mask_small,
                        counts_small=np.array([nauto_s, nuser_s]))
    print("findbadpix.npz", nauto, nuser, nauto_s, nuser_s)


def mint_combine_kat():
    """Known-answer stacks, expected values derived by hand.

    Layout: ``stack`` is (N, 1, P): one row of P independent pixels.
    S = 1.482602218505602.
    """
    S = 1.482602218505602
    nan = np.nan
    cases = {}

    # ---- KAT-A: N=5, median/MAD 5-sigma single pass (ApMasterCal setting) ----
    # px0: [1,2,3,4,100]  med=3, |d|=[2,1,0,1,97] -> MAD=1, s=S, bounds 3-+5S=[-4.413,10.413]
    #      100 rejected -> mean(1,2,3,4)=2.5, nrej=1
    # px1: [10,10,10,10,10] MAD=0 -> bounds [10,10], nothing <10 or >10 -> mean 10, nrej 0
    # px2: [10,10,10,10,11] med=10, |d|=[0,0,0,0,1] MAD=0 -> bounds [10,10]; 11>10 rejected -> mean 10, nrej 1
    # px3: [nan,1,2,3,4]   nan rejected up front; kept [1,2,3,4]: med=2.5, |d|=[1.5,.5,.5,1.5] MAD=1.0
    #      bounds 2.5-+5S -> all kept -> mean 2.5, nrej 1
    # px4: all nan -> data nan, nrej 5, allmasked 1
    # px5: [0,0,0,0,inf]  inf rejected up front; rest MAD=0, kept -> mean 0, nrej 1
    a = np.array([[1, 2, 3, 4, 100], [10, 10, 10, 10, 10], [10, 10, 10, 10, 11],
                  [nan, 1, 2, 3, 4], [nan] * 5, [0, 0, 0, 0, np.inf]], dtype=np.float32).T
    cases["A_stack"] = a[:, None, :]
    cases["A_params"] = np.array([5.0, 5.0, 1, 1, 1])         # klo khi maxiters cen(1=median) dev(1=mad)
    cases["A_mean"] = np.array([2.5, 10.0, 10.0, 2.5, nan, 0.0])
    cases["A_nrej"] = np.array([1, 0, 1, 1, 5, 1])
    cases["A_allmasked"] = np.array([0, 0, 0, 0, 1, 0])
    # uncert = std(kept, ddof=0)/sqrt(nkept): px0 kept 1,2,3,4: var=1.25 -> sqrt(1.25)/2
    cases["A_uncert"] = np.array([np.sqrt(1.25) / 2, 0.0, 0.0, np.sqrt(1.25) / 2, nan, 0.0])

    # ---- KAT-B: N=4 / N=5 plain median, min, max (no clipping) ----
    # px0: [4,1,3,2] median=(2+3)/2=2.5 ; px1: [1,1,2,2] -> 1.5 ; px2: [-0.0,0.0,5,-5] -> 0
    # px3: [1e30,-1e30,1,2] -> 1.5 ; px4: [nan,1,2,3] nanmedian -> 2 ; px5: [3,3,3,3] -> 3
    b = np.array([[4, 1, 3, 2], [1, 1, 2, 2], [-0.0, 0.0, 5, -5], [1e30, -1e30, 1, 2],
                  [nan, 1, 2, 3], [3, 3, 3, 3]], dtype=np.float32).T
    cases["B_stack"] = b[:, None, :]
    cases["B_median"] = np.array([2.5, 1.5, 0.0, 1.5, 2.0, 3.0])
    cases["B_min"] = np.array([1, 1, -5, np.float32(-1e30), 1, 3], dtype=np.float64)
    cases["B_max"] = np.array([4, 2, 5, np.float32(1e30), 3, 3], dtype=np.float64)
    cases["B_mean"] = np.array([2.5, 1.5, 0.0, 0.75, 2.0, 3.0])
    cases["B_nrej"] = np.array([0, 0, 0, 0, 1, 0])       # NaN samples are never used

    # ---- KAT-C: iterative mean/std clip, k=1.5 (so that small N can reject), maxiters=5 ----
    # px0: [0,0,0,0,0,0,0,10]  it1: mean=1.25, var=(7*1.5625+76.5625)/8=10.9375 std=3.30719
    #      bounds 1.25-+4.9608=[-3.71,6.21] -> 10 rejected.  it2: all 0: mean 0 std 0 bounds [0,0] nothing
    #      rejected -> stop.  mean 0, nrej 1
    # px1: [1,2,3,4,5,6,7,8] mean 4.5 std=sqrt(5.25)=2.2913 bounds [1.063,7.937] -> 1 and 8 rejected
    #      it2: [2..7] mean 4.5 std=sqrt(17.5/6)=1.7078 bounds [1.938,7.062] none -> mean 4.5 nrej 2
    # px2: [5]*8 -> mean 5, nrej 0
    c = np.array([[0, 0, 0, 0, 0, 0, 0, 10], [1, 2, 3, 4, 5, 6, 7, 8], [5] * 8],
                 dtype=np.float32).T
    cases["C_stack"] = c[:, None, :]
    cases["C_params"] = np.array([1.5, 1.5, 5, 0, 0])         # cen 0=mean, dev 0=std
    cases["C_mean"] = np.array([0.0, 4.5, 5.0])
    cases["C_nrej"] = np.array([1, 2, 0])

    # ---- KAT-D: asymmetric bounds, median/std, maxiters=1 ----
    # px0: [1,2,3,4,20] median 3, mean 6, var=(25+16+9+4+196)/5=50 std=7.0711
    #      k_lo=0.2 -> lo=3-1.4142=1.5858 -> 1 rejected; k_hi=3 -> hi=24.2 none. mean(2,3,4,20)=7.25 nrej 1
    d = np.array([[1, 2, 3, 4, 20]], dtype=np.float32).T
    cases["D_stack"] = d[:, None, :]
    cases["D_params"] = np.array([0.2, 3.0, 1, 1, 0])
    cases["D_mean"] = np.array([7.25])
    cases["D_nrej"] = np.array([1])
    np.savez_compressed(os.path.join(GOLD, "combine_kat.npz"), **cases)
    print("combine_kat.npz")


def mint_imarith():
    """Inputs and the outputs of the reference's ApImArith.process_files executed verbatim."""
    rng = np.random.default_rng(77)
    shape = (23, 41)                        # odd sizes: vector body + scalar tail
    a = rng.normal(2000, 300, shape).astype(np.float32)
    a[0, :3] = [0.0, -0.0, np.inf]
    b32 = rng.normal(10, 4, shape).astype(np.float32)
    b32[1, :3] = [0.0, np.nan, -0.0]
    b64 = rng.normal(10, 4, shape)
    store = dict(a=a, b32=b32, b64=b64)
    meta = []
    for op in ("ADD", "SUB", "MUL", "DIV"):
        for name, val in (("scalar", 3.3), ("scalar0", 0.0), ("b32", b32), ("b64", b64)):
            out, hdr, hist = ref_exec.ref_imarith(a, op, val, units="adu/s" if op == "DIV" else None)
            store[f"out_{op}_{name}"] = out
            meta.append((op, name, str(out.dtype), hdr.get("BUNIT"), len(hist)))
    out, hdr, hist = ref_exec.ref_imarith(a, " sub ", 1.5, hdr1={"PEDESTAL": -100, "BUNIT": "adu"})
    store["out_ped"] = out
    meta.append(("sub_ped", "scalar", str(out.dtype), hdr.get("BUNIT"), "PEDESTAL" in hdr))
    store["meta_json"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(GOLD, "imarith.npz"), **store)
    print("imarith.npz", len(meta), "cases")


ORDER_CASES = {
    # name: (generator, N, shape, seed, k_lo, k_hi, maxiters, cen, dev)
    "dark30_apmastercal": ("dark", 30, (64, 96), 0, 5.0, 5.0, 1, "median", "mad_std"),
    "dark30q_apmastercal": ("darkq", 30, (64, 96), 0, 5.0, 5.0, 1, "median", "mad_std"),
    "dark100_kappa3": ("dark", 100, (32, 96), 0, 3.0, 3.0, 5, "mean", "std"),
    "dark100q_kappa3": ("darkq", 100, (32, 96), 0, 3.0, 3.0, 5, "mean", "std"),
    "dark200q_kappa3": ("darkq", 200, (16, 96), 0, 3.0, 3.0, 5, "mean", "std"),
    "dark30q_astropy_default": ("darkq", 30, (64, 96), 0, 3.0, 3.0, 5, "median", "std"),
    "tenths6_k1_mean_std": ("tenths", 6, (200, 200), 1, 1.0, 1.0, 1, "mean", "std"),
    "tenths6_k1p5_mean_std_3it": ("tenths", 6, (200, 200), 2, 1.5, 1.5, 3, "mean", "std"),
    "tenths8_k1_medmad": ("tenths", 8, (200, 200), 3, 1.0, 1.0, 1, "median", "mad_std"),
}


def order_case_stack(name):
    """The seeded stack of one ORDER_CASES entry (the GPU tests rebuild it with the same call)."""
    gen, n, shape, seed, *_ = ORDER_CASES[name]
    if gen in ("dark", "darkq"):
        return synth.dark_stack(n, shape, exptime=300.0, quantise=(gen == "darkq"))
    rng = np.random.default_rng(seed)
    # multiples of 0.1 in float32: sample, centre and k*std all land on few-digit decimals, so
    # samples sit exactly on clip bounds and the two operation orders round differently
    return rng.integers(0, 12, (n,) + tuple(shape)).astype(np.float32) * np.float32(0.1)


def mint_order_census():
    """Tie census per configuration: where ``lo = c - k*s; x < lo`` (astropy / ccdproc >= 2.4) and
    ``x - c < -k*s`` (ccdproc <= 2.3) disagree.  Every disagreement must be a boundary tie."""
    from oracle import combine_oracle as C
    out = {}
    for name, (_gen, _n, _shape, _seed, k_lo, k_hi, maxiters, cen, dev) in ORDER_CASES.items():
        out[name] = C.order_census(order_case_stack(name), k_lo, k_hi, maxiters, cen, dev)
        assert out[name]["differences_are_ties"], name
    with open(os.path.join(GOLD, "combine_order_census.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("combine_order_census.json", {k: (v["orders_differ"], v["boundary_ties"]) for k, v in out.items()})


def main():
    os.makedirs(GOLD, exist_ok=True)
    if "--order-census" in sys.argv:
        mint_order_census()
        return
    if "--imarith" in sys.argv:
        mint_imarith()
        return
    if not ref_exec.reference_available():
        raise SystemExit("needs /root/reference (authoring container)")
    mint_calibrate()
    mint_badpix()
    mint_findbadpix()
    mint_combine_kat()
    mint_order_census()
    mint_imarith()


if __name__ == "__main__":
    main()
