"""GPU parity: frame-stack reducer vs the float64 numpy oracle (oracle/combine_oracle.py).

Bars: median/min/max selection bit-exact; clipped means
within 1e-6 relative in float32 (the generic kernel and every float64 output
are checked far tighter); rejection-count maps identical.
"""
import os

import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu

RTOL32 = 1e-6          # north_star tolerance for float32 clipped means


def _stack(n, shape, seed=0, quantise=False, specials=True):
    rng = np.random.default_rng(seed * 1000 + n)
    h, w = shape
    st = rng.normal(1000.0, 12.0, size=(n, h, w)).astype(np.float32)
    hits = rng.random((n, h, w)) < 0.004                       # cosmic-ray-like outliers
    st[hits] += rng.uniform(500, 30000, size=int(hits.sum())).astype(np.float32)
    st[:, 0, : min(w, 6)] += 4000.0                            # hot pixels: high in every frame
    if quantise:
        st = np.rint(st).astype(np.float32)
    if specials and h * w > 40:
        st[0, 1, 1] = np.nan
        st[n // 2, 1, 2] = np.inf
        st[n - 1, 1, 3] = -np.inf
        st[:, 1, 4] = np.nan                                   # an all-NaN pixel
        st[:, 1, 5] = 7.0                                      # all samples equal
        st[: n // 2, 1, 6] = 5.0                               # two values only
        st[n // 2:, 1, 6] = 6.0
    return st


def _run(torch, st, **kw):
    from astrophotography_b200 import kernels
    cube = torch.from_numpy(np.ascontiguousarray(st)).cuda()
    res = kernels.stack_reduce(cube, want_nrej=True, want_allmasked=True, **kw)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in res.items()}


def _oracle(st, method, k_lo, k_hi, maxiters, cen, dev):
    from oracle import combine_oracle as C
    return C.combine(st, method, k_lo, k_hi, maxiters, cen, dev)


def _assert_close_data(got, exp, rtol, scale):
    both_nan = np.isnan(got) & np.isnan(exp)
    with np.errstate(invalid="ignore"):
        same_inf = np.isinf(exp) & (got == exp)
        ok = both_nan | same_inf | (np.abs(got - exp) <= rtol * np.maximum(np.abs(exp), scale))
    assert ok.all(), (np.argwhere(~ok)[:5], got[~ok][:5], exp[~ok][:5])


def test_kat_hand_computed(cuda, golden_dir):
    """The hand-derived known-answer stacks (tests/golden/combine_kat.npz)."""
    torch = cuda
    k = np.load(os.path.join(golden_dir, "combine_kat.npz"))
    def close(got, exp, force, skip=()):
        # The generic kernel is the oracle's arithmetic (1e-15).  The fast kernels promise
        # 1e-6 of max(|mean|, spread): KAT pixel B[3] = [1e30, -1e30, 1, 2] is ill-conditioned
        # by design and is only checked on the generic path.
        keep = np.ones(exp.shape, bool)
        if not force:
            keep[list(skip)] = False
        return np.allclose(got[keep], exp[keep], rtol=1e-15 if force else 1e-6, atol=0, equal_nan=True)

    for force in (False, True):
        a = _run(torch, k["A_stack"], method="average", k_lo=5, k_hi=5, maxiters=1, cen="median",
                 dev="mad_std", out_f64=True, want_uncert=True, force_generic=force)
        assert close(a["data"][0], k["A_mean"], force)
        assert np.array_equal(a["nrej"][0], k["A_nrej"])
        assert np.array_equal(a["allmasked"][0], k["A_allmasked"])
        assert np.allclose(a["uncert"][0], k["A_uncert"], rtol=1e-15, atol=0, equal_nan=True)
        for m, key in (("median", "B_median"), ("min", "B_min"), ("max", "B_max"), ("average", "B_mean")):
            b = _run(torch, k["B_stack"], method=m, maxiters=0, out_f64=True, force_generic=force)
            assert close(b["data"][0], k[key], force, skip=(3,) if m == "average" else ()), (m, b["data"][0])
            assert np.array_equal(b["nrej"][0], k["B_nrej"])
        c = _run(torch, k["C_stack"], method="average", k_lo=1.5, k_hi=1.5, maxiters=5, cen="mean",
                 dev="std", out_f64=True, force_generic=force)
        assert close(c["data"][0], k["C_mean"], force) and np.array_equal(c["nrej"][0], k["C_nrej"])
        d = _run(torch, k["D_stack"], method="average", k_lo=0.2, k_hi=3.0, maxiters=1, cen="median",
                 dev="std", out_f64=True, force_generic=force)
        assert close(d["data"][0], k["D_mean"], force) and np.array_equal(d["nrej"][0], k["D_nrej"])


GENERIC_CASES = [
    # method, k_lo, k_hi, maxiters, cen, dev
    ("average", 5.0, 5.0, 1, "median", "mad_std"),     # ApMasterCal (ap_combine_darks.py:394-399)
    ("average", 3.0, 3.0, 5, "median", "std"),         # astropy.stats.sigma_clip defaults
    ("average", 3.0, 3.0, 5, "mean", "std"),           # kappa-sigma, BASELINE config 3
    ("average", 2.5, 4.0, None, "mean", "mad_std"),    # asymmetric, to convergence
    ("median", 3.0, 3.0, 2, "median", "mad_std"),
    ("median", 5.0, 5.0, 0, "median", "mad_std"),      # plain median
    ("average", 5.0, 5.0, 0, "median", "mad_std"),     # plain mean
    ("min", 3.0, 3.0, 1, "mean", "std"),
    ("max", 3.0, 3.0, 0, "mean", "std"),
]


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 10, 31, 100, 130])
@pytest.mark.parametrize("case", GENERIC_CASES, ids=lambda c: "-".join(map(str, c)))
def test_generic_kernel_bit_exact_f64(cuda, n, case):
    """The generic kernel is the oracle's arithmetic in its operation order: float64
    outputs must be bit-identical, rejection maps identical, on every parameter set."""
    torch = cuda
    method, k_lo, k_hi, maxiters, cen, dev = case
    st = _stack(n, (12, 40), seed=1, quantise=(n % 2 == 0))
    exp = _oracle(st, method, k_lo, k_hi, maxiters, cen, dev)
    got = _run(torch, st, method=method, k_lo=k_lo, k_hi=k_hi, maxiters=maxiters, cen=cen, dev=dev,
               out_f64=True, want_uncert=True, force_generic=True)
    assert np.array_equal(got["nrej"].astype(np.int64), exp["nrej"])
    assert np.array_equal(got["allmasked"], exp["allmasked"])
    assert bits_equal(got["data"], exp["data"])
    if not (maxiters == 0 and not np.isfinite(st).all()):
        assert bits_equal(got["uncert"], exp["uncert"])


FAST_CASES = [
    ("average", 5.0, 5.0, 1, "median", "mad_std", "sorted_medmad1"),
    ("average", 3.0, 3.0, 1, "median", "mad_std", "sorted_medmad1"),
    ("median", 5.0, 5.0, 0, "median", "mad_std", "sorted_median"),
    ("average", 3.0, 3.0, 5, "mean", "std", "meanclip"),
    ("average", 2.0, 3.5, None, "mean", "std", "meanclip"),
    ("average", 3.0, 3.0, 1, "mean", "std", "meanclip"),
    ("average", 3.0, 3.0, 0, "mean", "std", "meanclip"),
]


@pytest.mark.parametrize("n", [3, 4, 5, 8, 9, 16, 17, 24, 30, 33, 50, 64, 65, 81, 90, 99, 100, 101, 113, 128, 129, 160, 161, 200])
@pytest.mark.parametrize("case", FAST_CASES, ids=lambda c: "-".join(map(str, c)))
@pytest.mark.parametrize("quantise", [False, True])
def test_fast_kernels_match_oracle(cuda, n, case, quantise):
    torch = cuda
    from astrophotography_b200 import kernels
    method, k_lo, k_hi, maxiters, cen, dev, family = case
    name = kernels.stack_kernel_name(n, method, k_lo, k_hi, maxiters, cen, dev)
    assert name.startswith(family), name
    st = _stack(n, (9, 70), seed=2, quantise=quantise)
    exp = _oracle(st, method, k_lo, k_hi, maxiters, cen, dev)
    # the kappa-sigma family has a register-resident and a shared-memory-resident kernel: check both
    variants = [(f, p) for f in (False, True)
                for p in (("registers", "registers_tma", "registers_direct", "registers_cpasync", "shared") if family == "meanclip" else (None,))]
    for out_f64, prefer in variants:
        if prefer is not None:
            kn = kernels.stack_kernel_name(n, method, k_lo, k_hi, maxiters, cen, dev, prefer=prefer)
            assert kn == ("meanclip_smem" if prefer == "shared" else kn) and kn.startswith("meanclip"), kn
        # (9, 70) = 630 pixels: 4 full TMA tiles + a 118-pixel tail through the direct kernel
        got = _run(torch, st, method=method, k_lo=k_lo, k_hi=k_hi, maxiters=maxiters, cen=cen, dev=dev,
                   out_f64=out_f64, prefer=prefer)
        assert np.array_equal(got["nrej"].astype(np.int64), exp["nrej"]), (name, out_f64)
        assert np.array_equal(got["allmasked"], exp["allmasked"])
        if family == "sorted_median":
            # selection is comparison-only: exact (the even-N midpoint of two float32 is exact in float64)
            e = exp["data"] if out_f64 else exp["data"].astype(np.float32)
            assert bits_equal(got["data"], e)
        else:
            # float64 output: the sorted kernel sums in float64 (<= a few ulp of the oracle);
            # the meanclip kernel sums pivot-shifted float32 values exactly in float64.
            rt = RTOL32 if not out_f64 else (1e-13 if family.startswith("sorted") else 2e-7)
            _assert_close_data(got["data"].astype(np.float64), exp["data"], rt, 12.0)


@pytest.mark.parametrize("n", [130, 160, 200, 257, 448])
@pytest.mark.parametrize("prefer", ["registers", "shared"])
def test_meanclip_large_n(cuda, n, prefer):
    torch = cuda
    from astrophotography_b200 import kernels
    assert kernels.stack_kernel_name(n, "average", 3, 3, 5, "mean", "std").startswith("meanclip")
    st = _stack(n, (6, 64), seed=4)
    exp = _oracle(st, "average", 3.0, 3.0, 5, "mean", "std")
    got = _run(torch, st, method="average", k_lo=3.0, k_hi=3.0, maxiters=5, cen="mean", dev="std", want_uncert=True,
               prefer=prefer)
    assert np.array_equal(got["nrej"].astype(np.int64), exp["nrej"])
    _assert_close_data(got["data"].astype(np.float64), exp["data"], RTOL32, 1.0)
    _assert_close_data(got["uncert"].astype(np.float64), exp["uncert"], 1e-5, 1e-3)


@pytest.mark.parametrize("n", [3, 8, 30, 31, 64, 81, 100, 127, 160, 200])
@pytest.mark.parametrize("case", [c for c in FAST_CASES if c[-1] == "meanclip"], ids=lambda c: "-".join(map(str, c)))
def test_meanclip_tensormap_staging(cuda, n, case):
    """Equally spaced frames (a torch cube whose frame size is a multiple of 16 bytes) go through the
    warp-granular tensor-map TMA pipeline: same results as the oracle, and the path really ran."""
    torch = cuda
    from astrophotography_b200 import kernels
    method, k_lo, k_hi, maxiters, cen, dev, _ = case
    st = _stack(n, (11, 76), seed=21 + n, quantise=(n % 2 == 1))     # 836 pixels: 26 warp tiles + 4-pixel tail
    exp = _oracle(st, method, k_lo, k_hi, maxiters, cen, dev)
    for out_f64 in (False, True):
        got = _run(torch, st, method=method, k_lo=k_lo, k_hi=k_hi, maxiters=maxiters, cen=cen, dev=dev,
                   out_f64=out_f64, prefer="registers_tensormap")
        assert kernels.stack_last_staging() == 3
        assert np.array_equal(got["nrej"].astype(np.int64), exp["nrej"])
        assert np.array_equal(got["allmasked"], exp["allmasked"])
        _assert_close_data(got["data"].astype(np.float64), exp["data"], RTOL32 if not out_f64 else 2e-7, 12.0)
    # frames that are not 16-byte spaced fall back to direct loads
    st2 = _stack(n, (9, 70), seed=3)
    _run(torch, st2, method=method, k_lo=k_lo, k_hi=k_hi, maxiters=maxiters, cen=cen, dev=dev, prefer="registers_tensormap")
    assert kernels.stack_last_staging() == 0


@pytest.mark.parametrize("n", [101, 127, 128, 129, 160, 161, 199, 200, 201, 256, 257, 320, 399, 448, 512, 513, 641, 800, 1000, 1024])
@pytest.mark.parametrize("case", [c for c in FAST_CASES if c[-1] == "meanclip"][:2], ids=lambda c: "-".join(map(str, c)))
def test_meanclip_lane_split_long_stacks(cuda, n, case):
    """Long stacks on equally spaced frames: above N = 128, 2, 4 or 8 lanes share a pixel (lane-split tensor-map kernels)."""
    torch = cuda
    from astrophotography_b200 import kernels
    method, k_lo, k_hi, maxiters, cen, dev, _ = case
    st = _stack(n, (7, 76), seed=40 + n, quantise=(n % 2 == 1))     # 532 pixels: full warp tiles + a 20-pixel tail
    exp = _oracle(st, method, k_lo, k_hi, maxiters, cen, dev)
    for out_f64 in (False, True):
        got = _run(torch, st, method=method, k_lo=k_lo, k_hi=k_hi, maxiters=maxiters, cen=cen, dev=dev,
                   out_f64=out_f64, want_uncert=True)
        # (N <= 128: the single-thread register kernel on warp-granular tensor-map tiles leads there)
        assert kernels.stack_last_staging() == (3 if n <= 128 else (5 if n <= 512 else 4))
        assert np.array_equal(got["nrej"].astype(np.int64), exp["nrej"])
        assert np.array_equal(got["allmasked"], exp["allmasked"])
        _assert_close_data(got["data"].astype(np.float64), exp["data"], RTOL32 if not out_f64 else 2e-7, 12.0)
        _assert_close_data(got["uncert"].astype(np.float64), exp["uncert"], 1e-5, 1e-3)


@pytest.mark.parametrize("n", [30, 100, 150, 300, 600])
def test_meanclip_separately_allocated_frames(cuda, n):
    """N frame pointers that are NOT equally spaced (the C ABI allows it): no tensor map can describe them,
    the dispatcher must fall back to the pointer-table kernels and still match the oracle."""
    torch = cuda
    from astrophotography_b200 import kernels
    st = _stack(n, (9, 70), seed=60 + n)
    exp = _oracle(st, "average", 3.0, 3.0, 5, "mean", "std")
    frames, pad = [], []
    for i in range(n):
        frames.append(torch.from_numpy(np.ascontiguousarray(st[i])).cuda())
        pad.append(torch.empty(640 * (1 + i % 3), device="cuda"))            # unequal gaps between the frames
    gaps = {frames[i + 1].data_ptr() - frames[i].data_ptr() for i in range(n - 1)}
    assert len(gaps) > 1
    res = kernels.stack_reduce(frames, method="average", k_lo=3.0, k_hi=3.0, maxiters=5, cen="mean", dev="std",
                               want_nrej=True, want_allmasked=True)
    torch.cuda.synchronize()
    assert kernels.stack_last_staging() in (-1, 0, 2)
    assert np.array_equal(res["nrej"].cpu().numpy().astype(np.int64), exp["nrej"])
    _assert_close_data(res["data"].cpu().numpy().astype(np.float64), exp["data"], RTOL32, 12.0)
    del pad


@pytest.mark.parametrize("n,shape,row0,nrows", [(100, (8, 75), 3, 4), (64, (16, 76), 5, 9), (200, (8, 75), 1, 6),
                                                (300, (12, 68), 2, 9)])
def test_meanclip_row_band_on_cube(cuda, n, shape, row0, nrows):
    """Row bands of a cube (multi-GPU sharding): the band may start at a pixel that is not 16-byte aligned
    (a TMA box must start on one: the dispatcher then falls back to the pointer-table kernels, which this test
    found out the hard way); rows outside the band are left untouched."""
    torch = cuda
    from astrophotography_b200 import kernels
    st = _stack(n, shape, seed=70 + n)
    exp = _oracle(st, "average", 3.0, 3.0, 5, "mean", "std")
    cube = torch.from_numpy(st).cuda()
    out = {"data": torch.full(shape, -7.0, dtype=torch.float32, device="cuda"),
           "nrej": torch.full(shape, 255, dtype=torch.uint8 if n <= 255 else torch.uint16, device="cuda")}
    kernels.stack_reduce(cube, method="average", k_lo=3.0, k_hi=3.0, maxiters=5, cen="mean", dev="std",
                         row0=row0, nrows=nrows, out=out)
    torch.cuda.synchronize()
    aligned = (row0 * shape[1]) % 4 == 0
    assert (kernels.stack_last_staging() in (3, 5)) == aligned
    data, nrej = out["data"].cpu().numpy(), out["nrej"].cpu().numpy()
    band = slice(row0, row0 + nrows)
    assert np.array_equal(nrej[band].astype(np.int64), exp["nrej"][band])
    _assert_close_data(data[band].astype(np.float64), exp["data"][band], RTOL32, 12.0)
    keep = np.ones(shape[0], bool); keep[band] = False
    assert (data[keep] == -7.0).all() and (nrej[keep] == 255).all()


@pytest.mark.parametrize("n", [3, 4, 9, 30, 31, 64, 100, 127, 160, 200])
@pytest.mark.parametrize("case", [c for c in FAST_CASES if c[-1].startswith("sorted")], ids=lambda c: "-".join(map(str, c)))
def test_sorted_tensormap_staging(cuda, n, case):
    """Equally spaced frames: the sorted kernels take their 256-pixel tiles through one tensor-map TMA box
    (partial last tile, row band starting off a tile boundary); same results as the oracle."""
    torch = cuda
    from astrophotography_b200 import kernels
    method, k_lo, k_hi, maxiters, cen, dev, family = case
    st = _stack(n, (11, 76), seed=90 + n, quantise=(n % 2 == 1))     # 836 pixels: 3 full tiles + 68 pixels
    exp = _oracle(st, method, k_lo, k_hi, maxiters, cen, dev)
    for out_f64 in (False, True):
        got = _run(torch, st, method=method, k_lo=k_lo, k_hi=k_hi, maxiters=maxiters, cen=cen, dev=dev, out_f64=out_f64)
        assert kernels.stack_last_staging() == 1
        assert np.array_equal(got["nrej"].astype(np.int64), exp["nrej"])
        if family == "sorted_median":
            e = exp["data"] if out_f64 else exp["data"].astype(np.float32)
            assert bits_equal(got["data"], e)
        else:
            _assert_close_data(got["data"].astype(np.float64), exp["data"], RTOL32 if not out_f64 else 1e-13, 12.0)
    # a row band (rows 2..8) of the same cube
    cube = torch.from_numpy(st).cuda()
    out = {"data": torch.full((11, 76), -7.0, dtype=torch.float32, device="cuda")}
    kernels.stack_reduce(cube, method=method, k_lo=k_lo, k_hi=k_hi, maxiters=maxiters, cen=cen, dev=dev,
                         row0=2, nrows=7, out=out, want_nrej=False)
    torch.cuda.synchronize()
    d = out["data"].cpu().numpy()
    _assert_close_data(d[2:9].astype(np.float64), exp["data"][2:9], RTOL32, 12.0)
    assert (d[:2] == -7.0).all() and (d[9:] == -7.0).all()


def test_fast_uncert(cuda):
    torch = cuda
    st = _stack(30, (8, 64), seed=5)
    exp = _oracle(st, "average", 5.0, 5.0, 1, "median", "mad_std")
    got = _run(torch, st, method="average", out_f64=True, want_uncert=True)
    assert np.array_equal(got["nrej"].astype(np.int64), exp["nrej"])
    _assert_close_data(got["uncert"], exp["uncert"], 1e-12, 1e-6)
    _assert_close_data(got["data"], exp["data"], 1e-13, 1.0)


def test_unclipped_pixels_mean_is_bit_exact(cuda):
    """Where nothing is rejected the sorted kernel sums in frame order like np.nanmean."""
    torch = cuda
    rng = np.random.default_rng(8)
    st = rng.normal(1000, 10, size=(30, 16, 64)).astype(np.float32)
    exp = _oracle(st, "average", 5.0, 5.0, 1, "median", "mad_std")
    got = _run(torch, st, method="average", out_f64=True)
    keep_all = exp["nrej"] == 0
    assert keep_all.mean() > 0.9
    assert np.array_equal(got["data"][keep_all], exp["data"][keep_all])


def test_row_band_only_touches_its_rows(cuda):
    torch = cuda
    from astrophotography_b200 import kernels
    st = _stack(10, (32, 48), seed=6, specials=False)
    exp = _oracle(st, "average", 5.0, 5.0, 1, "median", "mad_std")
    cube = torch.from_numpy(st).cuda()
    out = {"data": torch.full((32, 48), -1.0, device="cuda"),
           "nrej": torch.full((32, 48), 255, dtype=torch.uint8, device="cuda")}
    kernels.stack_reduce(cube, row0=5, nrows=11, out=out)
    d = out["data"].cpu().numpy()
    assert (d[:5] == -1).all() and (d[16:] == -1).all()
    _assert_close_data(d[5:16].astype(np.float64), exp["data"][5:16], RTOL32, 1.0)
    assert (out["nrej"].cpu().numpy()[16:] == 255).all()


def test_separate_frame_pointers_and_u16_counts(cuda):
    """Frames as N separate allocations (the reference stacks N files) and N > 255."""
    torch = cuda
    from astrophotography_b200 import kernels
    n = 260
    st = _stack(n, (4, 33), seed=7, specials=False)
    frames = [torch.from_numpy(st[i].copy()).cuda() for i in range(n)]
    res = kernels.stack_reduce(frames, method="average", k_lo=3, k_hi=3, maxiters=3, cen="median", dev="std",
                               out_f64=True)
    assert res["nrej"].dtype == torch.uint16
    exp = _oracle(st, "average", 3.0, 3.0, 3, "median", "std")
    assert np.array_equal(res["nrej"].cpu().numpy().astype(np.int64), exp["nrej"])
    assert bits_equal(res["data"].cpu().numpy(), exp["data"])


def test_stack_errors(cuda):
    torch = cuda
    from astrophotography_b200 import kernels
    a = torch.zeros((3, 4, 4), device="cuda")
    with pytest.raises(RuntimeError):
        kernels.stack_reduce(a.to(torch.float64))
    with pytest.raises(RuntimeError):
        kernels.stack_reduce([a[0], a[1, :2]])
    with pytest.raises(RuntimeError):
        kernels.stack_reduce(a, method="mode")
    with pytest.raises(RuntimeError, match="geometry"):
        kernels.stack_reduce(a, row0=3, nrows=2)
    with pytest.raises(RuntimeError):
        kernels.stack_reduce(torch.zeros((0, 4, 4), device="cuda"))


# ---------------------------------------------------------------- the two published operation orders
def _order_case_names():
    from oracle import make_golden as G
    return sorted(G.ORDER_CASES)


@pytest.mark.parametrize("name", _order_case_names())
def test_nrej_equals_both_operation_orders_off_ties(cuda, golden_dir, name):
    """Rejection maps equal the astropy order (``x < c - k*s``) everywhere and the ccdproc <= 2.3 order
    (``x - c < -k*s``) everywhere except on boundary-tie pixels; the tie census is a committed golden."""
    import json
    from oracle import combine_oracle as C, make_golden as G
    torch = cuda
    _gen, _n, _shape, _seed, k_lo, k_hi, maxiters, cen, dev = G.ORDER_CASES[name]
    st = G.order_case_stack(name)
    with open(os.path.join(golden_dir, "combine_order_census.json")) as f:
        census = json.load(f)[name]
    assert C.order_census(st, k_lo, k_hi, maxiters, cen, dev) == census
    a = C.combine(st, "average", k_lo, k_hi, maxiters, cen, dev, want_uncert=False, order="bounds")
    b = C.combine(st, "average", k_lo, k_hi, maxiters, cen, dev, want_uncert=False, order="deviation")
    ties = C.boundary_ties(st, k_lo, k_hi, maxiters, cen, dev)
    for force in (False, True):
        got = _run(torch, st, method="average", k_lo=k_lo, k_hi=k_hi, maxiters=maxiters, cen=cen, dev=dev,
                   force_generic=force)
        nrej = got["nrej"].astype(np.int64)
        assert np.array_equal(nrej, a["nrej"]), (name, force)
        assert np.array_equal(nrej[~ties], b["nrej"][~ties]), (name, force)
        _assert_close_data(got["data"], a["data"], RTOL32, 1.0)


def test_real_ccdproc_cross_check_on_this_box(cuda):
    """Attempted on the GPU box too (the authoring container has no ccdproc/astropy): when the real
    library is importable, the oracle AND the CUDA path are compared with it."""
    ccdproc = pytest.importorskip("ccdproc")
    from astropy.nddata import CCDData
    from astropy.stats import mad_std
    from oracle import combine_oracle as C
    torch = cuda
    rng = np.random.default_rng(9)
    st = rng.normal(1000, 12, (12, 16, 16)).astype(np.float32)
    st[2, 3, 3] += 5000
    ccds = [CCDData(f, unit="adu") for f in st]
    ref = ccdproc.combine(ccds, method="average", sigma_clip=True, sigma_clip_low_thresh=5, sigma_clip_high_thresh=5,
                          sigma_clip_func=np.ma.median, sigma_clip_dev_func=mad_std)
    mine = C.combine(st, "average", 5, 5, 1, "median", "mad_std")
    assert np.allclose(np.asarray(ref.data), mine["data"], rtol=1e-12)
    got = _run(torch, st, out_f64=True)
    assert np.allclose(got["data"], np.asarray(ref.data), rtol=1e-12)


# ---------------------------------------------------------------- median/MAD clip extremes, median + uncertainty
@pytest.mark.parametrize("n", [21, 30, 64, 100, 128, 200])
def test_medmad_uncert_and_frame_order_mean(cuda, n):
    """The reference's ApMasterCal setting with the uncertainty plane (what ApMasterCal always asks for):
    rejection maps identical, unclipped pixels' float64 mean bit-identical to np.nanmean (frame-order sum)."""
    torch = cuda
    st = _stack(n, (13, 84), seed=120 + n, quantise=(n % 2 == 0))
    for k_lo, k_hi in ((5.0, 5.0), (2.0, 3.0)):
        e = _oracle(st, "average", k_lo, k_hi, 1, "median", "mad_std")
        for out_f64 in (False, True):
            got = _run(torch, st, method="average", k_lo=k_lo, k_hi=k_hi, maxiters=1, cen="median", dev="mad_std",
                       out_f64=out_f64, want_uncert=True)
            assert np.array_equal(got["nrej"].astype(np.int64), e["nrej"])
            assert np.array_equal(got["allmasked"], e["allmasked"])
            _assert_close_data(got["data"].astype(np.float64), e["data"], RTOL32 if not out_f64 else 1e-13, 12.0)
            _assert_close_data(got["uncert"].astype(np.float64), e["uncert"], 1e-5 if not out_f64 else 1e-12, 1e-3)
            if out_f64:
                keep_all = (e["nrej"] == 0) & np.isfinite(e["data"])
                assert np.array_equal(got["data"][keep_all], e["data"][keep_all])


def test_medmad_constant_frames_and_outlier_in_every_pixel(cuda):
    """MAD = 0 (every sample equals the median: nothing clipped) and a stack where every pixel holds an outlier."""
    torch = cuda
    n, shape = 40, (6, 96)
    const = np.full((n,) + shape, 123.0, np.float32)
    got = _run(torch, const, method="average", out_f64=True, want_uncert=True)
    assert (got["data"] == 123.0).all() and (got["nrej"] == 0).all() and (got["uncert"] == 0).all()
    rng = np.random.default_rng(3)
    st = rng.normal(500, 5, (n,) + shape).astype(np.float32)
    st[rng.integers(0, n, shape), np.arange(shape[0])[:, None], np.arange(shape[1])[None, :]] += 9000.0
    exp = _oracle(st, "average", 5.0, 5.0, 1, "median", "mad_std")
    assert (exp["nrej"] >= 1).all()
    got = _run(torch, st, method="average", out_f64=True)
    assert np.array_equal(got["nrej"].astype(np.int64), exp["nrej"])
    _assert_close_data(got["data"], exp["data"], 1e-13, 1.0)


@pytest.mark.parametrize("n", [3, 4, 9, 16, 30, 31, 64, 100, 101, 200])
def test_median_with_uncertainty_fast_kernel(cuda, n):
    """ApMasterCal(method='median') always asks for the uncertainty plane: median (bit-exact) +
    1.4826 * MAD / sqrt(N) from the sorted column, no generic kernel."""
    torch = cuda
    from astrophotography_b200 import kernels
    assert kernels.stack_kernel_name(n, "median", maxiters=0, want_uncert=True).startswith("sorted_median_mad")
    st = _stack(n, (11, 76), seed=200 + n, quantise=(n % 2 == 1))
    exp = _oracle(st, "median", 5.0, 5.0, 0, "median", "mad_std")
    for out_f64 in (False, True):
        got = _run(torch, st, method="median", maxiters=0, out_f64=out_f64, want_uncert=True)
        e = exp["data"] if out_f64 else exp["data"].astype(np.float32)
        assert bits_equal(got["data"], e)
        assert np.array_equal(got["nrej"].astype(np.int64), exp["nrej"])
        ok = np.isfinite(exp["uncert"]) & np.isfinite(st).all(axis=0)
        _assert_close_data(got["uncert"].astype(np.float64)[ok], exp["uncert"][ok], 1e-6 if not out_f64 else 1e-14, 1e-3)


# ---------------------------------------------------------------- uint16 frames (f2)
def _u16_stack(n, shape, seed):
    from astrophotography_b200 import synth
    rng = np.random.default_rng(seed)
    st = synth.dark_stack(n, shape, exptime=300.0, quantise=True)
    st[:, 0, :4] = np.array([0, 65535, 32767, 32768], np.float32)          # the corners of the format
    st[rng.integers(0, n), 1, :] += 40000.0
    return np.clip(st, 0, 65535).astype(np.uint16)


def _u16_tensor(torch, u16, fmt):
    if fmt == "fits":      # the data unit of a BITPIX=16 / BZERO=32768 file: big-endian int16 of (value - 32768)
        raw = (u16.astype(np.int32) - 32768).astype(">i2")
        return torch.from_numpy(np.ascontiguousarray(raw).view(np.int16).copy()).cuda()
    return torch.from_numpy(u16.view(np.int16).copy()).cuda().view(torch.uint16)


U16_CASES = [
    ("average", 5.0, 5.0, 1, "median", "mad_std"),     # ApMasterCal: two-phase sorted kernels
    ("average", 3.0, 3.0, 5, "mean", "std"),           # kappa-sigma: tensor-map meanclip kernel, 64-pixel tiles
    ("median", 5.0, 5.0, 0, "median", "mad_std"),      # plain median
    ("average", 3.0, 3.0, 5, "median", "std"),         # astropy default: generic kernel
    ("max", 3.0, 3.0, 0, "mean", "std"),
]


@pytest.mark.parametrize("n", [3, 10, 30, 31, 64, 100, 101, 160, 200])
@pytest.mark.parametrize("case", U16_CASES, ids=lambda c: "-".join(map(str, c)))
@pytest.mark.parametrize("fmt", ["native", "fits"])
def test_u16_frames_match_oracle_on_float32_of_them(cuda, n, case, fmt):
    """apgpu_stack_reduce_u16: raw 16-bit frames (host order, or the big-endian BZERO=32768 data unit of a
    FITS file) give exactly what the float32 path gives on float32(frames): identical rejection maps,
    bit-exact selection, means within 1e-6."""
    torch = cuda
    from astrophotography_b200 import kernels
    method, k_lo, k_hi, maxiters, cen, dev = case
    shape = (9, 200)                                   # 1800 pixels: 64-pixel warp tiles + tails
    u16 = _u16_stack(n, shape, seed=300 + n)
    exp = _oracle(u16.astype(np.float32), method, k_lo, k_hi, maxiters, cen, dev)
    cube = _u16_tensor(torch, u16, fmt)
    for out_f64 in (False, True):
        res = kernels.stack_reduce(cube, method=method, k_lo=k_lo, k_hi=k_hi, maxiters=maxiters, cen=cen, dev=dev,
                                   out_f64=out_f64, want_nrej=True, want_allmasked=True, want_uncert=True, u16_format=fmt)
        torch.cuda.synchronize()
        got = {k: v.cpu().numpy() for k, v in res.items()}
        assert np.array_equal(got["nrej"].astype(np.int64), exp["nrej"]), (case, fmt, out_f64)
        assert np.array_equal(got["allmasked"], exp["allmasked"])
        if method in ("median", "max"):
            e = exp["data"] if out_f64 else exp["data"].astype(np.float32)
            assert bits_equal(got["data"], e)
        else:
            _assert_close_data(got["data"].astype(np.float64), exp["data"], RTOL32 if not out_f64 else 2e-7, 12.0)
        _assert_close_data(got["uncert"].astype(np.float64), exp["uncert"], 1e-5, 1e-3)
    # the float32 path on the converted frames gives the same rejection map (same kernels, other loader)
    f32 = _run(torch, u16.astype(np.float32), method=method, k_lo=k_lo, k_hi=k_hi, maxiters=maxiters, cen=cen, dev=dev)
    assert np.array_equal(f32["nrej"], got["nrej"])


@pytest.mark.parametrize("n,row0,nrows", [(30, 3, 5), (100, 4, 4), (100, 1, 7), (64, 2, 5)])
def test_u16_row_bands_and_staging(cuda, n, row0, nrows):
    """Row bands of a uint16 cube: a band starting on a 16-byte boundary keeps the tensor-map path (staging 3),
    any other start falls back to direct loads; rows outside the band stay untouched."""
    torch = cuda
    from astrophotography_b200 import kernels
    shape = (8, 100)                                   # 200-byte rows: every other row starts on a 16-byte boundary
    u16 = _u16_stack(n, shape, seed=400 + n)
    exp = _oracle(u16.astype(np.float32), "average", 3.0, 3.0, 5, "mean", "std")
    cube = _u16_tensor(torch, u16, "native")
    out = {"data": torch.full(shape, -7.0, dtype=torch.float32, device="cuda"),
           "nrej": torch.full(shape, 255, dtype=torch.uint8, device="cuda")}
    kernels.stack_reduce(cube, method="average", k_lo=3.0, k_hi=3.0, maxiters=5, cen="mean", dev="std",
                         row0=row0, nrows=nrows, out=out)
    torch.cuda.synchronize()
    aligned = (row0 * shape[1] * 2) % 16 == 0
    assert (kernels.stack_last_staging() == 3) == aligned
    data, nrej = out["data"].cpu().numpy(), out["nrej"].cpu().numpy()
    band = slice(row0, row0 + nrows)
    assert np.array_equal(nrej[band].astype(np.int64), exp["nrej"][band])
    _assert_close_data(data[band].astype(np.float64), exp["data"][band], RTOL32, 12.0)
    keep = np.ones(shape[0], bool); keep[band] = False
    assert (data[keep] == -7.0).all() and (nrej[keep] == 255).all()


# ---------------------------------------------------------------- long medians: lane-cooperative selection
@pytest.mark.parametrize("n", [201, 202, 224, 225, 255, 256, 257, 300, 320, 321, 383, 384, 400, 448, 449, 511, 512])
@pytest.mark.parametrize("quantise", [False, True])
def test_long_median_lane_cooperative(cuda, n, quantise):
    """200 < N <= 512: 4 or 8 lanes sort a share of the samples each and select the median of the union of their
    sorted runs -- comparison-only, bit-exact (ties from quantised data, NaN / inf pixels via the generic routine,
    a tail shorter than a 32-pixel tile)."""
    torch = cuda
    from astrophotography_b200 import kernels
    assert kernels.stack_kernel_name(n, "median", maxiters=0).startswith("median_coop")
    st = _stack(n, (7, 76), seed=500 + n, quantise=quantise)            # 532 pixels: 16 tiles + a 20-pixel tail
    if quantise:
        st[:, 2, :] = np.rint(st[:, 2, :] / 8) * 8                      # many equal samples around the median
    exp = _oracle(st, "median", 5.0, 5.0, 0, "median", "mad_std")
    for out_f64 in (False, True):
        got = _run(torch, st, method="median", maxiters=0, out_f64=out_f64)
        assert kernels.stack_last_staging() == 5
        e = exp["data"] if out_f64 else exp["data"].astype(np.float32)
        assert bits_equal(got["data"], e), (n, np.argwhere(got["data"] != e)[:5])
        assert np.array_equal(got["nrej"].astype(np.int64), exp["nrej"])
    # a row band that does not start on a 16-byte boundary falls back to the generic kernel, same answer
    cube = torch.from_numpy(st).cuda()
    out = {"data": torch.full((7, 76), -7.0, dtype=torch.float32, device="cuda")}
    kernels.stack_reduce(cube, method="median", maxiters=0, row0=1, nrows=5, out=out, want_nrej=False)
    torch.cuda.synchronize()
    d = out["data"].cpu().numpy()
    assert bits_equal(d[1:6], exp["data"][1:6].astype(np.float32)) and (d[0] == -7.0).all() and (d[6] == -7.0).all()


@pytest.mark.parametrize("n", [201, 224, 256, 257, 300, 384, 449, 512])
@pytest.mark.parametrize("quantise", [False, True])
def test_long_medmad_and_median_uncertainty_lane_cooperative(cuda, n, quantise):
    """200 < N <= 512: the reference's median/MAD clip and the median with its uncertainty plane on the
    lane-cooperative kernels (MAD = a second rank selection over the deviation lists of the sorted runs, float64
    like the oracle): rejection maps identical, medians bit-exact, means within a few ulp."""
    torch = cuda
    from astrophotography_b200 import kernels
    assert kernels.stack_kernel_name(n).startswith("medmad1_coop")
    assert kernels.stack_kernel_name(n, "median", maxiters=0, want_uncert=True).startswith("median_mad_coop")
    st = _stack(n, (7, 76), seed=700 + n, quantise=quantise)
    if quantise:
        st[:, 2, :] = np.rint(st[:, 2, :] / 8) * 8                      # many ties in samples and deviations
    for k_lo, k_hi in ((5.0, 5.0), (1.5, 2.5)):
        exp = _oracle(st, "average", k_lo, k_hi, 1, "median", "mad_std")
        for out_f64 in (False, True):
            got = _run(torch, st, method="average", k_lo=k_lo, k_hi=k_hi, maxiters=1, cen="median", dev="mad_std",
                       out_f64=out_f64, want_uncert=True)
            assert kernels.stack_last_staging() == 5
            assert np.array_equal(got["nrej"].astype(np.int64), exp["nrej"]), (n, k_lo)
            assert np.array_equal(got["allmasked"], exp["allmasked"])
            _assert_close_data(got["data"].astype(np.float64), exp["data"], RTOL32 if not out_f64 else 1e-13, 12.0)
            _assert_close_data(got["uncert"].astype(np.float64), exp["uncert"], 1e-5 if not out_f64 else 1e-12, 1e-3)
    expm = _oracle(st, "median", 5.0, 5.0, 0, "median", "mad_std")
    for out_f64 in (False, True):
        got = _run(torch, st, method="median", maxiters=0, out_f64=out_f64, want_uncert=True)
        assert kernels.stack_last_staging() == 5
        e = expm["data"] if out_f64 else expm["data"].astype(np.float32)
        assert bits_equal(got["data"], e)
        ok = np.isfinite(expm["uncert"]) & np.isfinite(st).all(axis=0)
        _assert_close_data(got["uncert"].astype(np.float64)[ok], expm["uncert"][ok], 1e-6 if not out_f64 else 1e-14, 1e-3)


# ---------------------------------------------------------------------------
# marked pixels: what the fast kernels cannot finish is redone by the scan launch (csrc/stack_generic.cu)
# ---------------------------------------------------------------------------
MARK_LAYOUTS = {"f32+u8": dict(out_f64=False, want_nrej=True),
                "f64+uncert+counts": dict(out_f64=True, want_nrej=True, want_uncert=True),
                "f32 alone": dict(out_f64=False, want_nrej=False),
                "f64 alone": dict(out_f64=True, want_nrej=False)}
MARK_PARAMS = {"kappa": dict(method="average", k_lo=3.0, k_hi=3.0, maxiters=5, cen="mean", dev="std"),
               "plain mean": dict(method="average", k_lo=3.0, k_hi=3.0, maxiters=0, cen="mean", dev="std"),
               "medmad": dict(method="average", k_lo=5.0, k_hi=5.0, maxiters=1, cen="median", dev="mad_std"),
               "median": dict(method="median", k_lo=3.0, k_hi=3.0, maxiters=0, cen="mean", dev="std")}


@pytest.mark.parametrize("n", [20, 100, 256, 300])
@pytest.mark.parametrize("layout", list(MARK_LAYOUTS))
@pytest.mark.parametrize("params", list(MARK_PARAMS))
def test_marked_pixels_are_finished_by_the_scan_launch(cuda, n, layout, params):
    """Many pixels with NaN / inf samples (each is marked by the fast kernel), constant pixels with one NaN (the
    warp-cooperative float64 routine gives up on sd = 0 and hands over to the generic one), an odd width and a row
    band that starts off every 16-byte boundary, with and without a rejection map to scan."""
    torch = cuda
    from astrophotography_b200 import kernels
    h, w, row0, nrows = 37, 203, 3, 29
    st = _stack(n, (h, w), seed=11)
    rng = np.random.default_rng(n)
    flat = st.reshape(n, -1)
    for value, frac in ((np.nan, 0.10), (np.inf, 0.02), (-np.inf, 0.01)):
        px = np.flatnonzero(rng.random(h * w) < frac)
        flat[rng.integers(0, n, px.size), px] = value
    flat[:, 5 * w + 7] = np.nan                                   # nothing left
    flat[:, 6 * w + 9] = 5.0
    flat[n // 3, 6 * w + 9] = np.nan                              # constant but for one NaN
    kw, lay = MARK_PARAMS[params], MARK_LAYOUTS[layout]
    exp = _oracle(st, kw["method"], kw["k_lo"], kw["k_hi"], kw["maxiters"], kw["cen"], kw["dev"])
    cube = torch.from_numpy(st).cuda()
    dt = torch.float64 if lay["out_f64"] else torch.float32
    out = {"data": torch.full((h, w), -1.0, dtype=dt, device="cuda")}
    if lay["want_nrej"]:
        out["nrej"] = torch.full((h, w), 7, dtype=torch.uint16 if n > 255 else torch.uint8, device="cuda")
    res = kernels.stack_reduce(cube, row0=row0, nrows=nrows, out=out, want_allmasked=True, **kw, **lay)
    torch.cuda.synchronize()
    band = slice(row0, row0 + nrows)
    d = res["data"].cpu().numpy().astype(np.float64)
    assert (d[:row0] == -1).all() and (d[row0 + nrows:] == -1).all()
    if params == "median":
        e = exp["data"][band] if lay["out_f64"] else exp["data"][band].astype(np.float32).astype(np.float64)
        _assert_close_data(d[band], e, 0.0, 1.0)
    else:
        _assert_close_data(d[band], exp["data"][band], 1e-12 if lay["out_f64"] and params == "medmad" else RTOL32, 1.0)
    # no mark survives: the marks are NaNs with a payload in the low mantissa bits
    raw = res["data"].view(torch.int64 if lay["out_f64"] else torch.int32).cpu().numpy()
    nan_bits = raw[np.isnan(res["data"].cpu().numpy())]
    assert np.all((nan_bits & 0xfffff) == 0)
    if lay["want_nrej"]:
        nr = res["nrej"].cpu().numpy().astype(np.int64)
        assert np.array_equal(nr[band], exp["nrej"][band])
        assert (nr[:row0] == 7).all() and (nr[row0 + nrows:] == 7).all()
    assert np.array_equal(res["allmasked"].cpu().numpy()[band] != 0, exp["allmasked"][band] != 0)
    if lay.get("want_uncert"):
        u = res["uncert"].cpu().numpy().astype(np.float64)
        _assert_close_data(u[band], exp["uncert"][band], 1e-6, 1e-3)


@pytest.mark.parametrize("n", [2, 3, 4, 5, 8, 9, 30, 31, 64, 99, 100, 128, 200])
@pytest.mark.parametrize("k", [(5.0, 5.0), (3.0, 3.0), (1.5, 4.0), (0.25, 0.5)])
def test_medmad_tie_heavy_stacks(cuda, n, k):
    """Median/MAD clip on few-valued integer samples (deviations tie with each other and with the clip bounds,
    MAD = 0 where more than half of the samples agree), narrow and asymmetric clip factors (most pixels DO reject),
    float32 and float64 + uncertainty outputs."""
    torch = cuda
    rng = np.random.default_rng(100 * n + int(10 * k[0]))
    h, w = 24, 512                                               # whole 256-pixel tiles and 16-byte groups
    st = rng.integers(0, 4, size=(n, h, w)).astype(np.float32) + 100.0
    st[:, :8] = rng.normal(1000.0, 12.0, size=(n, 8, w)).astype(np.float32)
    hits = rng.random((n, h, w)) < 0.01
    st[hits] += 5000.0
    st[:, 9, :7] = 42.0                                          # constant pixels
    st[0, 9, :3] = 43.0                                          # ... but for one sample
    exp = _oracle(st, "average", k[0], k[1], 1, "median", "mad_std")
    for lay in ({"out_f64": True, "want_uncert": True}, {"out_f64": False}):
        got = _run(torch, st, method="average", k_lo=k[0], k_hi=k[1], **lay)
        assert np.array_equal(got["nrej"].astype(np.int64), exp["nrej"]), (n, k, lay)
        _assert_close_data(got["data"].astype(np.float64), exp["data"], 1e-13 if lay["out_f64"] else RTOL32, 1.0)
        if lay.get("want_uncert"):
            _assert_close_data(got["uncert"], exp["uncert"], 1e-12, 1e-6)
