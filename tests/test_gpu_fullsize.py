"""GPU parity at BASELINE.json's full sizes, through checks that do not need the whole
CPU oracle: (i) the oracle on sampled row bands of the same frames, (ii) the exact float64
generic kernel against the fast kernels over the full frame, (iii) size-independent
properties (min <= median <= max, a clipped mean lies inside the clip bounds' hull,
row-band sharding reproduces the single-launch result, repair is the identity off-mask)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cube(torch, n, h, w, seed, quantise=False):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    cube = torch.empty((n, h, w), dtype=torch.float32, device="cuda")
    cube.normal_(1000.0, 12.0, generator=g)
    hits = torch.randint(0, n * h * w, (n * h * w // 5000,), generator=g, device="cuda")
    cube.view(-1)[hits] += 500.0 + 29500.0 * torch.rand(hits.numel(), generator=g, device="cuda")
    if quantise:
        cube.round_()
    return cube


def _rows_oracle(cube, rows, **p):
    from oracle import combine_oracle as C
    sample = cube[:, rows].cpu().numpy()
    return C.combine(sample, want_uncert=False, **p)


def _close(got, exp, rtol=1e-6, scale=12.0):
    return bool(np.all(np.abs(got.astype(np.float64) - exp) <= rtol * np.maximum(np.abs(exp), scale)))


@pytest.mark.parametrize("params,kernel", [
    (dict(method="average", k_lo=3.0, k_hi=3.0, maxiters=5, cen="mean", dev="std"), "meanclip<100>"),
    (dict(method="average", k_lo=5.0, k_hi=5.0, maxiters=1, cen="median", dev="mad_std"), "sorted_medmad1<100>"),
    (dict(method="median", k_lo=5.0, k_hi=5.0, maxiters=0, cen="median", dev="mad_std"), "sorted_median<100>"),
])
def test_full_frame_cmos_stack_100x9576x6388(cuda, params, kernel):
    """BASELINE config 4 shape (per GPU): 100 x (9576 x 6388) float32, 24.5 GB."""
    torch = cuda
    from astrophotography_b200 import kernels
    n, h, w = 100, 6388, 9576
    assert kernels.stack_kernel_name(n, **params) == kernel
    cube = _cube(torch, n, h, w, seed=5)
    res = kernels.stack_reduce(cube, **params)
    rows = [0, 1, 777, 3193, 6386, 6387]
    exp = _rows_oracle(cube, rows, **params)
    got = res["data"][rows].cpu().numpy()
    assert np.array_equal(res["nrej"][rows].cpu().numpy().astype(np.int64), exp["nrej"])
    if params["method"] == "median":
        assert np.array_equal(got, exp["data"].astype(np.float32))
    else:
        assert _close(got, exp["data"])
    # the exact generic kernel over a 512-row band of the same cube must agree everywhere
    band = slice(2900, 3412)
    gen = kernels.stack_reduce(cube[:, band].contiguous(), force_generic=True, **params)
    assert torch.equal(gen["nrej"], res["nrej"][band])
    d = (gen["data"].double() - res["data"][band].double()).abs()
    tol = 1e-6 * torch.maximum(gen["data"].double().abs(), torch.tensor(12.0, device="cuda", dtype=torch.float64))
    assert bool((d <= tol).all())
    # row-band sharding (2, 3 and 8 bands) reproduces the single launch bit for bit
    from astrophotography_b200 import pipeline
    for world in (2, 3, 8):
        out = {"data": torch.zeros((h, w), device="cuda"), "nrej": torch.zeros((h, w), dtype=torch.uint8, device="cuda")}
        for rank in range(world):
            r0, r1, _, _ = pipeline.row_band(h, world, rank)
            kernels.stack_reduce(cube, row0=r0, nrows=r1 - r0, out=out, **params)
        assert torch.equal(out["data"], res["data"]) and torch.equal(out["nrej"], res["nrej"])
    # properties
    lo = cube.amin(dim=0)
    hi = cube.amax(dim=0)
    assert bool(((res["data"] >= lo) & (res["data"] <= hi)).all())
    assert int(res["nrej"].max()) < n
    del cube


def test_full_frame_cmos_stack_200x9576x6388_config4(cuda):
    """BASELINE config 4 itself: 200 x (9576 x 6388) float32 (48.9 GB on one GPU), kappa-sigma (3, 5 iterations):
    oracle on sampled rows, and the 2 / 4 / 8 row bands a sharded run would reduce reproduce the single launch."""
    torch = cuda
    from astrophotography_b200 import kernels, pipeline
    n, h, w = 200, 6388, 9576
    p = dict(method="average", k_lo=3.0, k_hi=3.0, maxiters=5, cen="mean", dev="std")
    cube = _cube(torch, n, h, w, seed=9)
    res = kernels.stack_reduce(cube, **p)
    assert kernels.stack_last_staging() == 5
    rows = [0, 1, 798, 799, 3193, 3194, 6387]                        # incl. band boundaries of the 2 / 8 GPU splits
    exp = _rows_oracle(cube, rows, **p)
    assert np.array_equal(res["nrej"][rows].cpu().numpy().astype(np.int64), exp["nrej"])
    assert _close(res["data"][rows].cpu().numpy(), exp["data"])
    for world in (2, 4, 8):
        out = {"data": torch.zeros((h, w), device="cuda"), "nrej": torch.zeros((h, w), dtype=torch.uint8, device="cuda")}
        for rank in range(world):
            r0, r1, _, _ = pipeline.row_band(h, world, rank)
            kernels.stack_reduce(cube, row0=r0, nrows=r1 - r0, out=out, **p)
        # rejection maps identical; the means are bit-identical except in each band's < 32-pixel tail, which the
        # register kernel reduces instead of the lane-cooperative one (another float32 summation order: ~1e-8)
        assert torch.equal(out["nrej"], res["nrej"])
        diff = out["data"] != res["data"]
        assert int(diff.sum()) <= 32 * world
        d = (out["data"].double() - res["data"].double()).abs()
        assert bool((d <= 2e-7 * res["data"].double().abs().clamp_min(12.0)).all())
    del cube


def test_full_frame_uint16_stack_100x9576x6388(cuda):
    """100 raw uint16 frames of 9576 x 6388 (12.2 GB): the uint16 kernels give what the float32 kernels give on
    float32(frames) over the whole frame -- identical rejection maps, same clipped means / medians -- and the
    oracle agrees on sampled rows."""
    torch = cuda
    from astrophotography_b200 import kernels
    n, h, w = 100, 6388, 9576
    cube = _cube(torch, n, h, w, seed=13, quantise=True).clamp_(0, 65535)
    u16 = cube.to(torch.int32).to(torch.int16).view(torch.uint16)
    rows = [0, 3193, 6387]
    for p in (dict(method="average", k_lo=3.0, k_hi=3.0, maxiters=5, cen="mean", dev="std"),
              dict(method="average", k_lo=5.0, k_hi=5.0, maxiters=1, cen="median", dev="mad_std"),
              dict(method="median", k_lo=5.0, k_hi=5.0, maxiters=0, cen="median", dev="mad_std")):
        a = kernels.stack_reduce(u16, **p)
        b = kernels.stack_reduce(cube, **p)
        assert torch.equal(a["nrej"], b["nrej"])
        if p["method"] == "median":
            assert torch.equal(a["data"], b["data"])
        else:
            d = (a["data"].double() - b["data"].double()).abs()
            assert bool((d <= 1e-6 * b["data"].double().abs().clamp_min(12.0)).all())
        exp = _rows_oracle(cube, rows, **p)
        assert np.array_equal(a["nrej"][rows].cpu().numpy().astype(np.int64), exp["nrej"])
        assert _close(a["data"][rows].cpu().numpy(), exp["data"])
    del cube, u16


def test_dslr_stack_with_repair_100x6000x4000(cuda):
    """BASELINE config 3: kappa-sigma (3, 5 iterations) stack of 100 frames 6000x4000, then repair dp=2."""
    torch = cuda
    from astrophotography_b200 import kernels, synth
    from oracle import badpix_oracle as bo
    n, h, w = 100, 4000, 6000
    p = dict(method="average", k_lo=3.0, k_hi=3.0, maxiters=5, cen="mean", dev="std")
    cube = _cube(torch, n, h, w, seed=7, quantise=True)          # integer-valued samples: ties
    cube[:, ::2, ::2] = 0.0                                       # Bayer-split channel frames are 3/4 zeros
    res = kernels.stack_reduce(cube, **p)
    rows = [0, 1, 2001, 3999]
    exp = _rows_oracle(cube, rows, **p)
    assert np.array_equal(res["nrej"][rows].cpu().numpy().astype(np.int64), exp["nrej"])
    assert _close(res["data"][rows].cpu().numpy(), exp["data"])
    assert float(res["data"][::2, ::2].abs().max()) == 0.0 and int(res["nrej"][::2, ::2].max()) == 0
    mask_np = synth.badpix_mask((h, w), auto_fraction=1e-3)
    mask = torch.from_numpy(mask_np).cuda()
    fixed, counts = kernels.fix_badpix(res["data"], mask, 2)
    good = mask == 0
    assert torch.equal(fixed[good], res["data"][good])           # identity off-mask
    band = slice(190, 320)                                        # covers the unfixable 101 x 21 rectangle
    e, st = bo.fix_bad_pixels_vec(res["data"][band.start - 2:band.stop + 2].cpu().numpy(),
                                  mask_np[band.start - 2:band.stop + 2], 2)
    assert np.array_equal(fixed[band].cpu().numpy(), e[2:-2], equal_nan=True)
    c = counts.cpu().numpy()
    assert int(c[0]) == int((mask_np != 0).sum()) and 0 < int(c[1]) < int(c[0])


def test_master_and_calibrate_30x4096x4096(cuda):
    """BASELINE config 2: master bias/dark (30 x 4096^2, ApMasterCal setting) + calibrate + repair."""
    torch = cuda
    from astrophotography_b200 import kernels, synth
    from oracle import badpix_oracle as bo, calibrate_oracle as co
    n, h, w = 30, 4096, 4096
    ref = dict(method="average", k_lo=5.0, k_hi=5.0, maxiters=1, cen="median", dev="mad_std")
    masters = {}
    rows = [0, 17, 2048, 4095]
    for name, seed in (("bias", 11), ("dark", 12)):
        cube = _cube(torch, n, h, w, seed=seed)
        res = kernels.stack_reduce(cube, **ref)
        exp = _rows_oracle(cube, rows, **ref)
        assert np.array_equal(res["nrej"][rows].cpu().numpy().astype(np.int64), exp["nrej"])
        assert _close(res["data"][rows].cpu().numpy(), exp["data"])
        masters[name] = res["data"]
        del cube
    flat = torch.from_numpy(synth.flat_frame((h, w))).cuda()
    nf, norm = kernels.flat_normalise(flat)
    assert np.float32(norm.item()).tobytes() == np.float32(co.flat_norm_factor(flat.cpu().numpy())).tobytes()
    raw_np = synth.science_frame((h, w), nstars=30)
    raw = torch.from_numpy(raw_np.view(np.int16)).cuda().view(torch.uint16)
    cal = kernels.calibrate(raw, masters["bias"], masters["dark"], nf, 300.0 / 900.0, True, pedestal=-100.0)
    exp = co.calibrate(co.read_convert(raw_np, -100.0), masters["bias"].cpu().numpy(), masters["dark"].cpu().numpy(),
                       300.0 / 900.0, co.normalise_flat(flat.cpu().numpy()), True)
    assert np.array_equal(cal.cpu().numpy(), exp, equal_nan=True)                 # bit-exact at full size
    mask_np = synth.badpix_mask((h, w), auto_fraction=1e-3)
    fixed, counts = kernels.fix_badpix(cal, torch.from_numpy(mask_np).cuda(), 2)
    e, st = bo.fix_bad_pixels_vec(exp, mask_np, 2)
    assert np.array_equal(fixed.cpu().numpy(), e, equal_nan=True)
    c = counts.cpu().numpy()
    assert (int(c[0]), int(c[1])) == (st["BPIXNBAD"][0], st["BPIXNFIX"][0])
