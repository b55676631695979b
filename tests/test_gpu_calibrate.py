"""GPU parity: fused calibrate + flat normalisation vs the verbatim-reference goldens
and the numpy oracle (oracle/calibrate_oracle.py).  Bar: bit-exact float32."""
import json
import os

import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu


def _golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "calibrate_small.npz"))
    return z, json.loads(str(z["meta_json"]))


def test_calibrate_matches_reference_goldens(cuda, golden_dir):
    torch = cuda
    from astrophotography_b200 import kernels
    z, meta = _golden(golden_dir)
    dev = "cuda:0"
    bias = torch.from_numpy(z["bias"]).to(dev)
    for (name, rawk, darkk, ped, iexp, dexp, useflat, usemask, dp, sb, nbad, nfix, nrem) in meta:
        raw_np = z[rawk]
        dark = torch.from_numpy(z[darkk]).to(dev)
        normflat = None
        if useflat:
            flat = torch.from_numpy(z["flat"]).to(dev)
            normflat, norm = kernels.flat_normalise(flat)
            assert bits_equal(normflat.cpu().numpy(), z[f"normflat_{name}"]), name
        if raw_np.dtype == np.uint16:
            raw = torch.from_numpy(raw_np.view(np.int16)).to(dev).view(torch.uint16)
            cal = kernels.calibrate(raw, bias, dark, normflat, iexp / dexp, bool(sb), pedestal=ped)
        else:
            rf = raw_np.copy()
            if ped != 0:
                rf += ped          # _read_fits :322, float32 in-place add
            cal = kernels.calibrate(torch.from_numpy(rf).to(dev), bias, dark, normflat, iexp / dexp, bool(sb))
        if usemask:
            mask = torch.from_numpy(z["mask"].astype(np.float32)).to(dev)   # ApCalibrate hands a float32 mask
            cal, counts = kernels.fix_badpix(cal, mask, dp)
            c = counts.cpu().numpy()
            assert (int(c[0]), int(c[1]), int(c[0] - c[1])) == (nbad, nfix, nrem), name
        assert bits_equal(cal.cpu().numpy(), z[f"out_{name}"]), name


@pytest.mark.parametrize("shape", [(1, 1), (3, 5), (17, 33), (64, 64), (257, 1031), (1024, 1536)])
@pytest.mark.parametrize("biased", [False, True])
def test_calibrate_matches_oracle_shapes(cuda, shape, biased):
    torch = cuda
    from astrophotography_b200 import kernels, synth
    from oracle import calibrate_oracle as co
    rng = np.random.default_rng(shape[0] * 7919 + shape[1])
    raw = rng.integers(0, 65536, size=shape).astype(np.uint16)
    bias = rng.normal(1000, 12, size=shape).astype(np.float32)
    dark = rng.normal(1100, 20, size=shape).astype(np.float32)
    flat = synth.flat_frame(shape, seed=3, with_specials=shape[0] * shape[1] > 20)
    r = 300.0 / 900.0
    nf_o = co.normalise_flat(flat)
    exp = co.calibrate(co.read_convert(raw, -50.0), bias, dark, r, nf_o, biased)
    dev = "cuda:0"
    nf_g, norm = kernels.flat_normalise(torch.from_numpy(flat).to(dev))
    assert bits_equal(norm.cpu().numpy()[0:1], np.array([co.flat_norm_factor(flat)], dtype=np.float32))
    assert bits_equal(nf_g.cpu().numpy(), nf_o)
    raw_t = torch.from_numpy(raw.view(np.int16)).to(dev).view(torch.uint16)
    got = kernels.calibrate(raw_t, torch.from_numpy(bias).to(dev), torch.from_numpy(dark).to(dev),
                            nf_g, r, biased, pedestal=-50.0)
    assert bits_equal(got.cpu().numpy(), exp)
    # float32 raw, no flat
    rawf = co.read_convert(raw, None)
    exp2 = co.calibrate(rawf, bias, dark, r, None, biased)
    got2 = kernels.calibrate(torch.from_numpy(rawf).to(dev), torch.from_numpy(bias).to(dev),
                             torch.from_numpy(dark).to(dev), None, r, biased)
    assert bits_equal(got2.cpu().numpy(), exp2)


def test_calibrate_unaligned_views(cuda):
    """Row-band views whose base pointer is not 16-byte aligned take the scalar kernel."""
    torch = cuda
    from astrophotography_b200 import kernels
    from oracle import calibrate_oracle as co
    rng = np.random.default_rng(5)
    n = 4099
    arrs = [rng.normal(1000, 30, size=n + 1).astype(np.float32) for _ in range(4)]
    exp = co.calibrate(arrs[0][1:], arrs[1][1:], arrs[2][1:], 0.4, arrs[3][1:], True)
    t = [torch.from_numpy(a).cuda()[1:].view(1, n) for a in arrs]
    got = kernels.calibrate(t[0], t[1], t[2], t[3], 0.4, True)
    assert bits_equal(got.cpu().numpy()[0], exp)


def test_flat_norm_sizes(cuda):
    """np.nanmean bit-for-bit over sizes that exercise every branch of the pairwise tree."""
    torch = cuda
    from astrophotography_b200 import kernels
    for n in [1, 7, 8, 9, 127, 128, 129, 255, 256, 257, 1000, 4096, 4097, 65536, 100003, 1 << 20, 3 * 1000 * 1000 + 17]:
        rng = np.random.default_rng(n)
        a = rng.normal(30000, 2500, size=n).astype(np.float32)
        if n > 10:
            a[rng.integers(0, n, size=3)] = np.nan
        exp = np.nanmean(a)
        got = kernels.flat_norm(torch.from_numpy(a).cuda().view(1, n)).cpu().numpy()[0]
        assert np.float32(exp).tobytes() == np.float32(got).tobytes(), (n, exp, got)


def test_calibrate_errors(cuda):
    torch = cuda
    from astrophotography_b200 import kernels
    a = torch.zeros((4, 4), device="cuda")
    with pytest.raises(RuntimeError):
        kernels.calibrate(a, a[:2], a)
    with pytest.raises(RuntimeError):
        kernels.calibrate(a.cpu(), a, a)
    with pytest.raises(RuntimeError):
        kernels.calibrate(a.to(torch.float64), a, a)


# ---------------------------------------------------------------- fused calibrate + repair (batch driver's launch)
@pytest.mark.parametrize("raw_kind", ["f32", "u16", "u16_fits"])
@pytest.mark.parametrize("dp", [1, 2, 3])
@pytest.mark.parametrize("shape", [(64, 96), (37, 53)])           # vector path / scalar path (W % 4 != 0)
def test_fused_calibrate_repair_equals_two_launches(cuda, raw_kind, dp, shape):
    """apgpu_calibrate_repair == apgpu_calibrate_* followed by apgpu_fix_badpix_f32, bit for bit (float32 raw,
    host-order uint16 raw with PEDESTAL, and the big-endian BZERO=32768 FITS data unit), for every flat / bias
    mode; the big-endian output option is the byte-swapped plane."""
    torch = cuda
    from astrophotography_b200 import kernels, synth
    rng = np.random.default_rng(dp * 100 + shape[0])
    raw16 = synth.science_frame(shape, seed=3, as_uint16=True, nstars=4)
    raw16[0, :4] = [0, 65535, 32767, 32768]
    bias = rng.normal(1000, 5, shape).astype(np.float32)
    dark = rng.normal(1040, 6, shape).astype(np.float32)
    flat = synth.flat_frame(shape, seed=7)
    mask = synth.badpix_mask(shape, seed=13, auto_fraction=2e-2)
    mask[5:9, 10:14] = 1                                          # an unfixable block interior at dp=1
    nf, _ = kernels.flat_normalise(torch.from_numpy(flat).cuda())
    b_d, d_d, m_d = torch.from_numpy(bias).cuda(), torch.from_numpy(dark).cuda(), torch.from_numpy(mask).cuda()
    if raw_kind == "f32":
        raw_d, raw_ref, ped = torch.from_numpy(raw16.astype(np.float32) + np.float32(-100.0)).cuda(), None, None
    elif raw_kind == "u16":
        raw_d, ped = torch.from_numpy(raw16.view(np.int16)).cuda().view(torch.uint16), -100.0
    else:
        be = (raw16.astype(np.int32) - 32768).astype(">i2")
        raw_d, ped = torch.from_numpy(np.ascontiguousarray(be).view(np.int16).copy()).cuda(), -100.0
    for flat_on in (True, False):
        for biased in (True, False):
            nfl = nf if flat_on else None
            if raw_kind == "f32":
                cal = kernels.calibrate(raw_d, b_d, d_d, nfl, 1.0 / 3.0, biased)
            else:
                u16_d = torch.from_numpy(raw16.view(np.int16)).cuda().view(torch.uint16)
                cal = kernels.calibrate(u16_d, b_d, d_d, nfl, 1.0 / 3.0, biased, pedestal=ped)
            exp, ecounts = kernels.fix_badpix(cal, m_d, dp)
            got, counts = kernels.calibrate_repair(raw_d, b_d, d_d, nfl, 1.0 / 3.0, biased, pedestal=ped, mask=m_d,
                                                   deltapix=dp, raw_kind=raw_kind)
            assert bits_equal(got.cpu().numpy(), exp.cpu().numpy()), (raw_kind, dp, flat_on, biased)
            assert counts.tolist() == ecounts.tolist() and counts[0].item() == int((mask != 0).sum())
            nomask, _ = kernels.calibrate_repair(raw_d, b_d, d_d, nfl, 1.0 / 3.0, biased, pedestal=ped, raw_kind=raw_kind)
            assert bits_equal(nomask.cpu().numpy(), cal.cpu().numpy())
    be_out, _ = kernels.calibrate_repair(raw_d, b_d, d_d, nf, 1.0 / 3.0, True, pedestal=ped, mask=m_d, deltapix=dp,
                                         raw_kind=raw_kind, out_big_endian=True)
    ref = kernels.calibrate_repair(raw_d, b_d, d_d, nf, 1.0 / 3.0, True, pedestal=ped, mask=m_d, deltapix=dp, raw_kind=raw_kind)[0]
    assert np.array_equal(be_out.cpu().numpy().view(">f4").astype("=f4").view(np.uint32), ref.cpu().numpy().view(np.uint32))
