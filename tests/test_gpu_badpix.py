"""GPU parity: bad-pixel repair vs the verbatim-reference goldens and the numpy oracle.
Bar: repaired pixels bit-exact, counters identical."""
import os

import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu


def _to_cuda(torch, a):
    if a.dtype == np.uint16:
        return torch.from_numpy(a.view(np.int16)).cuda().view(torch.uint16)
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_badpix_matches_reference_goldens(cuda, golden_dir):
    torch = cuda
    from astrophotography_b200 import kernels
    z = np.load(os.path.join(golden_dir, "badpix_cases.npz"))
    for name in z["names"]:
        data, mask, exp = z[f"data_{name}"], z[f"mask_{name}"], z[f"out_{name}"]
        dp, nbad, nfix, nrem = [int(v) for v in z[f"stat_{name}"]]
        out, counts = kernels.fix_badpix(_to_cuda(torch, data), _to_cuda(torch, mask), dp)
        c = counts.cpu().numpy()
        assert (int(c[0]), int(c[1])) == (nbad, nfix), name
        assert bits_equal(out.cpu().numpy(), exp), name


@pytest.mark.parametrize("shape", [(5, 7), (33, 65), (200, 300), (1024, 1024), (513, 1027)])
@pytest.mark.parametrize("dp", [1, 2, 3, 5])
def test_badpix_matches_oracle(cuda, shape, dp):
    torch = cuda
    from astrophotography_b200 import kernels, synth
    from oracle import badpix_oracle as bo
    rng = np.random.default_rng(shape[0] + 31 * dp)
    data = rng.normal(500, 40, size=shape).astype(np.float32)
    data[rng.integers(0, shape[0], 5), rng.integers(0, shape[1], 5)] = np.nan
    data[rng.integers(0, shape[0], 3), rng.integers(0, shape[1], 3)] = np.inf
    data = np.where(rng.random(shape) < 0.3, np.rint(data), data).astype(np.float32)   # ties
    mask = synth.badpix_mask(shape, seed=13 + dp, auto_fraction=0.02)
    exp, st = bo.fix_bad_pixels_vec(data, mask, dp)
    out, counts = kernels.fix_badpix(_to_cuda(torch, data), _to_cuda(torch, mask), dp)
    c = counts.cpu().numpy()
    assert int(c[0]) == st["BPIXNBAD"][0] and int(c[1]) == st["BPIXNFIX"][0]
    assert bits_equal(out.cpu().numpy(), exp)


@pytest.mark.parametrize("mdtype", [np.uint8, np.int16, np.int32, np.float32, np.float64, np.bool_])
def test_badpix_mask_dtypes(cuda, mdtype):
    torch = cuda
    from astrophotography_b200 import kernels
    from oracle import badpix_oracle as bo
    rng = np.random.default_rng(3)
    data = rng.normal(100, 5, size=(40, 52)).astype(np.float32)
    mask = (rng.random((40, 52)) < 0.05).astype(mdtype)
    if mdtype not in (np.bool_,):
        mask = mask * 3
    exp, st = bo.fix_bad_pixels_vec(data, mask, 2)
    out, counts = kernels.fix_badpix(_to_cuda(torch, data), _to_cuda(torch, mask.astype(mdtype)), 2)
    assert bits_equal(out.cpu().numpy(), exp)


def test_badpix_row_bands_with_halo(cuda):
    """Row-band sharding: each band + deltapix halo reproduces the full-frame result."""
    torch = cuda
    from astrophotography_b200 import kernels, synth
    from oracle import badpix_oracle as bo
    shape = (96, 128)
    rng = np.random.default_rng(9)
    data = rng.normal(100, 5, size=shape).astype(np.float32)
    mask = synth.badpix_mask(shape, seed=2, auto_fraction=0.03)
    exp, st = bo.fix_bad_pixels_vec(data, mask, 2)
    d, m = _to_cuda(torch, data), _to_cuda(torch, mask)
    counts = torch.zeros(2, dtype=torch.int64, device="cuda")
    parts = []
    for (r0, r1) in [(0, 30), (30, 31), (31, 77), (77, 96)]:
        b0, b1 = max(0, r0 - 2), min(shape[0], r1 + 2)
        out, counts = kernels.fix_badpix(d[b0:b1], m[b0:b1], 2, image_rows=shape[0], band_row0=b0,
                                         row0=r0, nrows=r1 - r0, counts=counts)
        parts.append(out.cpu().numpy())
    assert bits_equal(np.concatenate(parts), exp)
    c = counts.cpu().numpy()
    assert int(c[0]) == st["BPIXNBAD"][0] and int(c[1]) == st["BPIXNFIX"][0]


def test_badpix_errors(cuda):
    torch = cuda
    from astrophotography_b200 import kernels
    a = torch.zeros((8, 8), device="cuda")
    with pytest.raises(RuntimeError, match="does not match"):
        kernels.fix_badpix(a, torch.zeros((8, 9), device="cuda"), 1)
    with pytest.raises(RuntimeError, match="deltapix"):
        kernels.fix_badpix(a, a, 0)
    with pytest.raises(RuntimeError, match="halo"):
        kernels.fix_badpix(a, a, 2, image_rows=20, band_row0=4, row0=4, nrows=8)
