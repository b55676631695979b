"""GPU parity of ApImArith / apgpu_imarith_f32 against goldens minted by executing the reference's
core/ApImArith.py verbatim (oracle/ref_exec.ref_imarith -> tests/golden/imarith.npz): bit-exact float32."""
import json
import os

import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu


def test_imarith_matches_reference_goldens(cuda, golden_dir):
    torch = cuda
    from astrophotography_b200 import ApImArith, kernels
    z = np.load(os.path.join(golden_dir, "imarith.npz"))
    a = torch.from_numpy(z["a"]).cuda()
    ops = ApImArith("ERROR")
    for op, name, dtype, _bunit, _nhist in json.loads(str(z["meta_json"]))[:-1]:
        val = {"scalar": 3.3, "scalar0": 0.0, "b32": z["b32"], "b64": z["b64"]}[name]
        exp = z[f"out_{op}_{name}"]
        assert exp.dtype == np.float32 == np.dtype(dtype)
        b = val if not isinstance(val, np.ndarray) else torch.from_numpy(val).cuda()
        got = kernels.imarith(a, op, b).cpu().numpy()
        assert bits_equal(got, exp), (op, name)
        assert bits_equal(ops.apply(z["a"], op.lower() + " ", val), exp)          # class level, numpy in / out
    # an unaligned view (scalar path of the kernel)
    got = kernels.imarith(a.reshape(-1)[1:].reshape(1, -1).contiguous(), "MUL", 3.3).cpu().numpy()
    assert bits_equal(got.reshape(-1), z["out_MUL_scalar"].reshape(-1)[1:])


def test_apimarith_files_and_errors(cuda, golden_dir, tmp_path):
    from astrophotography_b200 import ApImArith, fitsio
    from astrophotography_b200.scripts import ap_imarith
    z = np.load(os.path.join(golden_dir, "imarith.npz"))
    inp = str(tmp_path / "in.fits")
    fitsio.write_image(inp, z["a"], fitsio.new_header({"PEDESTAL": -100, "BUNIT": "adu", "OBJECT": "M31"}))
    second = str(tmp_path / "b64.fits")
    fitsio.write_image(second, z["b64"], fitsio.new_header({}))
    out = str(tmp_path / "out.fits")
    ApImArith("ERROR").process_files(inp, " sub ", "1.5", out, None)
    data, hdr = fitsio.read_image(out, 0)
    assert bits_equal(data, z["out_ped"]) and "PEDESTAL" not in hdr and hdr["BUNIT"] == "adu" and hdr["OBJECT"] == "M31"
    hist = fitsio.header_history(hdr)
    assert len(hist) == 2 and "Applied ApImArith" in hist[0] and hist[1].endswith("in.fits SUB 1.5")
    plain = str(tmp_path / "plain.fits")
    fitsio.write_image(plain, z["a"], fitsio.new_header({}))
    assert ap_imarith.main([plain, "DIV", second, out, "--units", "adu/s", "-l", "ERROR"]) == 0
    data, hdr = fitsio.read_image(out, 0)
    assert bits_equal(data, z["out_DIV_b64"]) and hdr["BUNIT"] == "adu/s"
    assert fitsio.header_history(hdr)[1].endswith("plain.fits DIV b64.fits")
    with pytest.raises(ValueError, match="not one of the allowed operations"):
        ApImArith("ERROR").process_files(plain, "POW", "2", out, None)
    with pytest.raises(ValueError, match="not a scalar or a valid file path"):
        ApImArith("ERROR").process_files(plain, "ADD", str(tmp_path / "nope.fits"), out, None)
    small = str(tmp_path / "small.fits")
    fitsio.write_image(small, z["a"][:5], fitsio.new_header({}))
    with pytest.raises(ValueError, match="not a valid FITS file"):
        ApImArith("ERROR").process_files(plain, "ADD", small, out, None)
    ints = str(tmp_path / "ints.fits")
    fitsio.write_image(ints, (z["a"][:, :8] > 0).astype(np.int32), fitsio.new_header({}))
    with pytest.raises(RuntimeError, match="float32 images only"):
        ApImArith("ERROR").process_files(ints, "ADD", "1", out, None)
