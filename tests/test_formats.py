"""Host-side data formats (tinychatengine_b200/formats.py) against the fixtures made by the reference quantizer."""
import numpy as np
import pytest

from tinychatengine_b200 import formats


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_quantize_qm_cuda_matches_reference(golden_dir, name):
    g = np.load(golden_dir / f"quant_{name}.npz")
    w, z, s = formats.quantize_qm_cuda(g["w"])
    assert np.array_equal(w, g["q4_6_qs"])
    assert np.array_equal(z, g["q4_6_zp"])
    assert np.array_equal(s.view(np.uint16), g["q4_6_d"].view(np.uint16))


def test_pack_unpack_round_trip_and_dir_io(tmp_path):
    rng = np.random.default_rng(3)
    oc, ic = 32, 1408  # 11 groups: zeros_w = 2, 5 padding nibbles
    q = rng.integers(0, 16, (oc, ic), dtype=np.uint8)
    s = rng.random((oc, ic // 128)).astype(np.float16)
    z = rng.integers(0, 16, (oc, ic // 128), dtype=np.uint8)
    w, zp, sp = formats.pack_qm_cuda(q, s, z)
    assert w.shape == (oc, ic // 8) and zp.shape == (oc, 2) and sp.shape == (oc, 16)
    q2, s2, z2 = formats.unpack_qm_cuda(w, zp, sp)
    assert np.array_equal(q, q2) and np.array_equal(s, s2) and np.array_equal(z, z2)
    assert np.all(sp[:, 11:] == 0)
    formats.save_qm_cuda_dir(tmp_path / "q_proj", w, zp, sp)
    w3, z3, s3 = formats.load_qm_cuda_dir(tmp_path / "q_proj", oc, ic)
    assert np.array_equal(w, w3) and np.array_equal(zp, z3) and np.array_equal(sp, s3)
    with pytest.raises(ValueError):
        formats.load_qm_cuda_dir(tmp_path / "q_proj", oc, ic + 128)
