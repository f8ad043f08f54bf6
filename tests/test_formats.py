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


def test_x86_import_matches_requantisation_rule():
    """tce_w4_import_x86 (C++, host only): a QM_x86 op -> QM_CUDA arrays == exact dequantisation + the reference's QM_CUDA rule
    (formats.quantize_qm_cuda follows quantize_row_q4_6, pinned to the reference quantizer by tests/golden/quant_*.npz)."""
    import ctypes as C

    from tinychatengine_b200 import _lib, formats

    rng = np.random.default_rng(3)
    oc, ic = 24, 1280  # 10 groups of 128: the scale / zero rows are padded to zeros_w * 8 = 16
    w = (rng.standard_normal((oc, ic)) * 0.02).astype(np.float32)
    w[3, 128:256] = 0.0  # an all-zero group: d = 0
    qs, d = formats.quantize_qm_x86(w)
    deq = formats.dequantize_qm_x86(qs, d)
    assert np.abs(deq - w).max() < 0.02  # sanity: it is a 4-bit code of w
    ew, ez, es = formats.quantize_qm_cuda(deq)
    zw = formats.zeros_width(ic)
    gw = np.zeros((oc, ic // 8), np.uint32)
    gs = np.zeros((oc, zw * 8), np.float16)
    gz = np.zeros((oc, zw), np.uint32)
    L = _lib.lib()
    rc = L.tce_w4_import_x86(qs.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), oc, ic, gw.ctypes.data_as(C.c_void_p),
                             gs.ctypes.data_as(C.c_void_p), gz.ctypes.data_as(C.c_void_p))
    assert rc == 0
    np.testing.assert_array_equal(gw, ew)
    np.testing.assert_array_equal(gz, ez)
    np.testing.assert_array_equal(gs.view(np.uint16), es.view(np.uint16))
    assert L.tce_w4_import_x86(qs.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), oc, 100, gw.ctypes.data_as(C.c_void_p),
                               gs.ctypes.data_as(C.c_void_p), gz.ctypes.data_as(C.c_void_p)) != 0
