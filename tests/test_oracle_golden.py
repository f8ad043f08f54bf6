"""The oracle (oracle/tce_oracle.c, oracle/quant.py) against fixtures produced by the reference itself
(tests/golden/make_golden.py: reference Python quantizer + reference C++ kernels compiled in place)."""
import numpy as np
import pytest

from oracle import capi, quant


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_quantizer_formats_match_reference_python(golden_dir, name):
    g = np.load(golden_dir / f"quant_{name}.npz")
    w = g["w"]
    qs, d, zp = quant.quantize_q4_6(w)
    assert np.array_equal(qs, g["q4_6_qs"])
    assert np.array_equal(d.view(np.uint16), g["q4_6_d"].view(np.uint16))
    assert np.array_equal(zp, g["q4_6_zp"])
    assert zp.shape[1] == quant.calculate_zeros_width(w.shape[1]) == capi.zeros_width(w.shape[1])
    if "q4_3_qs" in g:
        qs3, d3 = quant.quantize_q4_3(w)
        assert np.array_equal(qs3, g["q4_3_qs"])
        assert np.array_equal(d3, g["q4_3_d"])
    qs5, d5 = quant.quantize_q4_5(w)
    assert np.array_equal(qs5, g["q4_5_qs"])
    assert np.array_equal(d5.view(np.uint16), g["q4_5_d"].view(np.uint16))


def test_naive_mat_mul_int4_bit_exact(golden_dir):
    g = np.load(golden_dir / "kernels_generic.npz")
    for tag, blk in (("g128", 128), ("g32", 32)):
        C = capi.naive_mat_mul_int4(g[f"int4_{tag}_A"], g[f"int4_{tag}_B"], g[f"int4_{tag}_s"], 8.0, blk)
        assert np.array_equal(C.view(np.uint32), g[f"int4_{tag}_C"].view(np.uint32)), tag


def test_w4a16_gemv_oracle_consistent_with_naive(golden_dir):
    """The QM_CUDA-layout oracle (per-group zeros, fp16 scales) equals naive_mat_mul_int4 when zeros == 8 and the
    scales/activations are exactly representable in fp16."""
    rng = np.random.default_rng(5)
    OC, IC, M = 24, 1408, 2  # 11 groups -> zeros_w 2, scales padded to 16 (the IC=11008 padding quirk in small)
    w = (rng.standard_normal((OC, IC)) * 0.02).astype(np.float32)
    qs, d, zp = quant.quantize_q4_6(w)
    x = rng.standard_normal((M, IC)).astype(np.float16)
    y = capi.w4a16_gemv(x, qs, zp, d)
    y2 = capi.naive_mat_mul_int4(x.astype(np.float32), quant.qmcuda_to_sequential_bytes(qs), d[:, : IC // 128].astype(np.float32), 8.0, 128)
    # same products; (q-z)*s vs s*(q-z) commute exactly, accumulation order identical
    assert np.array_equal(y.view(np.uint32), y2.view(np.uint32))
    # and a dynamic-zero case against a float64 evaluation
    zp2 = rng.integers(0, 2**32, zp.shape, dtype=np.uint32)
    y3 = capi.w4a16_gemv(x, qs, zp2, d)
    ref = x.astype(np.float64) @ quant.dequant_qmcuda(qs, d, zp2).astype(np.float64).T
    assert np.max(np.abs(y3 - ref)) <= 1e-4 * max(1.0, np.max(np.abs(ref)))


def test_int8_family_bit_exact(golden_dir):
    g = np.load(golden_dir / "kernels_generic.npz")
    A, B, Bb, b8, bf = g["i8_A"], g["i8_B"], g["i8_Bb"], g["i8_b8"], g["i8_bf"]
    alpha, beta = float(g["i8_alpha"]), float(g["i8_beta"])
    for v in range(8):
        Bv = Bb if v in (3, 7) else B
        qmin = 0 if v == 1 else -128
        C = capi.int8_matmul(v, A, Bv, b8, bf, alpha, beta, qmin, 127)
        want = g[f"i8_C{v}"]
        if C.dtype == np.int8:
            assert np.array_equal(C, want), v
        else:
            assert np.array_equal(C.view(np.uint32), want.view(np.uint32)), v
    assert np.array_equal(capi.int8_matmul(0, A, B, b8, bf, 0.05, 1.0, -128, 127), g["i8_Csat"])
    assert np.abs(g["i8_Csat"].astype(np.int32)).max() == 128 or g["i8_Csat"].max() == 127


def test_naive_mat_mul_int8(golden_dir):
    g = np.load(golden_dir / "kernels_generic.npz")
    A, B = g["i8_A"], g["i8_B"]
    M, K = A.shape
    N = B.shape[0]
    C = np.zeros((M, N), np.int8)
    capi.lib().orc_naive_mat_mul_int8(A, np.ascontiguousarray(B.T), C, M, N, K, 3, -2, 0.02, 0.01, 0.35, -128, 127)
    assert np.array_equal(C, g["i8_naive_C"])


def test_fp16_int4_host_reference(golden_dir):
    g = np.load(golden_dir / "kernels_generic.npz")
    A, qs, d = g["f16_A"], g["f16_qs"], g["f16_d"]
    M, IC = A.shape
    OC = d.shape[1]
    C = np.zeros((M, OC), np.uint16)
    capi.lib().orc_naive_mat_mul_fp16_int4(A.view(np.uint16), qs, d.view(np.uint16), C, M, IC, OC, 128)
    assert np.array_equal(C, g["f16_C"])


def test_mat_mul_transposed(golden_dir):
    g = np.load(golden_dir / "kernels_generic.npz")
    C = np.zeros_like(g["t_C"])
    capi.lib().orc_mat_mul_transposed(g["t_A"], g["t_B"], C, 3, 7, 40)
    assert np.array_equal(C, g["t_C"])


def test_half_conversions_match_numpy():
    L = capi.lib()
    L.orc_half_to_float.restype = __import__("ctypes").c_float
    L.orc_float_to_half.restype = __import__("ctypes").c_uint16
    L.orc_float_to_half.argtypes = [__import__("ctypes").c_float]
    allh = np.arange(65536, dtype=np.uint16)
    f = allh.view(np.float16).astype(np.float32)
    for h in list(range(0, 65536, 97)) + [0x0001, 0x03ff, 0x0400, 0x7bff, 0x7c00, 0xfc00, 0x8000]:
        got = L.orc_half_to_float(int(h))
        if np.isnan(f[h]):
            assert np.isnan(got)
        else:
            assert got == f[h]
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.standard_normal(2000) * 10 ** rng.uniform(-9, 5, 2000), [65504, 65519.9, 65520, 1e-8, 2.0**-25, 2.0**-24, 0.0, -0.0]]).astype(np.float32)
    for x in xs:
        with np.errstate(over="ignore"):
            want = np.float32(x).astype(np.float16).view(np.uint16)
        assert L.orc_float_to_half(float(x)) == int(want), x
