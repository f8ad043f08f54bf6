"""The drop-in proof: the reference's OWN call sites of the path -- Linear_half_int4 (llm/src/ops/cuda/linear.cu), the W8A8 wrappers and
BMMs, Int8OPTAttention and Int8OPTDecoderLayer -- compiled unchanged with -DQM_CUDA against the reference's kernels/matmul.h, linked with
this repo's MatmulOperator definitions on libtce_b200.so (oracle/_ref/libtce_callsites_cuda.so, built by oracle/Makefile where
/root/reference exists; the prebuilt library travels to the GPU box).  Outputs must equal the oracle's, and -- for the modules -- the
same reference sources running on the reference's own CPU kernels (libtce_ref_modules.so), bit for bit."""
import numpy as np
import pytest

from helpers import assert_w4_close

pytestmark = pytest.mark.gpu


def _need(kind):
    from oracle import capi

    if not (capi.REF_DIR / f"libtce_{kind}.so").exists():
        pytest.skip(f"oracle/_ref/libtce_{kind}.so not built")


@pytest.mark.parametrize("OC,IC,M", [(256, 1024, 1), (4096, 4096, 1), (48, 11008, 1), (128, 2048, 3)])
def test_linear_half_int4_call_site(OC, IC, M, tmp_path):
    """Linear_half_int4::forward (the reference's CUDA call site, unchanged) -> MatmulOperator::gemv_forward_cuda -> tce_w4a16_gemv."""
    _need("callsites_cuda")
    from oracle import capi, quant
    from tinychatengine_b200.formats import save_qm_cuda_dir

    rng = np.random.default_rng(OC + IC)
    w = (rng.standard_normal((OC, IC)) * 0.02).astype(np.float32)
    qs, d, zp = quant.quantize_q4_6(w)
    zp = rng.integers(0, 2**32, zp.shape, dtype=np.uint32)  # exercise per-group zero points
    save_qm_cuda_dir(tmp_path, qs, zp, d)
    x = rng.standard_normal((M, IC)).astype(np.float16)
    y = np.zeros((M, OC), np.float16)
    rc = capi.modules_lib("callsites_cuda").ref_linear_half_int4(str(tmp_path).encode(), OC, IC, M, x.view(np.uint16), y.view(np.uint16))
    assert rc == 0
    assert_w4_close(y.astype(np.float32), capi.w4a16_gemv(x, qs, zp, d), f"Linear_half_int4 {OC}x{IC} M={M}")


def _opt_params(rng, E, F):
    W = {k: rng.integers(-127, 128, (E, E), dtype=np.int8) for k in "qkvo"}
    B = {k: rng.integers(-127, 128, (E,), dtype=np.int8) for k in "qkv"}
    bo = rng.standard_normal(E).astype(np.float32)
    ln = {"ln1w": (1 + 0.1 * rng.standard_normal(E)).astype(np.float32), "ln1b": rng.standard_normal(E).astype(np.float32),
          "ln2w": (1 + 0.1 * rng.standard_normal(E)).astype(np.float32), "ln2b": rng.standard_normal(E).astype(np.float32)}
    fc = {"w1": rng.integers(-127, 128, (F, E), dtype=np.int8), "b1": rng.integers(-127, 128, (F,), dtype=np.int8),
          "w2": rng.integers(-127, 128, (E, F), dtype=np.int8), "b2": rng.standard_normal(E).astype(np.float32)}
    scales = {"a_qkv": np.float32(0.0011), "b_qkv": np.float32(0.7), "qk_alpha": np.float32(0.0009), "pv_alpha": np.float32(0.013),
              "a_out": np.float32(0.0006), "a1": np.float32(0.0008), "b1": np.float32(0.5), "a2": np.float32(0.0005)}
    return W, B, bo, ln, fc, scales


@pytest.mark.parametrize("E,H,prefill,steps,seed", [(128, 2, 5, 2, 1), (384, 6, 17, 4, 2), (256, 4, 1, 6, 3)])
def test_int8_opt_attention_module_on_this_library(E, H, prefill, steps, seed, tmp_path):
    """Int8OPTAttention::forward, reference source unchanged, its int8 matmuls running on the GPU through this library: equal to the
    oracle (and to the CPU reference build) bit for bit -- outputs and the returned int8 KV cache."""
    _need("callsites_cuda")
    from oracle import capi

    rng = np.random.default_rng(seed)
    W, B, bo, _, _, sc = _opt_params(rng, E, 4 * E)
    hidden = rng.integers(-127, 128, (prefill + steps, E), dtype=np.int8)
    par = (sc["a_qkv"], sc["b_qkv"], sc["qk_alpha"], sc["pv_alpha"], sc["a_out"])
    capi.write_opt_attention_params(tmp_path, W, B, bo, *par)
    got, gk, gv = capi.run_opt_attention("callsites_cuda", tmp_path, hidden, E, H, prefill, steps)
    want, wk, wv = capi.oracle_int8_opt_attention(hidden, W, B, bo, *par, H, prefill, steps)
    assert np.array_equal(gk, wk) and np.array_equal(gv, wv)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    if (capi.REF_DIR / "libtce_ref_modules.so").exists():
        ref, rk, rv = capi.run_opt_attention("ref_modules", tmp_path, hidden, E, H, prefill, steps)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)) and np.array_equal(gk, rk) and np.array_equal(gv, rv)


@pytest.mark.parametrize("E,H,F,prefill,steps,seed", [(128, 2, 512, 5, 2, 1), (256, 4, 768, 9, 3, 2)])
def test_int8_opt_decoder_layer_reference_cpu_vs_this_library(E, H, F, prefill, steps, seed, tmp_path):
    """Int8OPTDecoderLayer::forward (LayerNormQ -> attention -> residual -> LayerNormQ -> fc1 ReLU -> fc2 -> residual), the reference's
    sources unchanged, once on the reference's own CPU kernels and once on this library: identical fp32 outputs and int8 caches."""
    _need("callsites_cuda")
    _need("ref_modules")
    from oracle import capi

    rng = np.random.default_rng(seed)
    W, B, bo, ln, fc, sc = _opt_params(rng, E, F)
    capi.write_opt_decoder_layer_params(tmp_path, W, B, bo, ln, fc, sc)
    hidden = (rng.standard_normal((prefill + steps, E)) * 40).astype(np.float32)
    got, gk, gv = capi.run_opt_decoder_layer("callsites_cuda", tmp_path, hidden, E, H, F, prefill, steps)
    ref, rk, rv = capi.run_opt_decoder_layer("ref_modules", tmp_path, hidden, E, H, F, prefill, steps)
    assert np.array_equal(gk, rk) and np.array_equal(gv, rv)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
