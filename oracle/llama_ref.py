"""The Llama forward pass COMPOSED from oracle pieces, and its pin against the reference's own model.  TEST INFRASTRUCTURE ONLY.

``llama_forward`` is the one statement of how the pieces of a Llama step follow each other -- embedding row, RMSNorm, q/k/v projections,
attention core (RoPE, GQA, softmax), o_proj, residual add, RMSNorm, gate/up, SiLU(gate)*up, down_proj, residual add, final norm, lm_head --
restating ``Int4LlamaForCausalLM::forward`` (llm/src/nn_modules/non_cuda/Int4llamaForCausalLM.cc:17-56), ``Int4llamaDecoder::forward``
(non_cuda/Int4llamaDecoder.cc:66-131) and ``Int4llamaDecoderLayer::forward`` (non_cuda/Int4llamaDecoderLayer.cc:48-114).  It is parameterised by

* ``linear(x, handle)``: the projection arithmetic -- ``w4a8_linear`` below (the reference's CPU build) when the composition is PINNED against the
  compiled reference model (oracle/_ref/libtce_ref_llama_model.so, tests/golden/llama_model.npz), the W4A16 GEMV oracle when the GPU decode step is
  checked (tests/helpers.py::oracle_decode_step) -- and
* ``rnd(x)``: the rounding applied at the points where the GPU path holds fp16 (identity for the fp32 CPU reference).

Both users run the same Python lines, so the pinned composition IS the one the GPU parity tests compare with.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi, quant

LINEARS = ("q", "k", "v", "o", "gate", "up", "down")


def _identity(x):
    return x


def silu_mul(gate, up):
    """SiLuMul (non_cuda/Int4llamaDecoderLayer.cc:32-46): a = a / (1 + exp(-a)) * b in fp32."""
    gate = np.asarray(gate, np.float32)
    return (gate / (np.float32(1.0) + np.exp(-gate)) * np.asarray(up, np.float32)).astype(np.float32)


def llama_forward(tokens, past_k, past_v, *, embed_row, layers, final_norm, lm_head, linear, cosb, sinb, H, KVH, hd, eps, rnd=_identity,
                  round_new_k=None):
    """One call of the model on ``tokens`` (a prompt pass when len > 1, a decode step when 1) after ``past`` cached positions.

    embed_row(token) -> fp32 [E]; layers: list of dicts with "input_norm", "post_norm" (fp32 [E]) and one opaque handle per name in LINEARS;
    lm_head: handle; past_k / past_v: per-layer lists of fp32 [KVH, past, hd] arrays (or None).  ``round_new_k``: optional rounding of the
    freshly appended key rows (the GPU cache holds fp16).  Returns (logits fp32 [T, vocab], new_k, new_v)."""
    T = len(tokens)
    past = 0 if past_k[0] is None else past_k[0].shape[1]
    x = np.stack([np.asarray(embed_row(t), np.float32) for t in tokens]).astype(np.float32)  # residual stream fp32 [T, E]
    mask = capi.causal_mask(T, past)
    alpha = 1.0 / np.sqrt(hd)
    new_k, new_v = [], []
    for l, lt in enumerate(layers):
        xn = rnd(capi.rmsnorm(x, lt["input_norm"], eps))
        q, k, v = (rnd(linear(xn, lt[n])).astype(np.float32) for n in ("q", "k", "v"))
        core, fk, fv = capi.llama_attention_core(q, k, v, past_k[l], past_v[l], mask, cosb, sinb, alpha, H, KVH, hd)
        if round_new_k is not None:
            fk[:, past:, :] = round_new_k(fk[:, past:, :])
        new_k.append(fk)
        new_v.append(fv)
        x = x + linear(rnd(core), lt["o"])  # residual add (Int4llamaDecoderLayer.cc:76-79)
        xn = rnd(capi.rmsnorm(x, lt["post_norm"], eps))
        act = rnd(silu_mul(linear(xn, lt["gate"]), linear(xn, lt["up"])))
        x = x + linear(act, lt["down"])  # residual add (:111)
    xn = rnd(capi.rmsnorm(x, final_norm, eps))
    return linear(xn, lm_head), new_k, new_v


# ----------------------------------------------------------------------------------------------
# the reference CPU build's projection: Linear_FP_int4::forward -> mat_mul_accelerator_int8_int4_fast_no_offset (QM_x86, W4A8)
# ----------------------------------------------------------------------------------------------
def unpack_q4_3(qs: np.ndarray) -> np.ndarray:
    """uint8 [OC, IC/2] in the QM_x86 layout (byte e of a 64-weight pair of blocks = w[e] | w[32+e] << 4) -> int8 [OC, IC] of (q - 8)."""
    oc = qs.shape[0]
    b = qs.reshape(oc, -1, 32)
    q = np.concatenate([b & 0xF, b >> 4], axis=2)  # [OC, IC/64, 64]
    return (q.astype(np.int16) - 8).astype(np.int8).reshape(oc, -1)


def w4a8_linear(x, handle):
    """fp32 [M, OC] = the reference's W4A8 product (kernels/avx/matmul_avx_int8_int4.cc:179-256,259-350): activations quantised per 32-block to
    int8 (d = amax/127, round to nearest even), exact integer dot per block, one fp32 multiply by (s_w * s_a) per block, fp32 sum over blocks
    (summation order differs from the 8-lane FMA chain of the AVX code: a few 1e-7 relative).  handle = (Wq int8 [OC, IC], S fp32 [OC, IC/32])."""
    Wq, S = handle
    x = np.ascontiguousarray(x, np.float32)
    M, IC = x.shape
    xb = x.reshape(M, IC // 32, 32)
    amax = np.abs(xb).max(2)
    d = (amax / np.float32(127)).astype(np.float32)
    inv = np.where(amax != 0, np.float32(127) / np.where(amax != 0, amax, 1), 0).astype(np.float32)
    x8 = np.rint((xb * inv[:, :, None]).astype(np.float32)).astype(np.int32)
    dots = np.einsum("mbk,obk->mob", x8, Wq.reshape(Wq.shape[0], IC // 32, 32).astype(np.int32)).astype(np.float32)  # exact: |dot| < 2^24
    return (dots * (S[None, :, :] * d[:, None, :]).astype(np.float32)).sum(2, dtype=np.float64).astype(np.float32)


def random_model(rng, E, H, KVH, L, F, vocab):
    """fp32 master weights of a synthetic Llama; the QM_x86 quantisation happens in write_llama_model_params / quantized_handles."""
    hd = E // H

    def w(oc, ic):
        return (rng.standard_normal((oc, ic)) / np.sqrt(ic)).astype(np.float32)

    layers = []
    for _ in range(L):
        lt = {"input_norm": (1 + 0.1 * rng.standard_normal(E)).astype(np.float32), "post_norm": (1 + 0.1 * rng.standard_normal(E)).astype(np.float32)}
        lt.update(q=w(E, E), k=w(KVH * hd, E), v=w(KVH * hd, E), o=w(E, E), gate=w(F, E), up=w(F, E), down=w(E, F))
        layers.append(lt)
    return {"embed": rng.standard_normal((vocab, E)).astype(np.float32), "layers": layers,
            "final_norm": (1 + 0.1 * rng.standard_normal(E)).astype(np.float32), "lm_head": w(vocab, E)}


def _write_linear(d, w):
    os.makedirs(d, exist_ok=True)
    qs, sc = quant.quantize_q4_3(w)
    qs.tofile(os.path.join(d, "weight_int4.bin"))
    sc.astype(np.float32).tofile(os.path.join(d, "scaling_factor_int4.bin"))
    np.array([8.0], np.float32).tofile(os.path.join(d, "zero_point_int4.bin"))
    return unpack_q4_3(qs), sc.astype(np.float32)


def _write_f32(d, name, a):
    os.makedirs(d, exist_ok=True)
    np.ascontiguousarray(a, np.float32).tofile(os.path.join(d, name))


def write_llama_model_params(root, model, cosb, sinb, alpha):
    """Writes the parameter tree the reference's CPU Int4LlamaForCausalLM loads (the layout model_quantizer.py produces for QM_x86) and returns the
    same weights as w4a8_linear handles: {"layers": [...], "lm_head": handle}."""
    root = str(root)
    dec = os.path.join(root, "decoder")
    _write_f32(os.path.join(dec, "embed_tokens"), "weight.bin", model["embed"])
    _write_f32(os.path.join(dec, "norm"), "weight.bin", model["final_norm"])
    layers = []
    for l, lt in enumerate(model["layers"]):
        ld = os.path.join(dec, f"layer{l}")
        _write_f32(os.path.join(ld, "input_layernorm"), "weight.bin", lt["input_norm"])
        _write_f32(os.path.join(ld, "post_attention_layernorm"), "weight.bin", lt["post_norm"])
        h = {"input_norm": lt["input_norm"], "post_norm": lt["post_norm"]}
        attn = os.path.join(ld, "self_attn")
        for n in ("q", "k", "v", "o"):
            h[n] = _write_linear(os.path.join(attn, n + "_proj"), lt[n])
        _write_f32(os.path.join(attn, "rotary_emb"), "cos_cached.bin", cosb)
        _write_f32(os.path.join(attn, "rotary_emb"), "sin_cached.bin", sinb)
        _write_f32(os.path.join(attn, "qk_bmm"), "alpha.bin", np.array([alpha], np.float32))
        for n in ("gate", "up", "down"):
            h[n] = _write_linear(os.path.join(ld, n + "_proj"), lt[n])
        layers.append(h)
    return {"layers": layers, "lm_head": _write_linear(os.path.join(root, "lm_head"), model["lm_head"])}


def quantized_handles(model):
    """The handles write_llama_model_params returns, without touching the disk."""
    def h(w):
        qs, sc = quant.quantize_q4_3(w)
        return unpack_q4_3(qs), sc.astype(np.float32)

    layers = [{**{n: h(lt[n]) for n in LINEARS}, "input_norm": lt["input_norm"], "post_norm": lt["post_norm"]} for lt in model["layers"]]
    return {"layers": layers, "lm_head": h(model["lm_head"])}


def ref_model_available() -> bool:
    return (capi.REF_DIR / "libtce_ref_llama_model.so").exists()


def ref_int4_llama_causal_lm(param_root, tokens, E, H, KVH, L, F, vocab, prefill, decode_steps, max_sqlen=640, eps=1e-5):
    """Runs the REFERENCE Int4LlamaForCausalLM::forward (CPU): a prompt pass over tokens[:prefill], then one call per remaining token with the returned
    past keys / values.  Returns logits fp32 [prefill + decode_steps, vocab]."""
    so = capi.REF_DIR / "libtce_ref_llama_model.so"
    if not so.exists():
        raise FileNotFoundError(f"{so} not built (needs /root/reference; run `make -C oracle ref`)")
    lib = C.CDLL(str(so))
    lib.ref_int4_llama_causal_lm.argtypes = [C.c_char_p] + [C.c_int] * 7 + [C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    tok = np.ascontiguousarray(tokens, np.int32)
    T = prefill + decode_steps
    assert tok.size == T
    logits = np.zeros((T, vocab), np.float32)
    n = lib.ref_int4_llama_causal_lm(str(param_root).encode(), E, H, KVH, L, F, vocab, max_sqlen, eps, tok.ctypes.data, prefill, decode_steps,
                                     logits.ctypes.data)
    assert n == T, n
    return logits


def oracle_int4_llama_causal_lm(model, handles, tokens, cosb, sinb, H, KVH, prefill, decode_steps, eps=1e-5):
    """The same calls through llama_forward with the reference CPU build's arithmetic (fp32 everywhere, W4A8 projections)."""
    E = model["embed"].shape[1]
    L = len(handles["layers"])
    pk, pv = [None] * L, [None] * L
    out, row = [], 0
    for call in range(1 + decode_steps):
        s = prefill if call == 0 else 1
        lg, pk, pv = llama_forward(list(tokens[row:row + s]), pk, pv, embed_row=lambda t: model["embed"][t], layers=handles["layers"],
                                   final_norm=model["final_norm"], lm_head=handles["lm_head"], linear=w4a8_linear, cosb=cosb, sinb=sinb, H=H, KVH=KVH,
                                   hd=E // H, eps=eps)
        out.append(lg)
        row += s
    return np.concatenate(out)
