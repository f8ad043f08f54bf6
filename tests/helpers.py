"""Shared helpers for the parity tests: tolerances stated once, oracle-composed decode step."""
import numpy as np

from oracle import capi

W4_REL_TOL = 1e-2  # north_star: <= 1e-2 relative for fp16 W4A16 (BASELINE.json)


def rel_err(y, ref):
    """max |y - ref| / max |ref|  (the metric of SURVEY.md 8(d) config 1)."""
    y = np.asarray(y, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.max(np.abs(y - ref)) / max(np.max(np.abs(ref)), 1e-30))


def assert_w4_close(y, ref, what=""):
    e = rel_err(y, ref)
    assert np.all(np.isfinite(np.asarray(y, np.float64))), what
    assert e <= W4_REL_TOL, f"{what}: rel err {e:.3e} > {W4_REL_TOL}"
    # the kernels accumulate in fp32 and round once to fp16: anything above ~2e-3 signals a bug long before 1e-2
    assert e <= 2e-3, f"{what}: rel err {e:.3e} is inside the 1e-2 contract but far above fp16 rounding"
    return e


def np_w4(t):
    """(w, zeros, scales) torch tensors -> numpy in oracle dtypes."""
    w, z, s = t
    return w.cpu().numpy().view(np.uint32), z.cpu().numpy().view(np.uint32), s.cpu().numpy()


def oracle_decode_step(model, token, pos, past_k, past_v):
    """One decode step of the synthetic Llama composed from oracle pieces (tce_oracle.c) by oracle/llama_ref.py::llama_forward -- the composition that
    is pinned against the reference's own Int4LlamaForCausalLM (tests/test_oracle_golden.py::test_llama_model_*) -- with the fused GPU path's
    arithmetic plugged in: W4A16 GEMV oracle for the projections and fp16 at the GPU path's rounding points (RMSNorm output, projections, attention
    out, SiLU*mul, the appended key row); fp32 residual and logits.  past_k/past_v: per-layer lists of [KVH, pos, hd] fp32 arrays (or None)."""
    from oracle import llama_ref

    g = model.geom
    cosb, sinb = capi.rope_tables(model.max_ctx, g.head_dim, g.rope_theta)
    assert pos == (0 if past_k[0] is None else past_k[0].shape[1])

    def linear(xh, t):
        w, z, s = np_w4(t)
        return capi.w4a16_gemv(np.asarray(xh).astype(np.float16), w, z, s)

    layers = []
    for l in range(g.num_layers):
        lt = model.layer_tensors(l)
        layers.append({**{n: lt[n] for n in llama_ref.LINEARS}, "input_norm": lt["input_norm"].cpu().numpy(), "post_norm": lt["post_norm"].cpu().numpy()})
    logits, new_k, new_v = llama_ref.llama_forward(
        [token], past_k, past_v, embed_row=lambda t: model.embed[t].float().cpu().numpy(), layers=layers, final_norm=model.final_norm.cpu().numpy(),
        lm_head=model.tensors[-1], linear=linear, cosb=cosb, sinb=sinb, H=g.num_heads, KVH=g.num_kv_heads, hd=g.head_dim, eps=g.rms_eps,
        rnd=lambda a: np.asarray(a).astype(np.float16), round_new_k=lambda a: a.astype(np.float16).astype(np.float32))
    return logits[0], new_k, new_v
