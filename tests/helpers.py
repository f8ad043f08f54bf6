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
    """One decode step of the synthetic Llama composed from oracle pieces (tce_oracle.c), mirroring the fused
    GPU path's rounding points: fp32 residual, RMSNorm output -> fp16, projections -> fp16, attention out -> fp16,
    SiLU*mul -> fp16, logits fp32.  past_k/past_v: per-layer lists of [KVH, pos, hd] fp32 arrays (or None)."""
    g = model.geom
    hd, H, KVH = g.head_dim, g.num_heads, g.num_kv_heads
    cosb, sinb = capi.rope_tables(model.max_ctx, hd, g.rope_theta)
    x = model.embed[token].float().cpu().numpy()[None, :].astype(np.float32)  # resid fp32 [1,E]
    new_k, new_v = [], []
    alpha = 1.0 / np.sqrt(hd)

    def gemv(xh, name, l=None):
        t = model.layer_tensors(l)[name] if l is not None else model.tensors[-1]
        w, z, s = np_w4(t)
        return capi.w4a16_gemv(xh.astype(np.float16), w, z, s)

    for l in range(g.num_layers):
        lt = model.layer_tensors(l)
        xn = capi.rmsnorm(x, lt["input_norm"].cpu().numpy(), g.rms_eps).astype(np.float16)
        q = gemv(xn, "q", l).astype(np.float16).astype(np.float32)
        k = gemv(xn, "k", l).astype(np.float16).astype(np.float32)
        v = gemv(xn, "v", l).astype(np.float16).astype(np.float32)
        mask = capi.causal_mask(1, pos)
        out, fk, fv = capi.llama_attention_core(q, k, v, past_k[l], past_v[l], mask, cosb, sinb, alpha, H, KVH, hd)
        # the cache holds fp16: round the appended row the way the kernel stores it
        fk[:, -1, :] = fk[:, -1, :].astype(np.float16).astype(np.float32)
        new_k.append(fk)
        new_v.append(fv)
        o = gemv(out.astype(np.float16), "o", l)
        x = x + o
        xn = capi.rmsnorm(x, lt["post_norm"].cpu().numpy(), g.rms_eps).astype(np.float16)
        gate = gemv(xn, "gate", l)
        up = gemv(xn, "up", l)
        act = (gate / (1.0 + np.exp(-gate)) * up).astype(np.float16)
        x = x + gemv(act, "down", l)
    xn = capi.rmsnorm(x, model.final_norm.cpu().numpy(), g.rms_eps).astype(np.float16)
    logits = gemv(xn, None, None)
    return logits[0], new_k, new_v
