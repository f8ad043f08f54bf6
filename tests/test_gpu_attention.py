"""Per-token KV-cache attention parity on the GPU (tce_attn_decode) vs the fp32 GQA oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HD = 128


@pytest.fixture(scope="module")
def ctx():
    from tinychatengine_b200.runtime import Context

    c = Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("H,KVH", [(8, 2), (4, 4), (8, 1), (16, 8)])
@pytest.mark.parametrize("past", [0, 1, 5, 127, 128, 129, 300, 1023])
def test_decode_attention_matches_oracle(ctx, H, KVH, past):
    from oracle import capi

    max_ctx = 1024
    rng = np.random.default_rng(1000 * H + past)
    cosb, sinb = capi.rope_tables(max_ctx, HD, 500000.0)
    qkv = rng.standard_normal((H + 2 * KVH) * HD).astype(np.float16)
    pk = (rng.standard_normal((KVH, past, HD)) * 0.7).astype(np.float16)
    pv = rng.standard_normal((KVH, past, HD)).astype(np.float16)
    alpha = 1.0 / np.sqrt(HD)
    q = qkv[: H * HD].astype(np.float32)[None]
    k = qkv[H * HD: (H + KVH) * HD].astype(np.float32)[None]
    v = qkv[(H + KVH) * HD:].astype(np.float32)[None]
    want, fk, fv = capi.llama_attention_core(q, k, v, pk.astype(np.float32) if past else None, pv.astype(np.float32) if past else None,
                                             capi.causal_mask(1, past), cosb, sinb, alpha, H, KVH, HD)
    dev = torch.device("cuda", 0)
    kc = torch.zeros((KVH, max_ctx, HD), dtype=torch.float16, device=dev)
    vc = torch.zeros_like(kc)
    # poison the not-yet-written part of the cache: the kernel must never read it
    kc[:, past:, :] = float("nan")
    vc[:, past:, :] = float("nan")
    if past:
        kc[:, :past] = torch.from_numpy(pk).to(dev)
        vc[:, :past] = torch.from_numpy(pv).to(dev)
    out = torch.zeros(H * HD, dtype=torch.float16, device=dev)
    pos = torch.tensor([past], dtype=torch.int32, device=dev)
    ctx.attn_decode(torch.from_numpy(qkv).to(dev), kc, vc, torch.from_numpy(cosb).to(dev), torch.from_numpy(sinb).to(dev), pos, out, alpha, H, KVH, HD, max_ctx)
    torch.cuda.synchronize()
    got = out.float().cpu().numpy()
    assert np.all(np.isfinite(got))
    scale = max(np.abs(want).max(), 1e-6)
    assert np.abs(got - want[0]).max() / scale <= 2e-3  # fp16 output rounding + fp16 K row
    # the appended rows: rotated K (fp16 rounded) and V, bit-for-bit V
    assert np.allclose(kc[:, past].float().cpu().numpy(), fk[:, past], atol=2e-3, rtol=1e-3)
    assert np.array_equal(vc[:, past].cpu().numpy(), fv[:, past].astype(np.float16))
    if past:
        assert torch.equal(kc[:, :past].cpu(), torch.from_numpy(pk))  # the cached prefix is untouched


def test_multi_step_append_is_consistent(ctx):
    """decode 40 tokens one by one through the in-place cache and compare every step with the oracle fed by its
    own accumulated past (reference test: test_Int4llamaAttention sqlen 9 then 1 with past 9)."""
    from oracle import capi

    H, KVH, max_ctx = 8, 2, 256
    rng = np.random.default_rng(77)
    cosb, sinb = capi.rope_tables(max_ctx, HD, 10000.0)
    dev = torch.device("cuda", 0)
    kc = torch.zeros((KVH, max_ctx, HD), dtype=torch.float16, device=dev)
    vc = torch.zeros_like(kc)
    dcos, dsin = torch.from_numpy(cosb).to(dev), torch.from_numpy(sinb).to(dev)
    out = torch.zeros(H * HD, dtype=torch.float16, device=dev)
    pk = pv = None
    alpha = 1.0 / np.sqrt(HD)
    for step in range(40):
        qkv = rng.standard_normal((H + 2 * KVH) * HD).astype(np.float16)
        pos = torch.tensor([step], dtype=torch.int32, device=dev)
        ctx.attn_decode(torch.from_numpy(qkv).to(dev), kc, vc, dcos, dsin, pos, out, alpha, H, KVH, HD, max_ctx)
        q = qkv[: H * HD].astype(np.float32)[None]
        k = qkv[H * HD: (H + KVH) * HD].astype(np.float32)[None]
        v = qkv[(H + KVH) * HD:].astype(np.float32)[None]
        want, fk, fv = capi.llama_attention_core(q, k, v, pk, pv, capi.causal_mask(1, step), cosb, sinb, alpha, H, KVH, HD)
        fk[:, -1] = fk[:, -1].astype(np.float16).astype(np.float32)
        pk, pv = fk, fv
        got = out.float().cpu().numpy()
        assert np.abs(got - want[0]).max() / max(np.abs(want).max(), 1e-6) <= 3e-3, step
