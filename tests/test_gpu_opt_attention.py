"""Int8OPTAttention core on the GPU (tce_opt_int8_attention): BIT-EXACT against the oracle restatement of
llm/src/nn_modules/Int8OPTAttention.cc:183-284, in the reference's copy mode (past -> fresh [H][tgz][hd] buffers) and in the
in-place cache mode, with the explicit mask tensor and with the built-in causal mask."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

QK_ALPHA, PV_ALPHA = 0.0007, 0.011


@pytest.fixture(scope="module")
def ctx():
    from tinychatengine_b200.runtime import Context

    c = Context(0)
    yield c
    c.close()


def rnd8(shape, seed):
    return np.random.default_rng(seed).integers(-127, 128, shape, dtype=np.int8)


@pytest.mark.parametrize("H,hd,sqlen,past", [(12, 64, 9, 0), (12, 64, 1, 9), (12, 64, 1, 300), (32, 64, 5, 17), (4, 128, 33, 0), (2, 32, 1, 2047), (3, 20, 2, 5)])
@pytest.mark.parametrize("explicit_mask", [True, False])
def test_copy_mode_bit_exact(ctx, H, hd, sqlen, past, explicit_mask):
    from oracle import capi

    E, tgz = H * hd, past + sqlen
    q, k, v = rnd8((sqlen, E), 1), rnd8((sqlen, E), 2), rnd8((sqlen, E), 3)
    pk = rnd8((H, past, hd), 4) if past else None
    pv = rnd8((H, past, hd), 5) if past else None
    mask = capi.causal_mask(sqlen, past)
    want, fk, fv = capi.opt_int8_attention_core(q, k, v, pk, pv, mask, QK_ALPHA, PV_ALPHA, H, hd)
    dev = lambda a: None if a is None else torch.from_numpy(a).cuda()
    final_k = torch.zeros((H, tgz, hd), dtype=torch.int8, device="cuda")
    final_v = torch.zeros_like(final_k)
    got = ctx.opt_int8_attention(dev(q), dev(k), dev(v), dev(pk), dev(pv), final_k, final_v, dev(mask) if explicit_mask else None, QK_ALPHA, PV_ALPHA,
                                 past, H, hd)
    assert np.array_equal(final_k.cpu().numpy(), fk) and np.array_equal(final_v.cpu().numpy(), fv)
    assert np.array_equal(got.cpu().numpy(), want)


def test_in_place_cache_decode_sequence(ctx):
    """Prefill 6 tokens, then 5 single-token steps into one preallocated [H][max_ctx][hd] cache."""
    from oracle import capi

    H, hd, max_ctx = 12, 64, 32
    E = H * hd
    ck = torch.zeros((H, max_ctx, hd), dtype=torch.int8, device="cuda")
    cv = torch.zeros_like(ck)
    pk = pv = None
    past = 0
    for step in range(6):
        sqlen = 6 if step == 0 else 1
        q, k, v = rnd8((sqlen, E), 10 + step), rnd8((sqlen, E), 20 + step), rnd8((sqlen, E), 30 + step)
        want, pk, pv = capi.opt_int8_attention_core(q, k, v, pk, pv, capi.causal_mask(sqlen, past), QK_ALPHA, PV_ALPHA, H, hd)
        got = ctx.opt_int8_attention(torch.from_numpy(q).cuda(), torch.from_numpy(k).cuda(), torch.from_numpy(v).cuda(), ck, cv, ck, cv, None, QK_ALPHA,
                                     PV_ALPHA, past, H, hd)
        assert np.array_equal(got.cpu().numpy(), want), f"step {step}"
        past += sqlen
        assert np.array_equal(ck[:, :past].cpu().numpy(), pk) and np.array_equal(cv[:, :past].cpu().numpy(), pv)


def test_rejects_bad_arguments(ctx):
    from tinychatengine_b200 import _lib

    q = torch.zeros((1, 64), dtype=torch.int8, device="cuda")
    ck = torch.zeros((1, 4, 64), dtype=torch.int8, device="cuda")
    with pytest.raises(_lib.TceError):
        ctx.opt_int8_attention(q, q, q, None, None, ck, ck, None, 1.0, 1.0, 4, 1, 64)  # past > 0 without a past cache
    with pytest.raises(_lib.TceError):
        ctx.opt_int8_attention(q, q, q, ck, ck, ck, ck, None, 1.0, 1.0, 4, 1, 64)  # final stride too small for past + sqlen


def test_module_golden_fixture(ctx, golden_dir):
    """Projections (tce_w8a8_matmul) + core (tce_opt_int8_attention, in-place cache) + out_proj on the GPU == the outputs the compiled
    reference Int8OPTAttention module produced (tests/golden/opt_attention_module.npz), bit for bit."""
    g = np.load(golden_dir / "opt_attention_module.npz")
    H, prefill, steps = int(g["H"]), int(g["prefill"]), int(g["steps"])
    E = g["hidden"].shape[1]
    hd = E // H
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    W = {k: dev(g["w" + k]) for k in "qkvo"}
    B = {k: dev(g["b" + k]) for k in "qkv"}
    bo = dev(g["bo"])
    a_qkv, b_qkv, qk_alpha, pv_alpha, a_out = (float(g[k]) for k in ("a_qkv", "b_qkv", "qk_alpha", "pv_alpha", "a_out"))
    ck = torch.zeros((H, 64, hd), dtype=torch.int8, device="cuda")
    cv = torch.zeros_like(ck)
    hidden = dev(g["hidden"])
    outs, past, row = [], 0, 0
    for call in range(1 + steps):
        s = prefill if call == 0 else 1
        x = hidden[row:row + s].contiguous()
        q, k, v = (ctx.w8a8_matmul(0, x, W[n], B[n], a_qkv, b_qkv) for n in "qkv")
        core = ctx.opt_int8_attention(q, k, v, ck, cv, ck, cv, None, qk_alpha, pv_alpha, past, H, hd)
        outs.append(ctx.w8a8_matmul(2, core, W["o"], bo, a_out, 0.0))
        past += s
        row += s
    got = torch.cat(outs).cpu().numpy()
    assert np.array_equal(ck[:, :past].cpu().numpy(), g["final_k"]) and np.array_equal(cv[:, :past].cpu().numpy(), g["final_v"])
    assert np.array_equal(got.view(np.uint32), g["out"].view(np.uint32))
