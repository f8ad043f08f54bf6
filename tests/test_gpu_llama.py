"""Fused decode step (tce_llama_*: one CUDA graph per token) vs the oracle-composed step."""
import numpy as np
import pytest
import torch

from helpers import oracle_decode_step, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mega", ["1", "0"])
@pytest.mark.parametrize("geom", ["tiny-gqa", "tiny-mha"])
def test_decode_steps_match_oracle(geom, mega, monkeypatch):
    """mega=1: one persistent cooperative kernel per token; mega=0: one kernel per op inside a CUDA graph."""
    monkeypatch.setenv("TCE_MEGAKERNEL", mega)
    from tinychatengine_b200.llama import GEOMETRIES, LlamaModel
    from tinychatengine_b200.runtime import Context

    ctx = Context(0)
    g = GEOMETRIES[geom]
    model = LlamaModel(ctx, g, max_ctx=256, seed=7, random_zeros=True)
    past_k = [None] * g.num_layers
    past_v = [None] * g.num_layers
    tokens = [3, 77, 1000, 5, 900, 17, 256, 999]
    logits_host = torch.empty(g.vocab_size, dtype=torch.float32).pin_memory()
    for pos, tok in enumerate(tokens):
        nxt = model.decode_host(tok, pos, logits_host)
        want, past_k, past_v = oracle_decode_step(model, tok, pos, past_k, past_v)
        got = logits_host.numpy().copy()
        assert np.all(np.isfinite(got))
        assert rel_err(got, want) <= 1e-2, (geom, pos, rel_err(got, want))
        assert nxt == int(np.argmax(got))
        # device-resident entry point gives the same logits as the host entry point
        tp = torch.tensor([tok, pos], dtype=torch.int32, device="cuda")
        model.decode(tp)
        torch.cuda.synchronize()
        # (o_proj / down_proj split tiles accumulate with RED.ADD by default: last-bit differences are expected)
        assert rel_err(model.logits().cpu().numpy(), got) <= 2e-3
        for l in range(g.num_layers):
            kc = model.kv_cache(l, 0)[:, : pos + 1].float().cpu().numpy()
            assert np.abs(kc - past_k[l]).max() <= 2e-2 * max(1.0, np.abs(past_k[l]).max())
    model.close()
    ctx.close()


def test_graph_and_eager_paths_agree(monkeypatch):
    import os

    monkeypatch.setenv("TCE_DETERMINISTIC", "1")  # ordered stream-K fix-up instead of RED.ADD: bit-reproducible
    monkeypatch.setenv("TCE_MEGAKERNEL", "0")

    from tinychatengine_b200.llama import GEOMETRIES, LlamaModel
    from tinychatengine_b200.runtime import Context

    outs = []
    for pdl in (1, 0):
        ctx = Context(0)
        ctx.set_option("use_pdl", pdl)
        model = LlamaModel(ctx, GEOMETRIES["tiny-gqa"], max_ctx=128, seed=3)
        lg = torch.empty(model.geom.vocab_size, dtype=torch.float32)
        seq = []
        for pos, tok in enumerate([1, 2, 3, 4]):
            model.decode_host(tok, pos, lg)
            seq.append(lg.clone())
        outs.append(torch.stack(seq))
        model.close()
        ctx.close()
    assert torch.equal(outs[0], outs[1])


def test_persistent_kernel_long_context_matches_graph_path(monkeypatch):
    """many steps through both decode paths on the same weights: the KV caches and logits must stay together (covers
    several attention splits per KV head inside the persistent kernel: chunk 64, ctx up to 200)."""
    from tinychatengine_b200.llama import GEOMETRIES, LlamaModel
    from tinychatengine_b200.runtime import Context

    g = GEOMETRIES["tiny-gqa"]
    outs = {}
    for mega in ("1", "0"):
        monkeypatch.setenv("TCE_MEGAKERNEL", mega)
        ctx = Context(0)
        model = LlamaModel(ctx, g, max_ctx=256, seed=11)
        lg = torch.empty(g.vocab_size, dtype=torch.float32)
        seq = []
        tok = 5
        for pos in range(200):
            tok = model.decode_host(tok % g.vocab_size, pos, lg)
            if pos % 25 == 24 or pos >= 196:
                seq.append(lg.clone())
            tok = (tok * 7 + pos) % g.vocab_size  # data-dependent but identical in both runs unless the paths diverge
        outs[mega] = (torch.stack(seq), model.kv_cache(1, 0)[:, :200].float().cpu().clone())
        model.close()
        ctx.close()
    assert rel_err(outs["1"][0].numpy(), outs["0"][0].numpy()) <= 5e-3
    assert rel_err(outs["1"][1].numpy(), outs["0"][1].numpy()) <= 5e-3
