"""Host-side tensor-parallel sharding of the QM_CUDA format (tinychatengine_b200/llama.py): an input-channel shard
re-packed as its own tensor must dequantise to the same columns as the full tensor."""
import numpy as np
import torch

from oracle import quant
from tinychatengine_b200.llama import shard_w4_cols, shard_w4_rows
from tinychatengine_b200.runtime import random_w4


def test_column_and_row_shards_dequantise_to_slices():
    oc, ic, P = 32, 14336 // 4, 4  # 28 groups -> 7 per shard (zeros_w 4 -> 1, nibble-level repack)
    t = random_w4(oc, ic, torch.device("cpu"), 5, random_zeros=True)
    full = quant.dequant_qmcuda(t[0].numpy().view(np.uint32), t[2].numpy(), t[1].numpy().view(np.uint32))
    for r in range(P):
        w, z, s = shard_w4_cols(t, ic, r, P)
        assert w.shape == (oc, ic // P // 8) and z.shape == (oc, 1) and s.shape == (oc, 8)
        part = quant.dequant_qmcuda(w.numpy().view(np.uint32), s.numpy(), z.numpy().view(np.uint32))
        assert np.array_equal(part, full[:, r * ic // P:(r + 1) * ic // P])
        w, z, s = shard_w4_rows(t, r, P)
        part = quant.dequant_qmcuda(w.numpy().view(np.uint32), s.numpy(), z.numpy().view(np.uint32))
        assert np.array_equal(part, full[r * oc // P:(r + 1) * oc // P])


def test_tp_rank_keeps_the_full_embedding_table():
    """A tensor-parallel rank shards lm_head over the vocabulary but looks up GLOBAL token ids: synthetic local weights must carry
    an embedding table with all vocab_size * P rows (a local-size table made an 8-GPU run read far out of bounds)."""
    from tinychatengine_b200.llama import GEOMETRIES, LlamaGeometry, make_random_weights, shard_weights

    g = GEOMETRIES["tiny-gqa"]
    P = 2
    gl = LlamaGeometry(g.name, 1, g.num_heads // P, g.num_kv_heads // P, g.embed_dim, g.hidden_dim // P, g.vocab_size // P, g.rms_eps, g.rope_theta)
    W = make_random_weights(gl, torch.device("cpu"), seed=3, embed_rows=gl.vocab_size * P)
    assert W["embed"].shape == (g.vocab_size, g.embed_dim) and W["lm_head"][0].shape[0] == g.vocab_size // P
    full = make_random_weights(LlamaGeometry(g.name, 1, g.num_heads, g.num_kv_heads, g.embed_dim, g.hidden_dim, g.vocab_size), torch.device("cpu"), seed=3)
    Wl, gl2 = shard_weights(full, LlamaGeometry(g.name, 1, g.num_heads, g.num_kv_heads, g.embed_dim, g.hidden_dim, g.vocab_size), 1, P)
    assert Wl["embed"].shape[0] == g.vocab_size and gl2.vocab_size == g.vocab_size // P
