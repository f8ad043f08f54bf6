"""Worker for tests/test_dist_gloo.py::test_tensor_parallel_decode_steps_world2_gloo (torchrun, gloo, CPU only): whole decode steps of a synthetic
GQA Llama run tensor-parallel on the host -- the weights cut by tinychatengine_b200.llama.shard_weights exactly as bench.py / the GPU path cut
them (q|k|v, gate|up, lm_head by output rows; o_proj, down_proj by input channels on 128-group boundaries; heads and KV heads by rank), each rank
running oracle/llama_ref.py::llama_forward on its LOCAL geometry with the two all-reduces per layer and the vocabulary gather plugged into
`linear` -- compared on every rank with the unsharded step of the same weights."""
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import capi, llama_ref  # noqa: E402
from tinychatengine_b200.llama import GEOMETRIES, make_random_weights, shard_weights  # noqa: E402


def np_w4(t):
    w, z, s = t
    return w.numpy().view(np.uint32), z.numpy().view(np.uint32), s.numpy()


def gemv(x, t):
    return capi.w4a16_gemv(np.asarray(x).astype(np.float16), *np_w4(t))


def layers_of(W, wrap=lambda name, t: t):
    return [{**{n: wrap(n, L[n]) for n in llama_ref.LINEARS}, "input_norm": L["input_norm"].numpy(), "post_norm": L["post_norm"].numpy()} for L in W["layers"]]


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    g = GEOMETRIES["tiny-gqa"]
    W = make_random_weights(g, torch.device("cpu"), seed=7, random_zeros=True)  # same seed on every rank = the same "checkpoint"
    Wl, gl = shard_weights(W, g, rank, world)
    assert (gl.num_heads, gl.num_kv_heads, gl.hidden_dim, gl.vocab_size) == (g.num_heads // world, g.num_kv_heads // world, g.hidden_dim // world, g.vocab_size // world)
    max_ctx = 16
    cosb, sinb = capi.rope_tables(max_ctx, g.head_dim, g.rope_theta)
    f16 = lambda a: np.asarray(a).astype(np.float16)  # noqa: E731
    k16 = lambda a: a.astype(np.float16).astype(np.float32)  # noqa: E731

    def tp_linear(x, h):
        kind, t = h
        y = torch.from_numpy(np.ascontiguousarray(gemv(x, t)))
        if kind == "row_parallel":  # o_proj / down_proj: partial sums over this rank's input channels -> all-reduce
            dist.all_reduce(y, op=dist.ReduceOp.SUM)
        elif kind == "vocab":  # lm_head: vocabulary shards -> gather
            parts = [torch.empty_like(y) for _ in range(world)]
            dist.all_gather(parts, y)
            y = torch.cat(parts, dim=1)
        return y.numpy()

    kinds = {"o": "row_parallel", "down": "row_parallel"}
    common = dict(cosb=cosb, sinb=sinb, hd=g.head_dim, eps=g.rms_eps, rnd=f16, round_new_k=k16, embed_row=lambda t: W["embed"][t].float().numpy(),
                  final_norm=W["final_norm"].numpy())
    full_layers = layers_of(W)
    tp_layers = layers_of(Wl, lambda n, t: (kinds.get(n, "column_parallel"), t))
    pk, pv = [None] * g.num_layers, [None] * g.num_layers
    qk, qv = [None] * g.num_layers, [None] * g.num_layers
    worst = 0.0
    for tok in (5, 1700, 42):
        want, pk, pv = llama_ref.llama_forward([tok], pk, pv, layers=full_layers, lm_head=W["lm_head"], linear=gemv, H=g.num_heads, KVH=g.num_kv_heads, **common)
        got, qk, qv = llama_ref.llama_forward([tok], qk, qv, layers=tp_layers, lm_head=("vocab", Wl["lm_head"]), linear=tp_linear, H=gl.num_heads,
                                              KVH=gl.num_kv_heads, **common)
        assert got.shape == want.shape == (1, g.vocab_size)
        err = float(np.abs(got - want).max() / np.abs(want).max())
        worst = max(worst, err)
        # fp32 partial sums are added in a different order and the fp16 rounding points may move by one ulp: far below the 1e-2 contract
        assert err <= 5e-3, f"rank {rank}: tensor-parallel step differs from the single-rank step by {err}"
        assert int(got.argmax()) == int(want.argmax())
        # each rank's cache holds exactly its KV heads of the full cache
        lo = rank * gl.num_kv_heads
        for l in range(g.num_layers):
            assert np.abs(qk[l] - pk[l][lo:lo + gl.num_kv_heads]).max() <= 2e-2 and np.abs(qv[l] - pv[l][lo:lo + gl.num_kv_heads]).max() <= 2e-2
    dist.barrier()
    if rank == 0:
        print("GLOO_TP_STEP_OK", world, f"{worst:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
