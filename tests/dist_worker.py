"""Worker for tests/test_dist_gloo.py (run under torchrun, gloo backend, CPU only): the two exchanges of the tensor-parallel decode
on the host -- row-sharded linears (q/k/v/gate/up/lm_head) are gathered, input-channel-sharded linears (o/down) are all-reduced --
computed per rank with the oracle on the rank's own shard of one QM_CUDA tensor, compared on every rank with the unsharded result."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import capi  # noqa: E402
from tinychatengine_b200.llama import shard_w4_cols, shard_w4_rows  # noqa: E402
from tinychatengine_b200.runtime import random_w4  # noqa: E402


def gemv(x16, t):
    w, z, s = t
    return capi.w4a16_gemv(x16, w.numpy().view(np.uint32), z.numpy().view(np.uint32), s.numpy())


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    oc, ic = 64, 1024 * world  # ic/world stays a multiple of the 128-channel group
    t = random_w4(oc, ic, torch.device("cpu"), 9, random_zeros=True)  # same seed on every rank = the same "checkpoint"
    x = torch.randn((1, ic), generator=torch.Generator().manual_seed(3)).to(torch.float16).numpy()
    full = gemv(x, t)[0]
    # output-channel shard -> all-gather
    part = torch.from_numpy(gemv(x, shard_w4_rows(t, rank, world))[0].copy())
    parts = [torch.empty_like(part) for _ in range(world)]
    dist.all_gather(parts, part)
    got = torch.cat(parts).numpy()
    assert np.array_equal(got, full), "row-sharded linear + all-gather differs from the unsharded result"
    # input-channel shard -> all-reduce(sum)
    lo, hi = rank * ic // world, (rank + 1) * ic // world
    partial = torch.from_numpy(gemv(np.ascontiguousarray(x[:, lo:hi]), shard_w4_cols(t, ic, rank, world))[0].copy())
    dist.all_reduce(partial, op=dist.ReduceOp.SUM)
    err = np.abs(partial.numpy() - full).max() / np.abs(full).max()
    assert err <= 1e-5, f"column-sharded linear + all-reduce: rel err {err}"
    # greedy token over vocabulary shards: (value, global index) max-reduce picks the global arg-max, ties to the lowest index
    logits = torch.from_numpy(full.copy())
    sh = logits[rank * oc // world:(rank + 1) * oc // world]
    key = torch.tensor([float(sh.max()), -float(rank * oc // world + int(sh.argmax()))], dtype=torch.float64)
    keys = [torch.empty_like(key) for _ in range(world)]
    dist.all_gather(keys, key)
    best = max((k.tolist() for k in keys))
    assert int(-best[1]) == int(logits.argmax()), "sharded arg-max"
    dist.barrier()
    if rank == 0:
        print("GLOO_TP_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
