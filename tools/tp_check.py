#!/usr/bin/env python
"""Tensor-parallel decode vs the single-GPU path on the same weights (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/tp_check.py
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tinychatengine_b200.llama import GEOMETRIES, LlamaModel, make_random_weights, shard_weights  # noqa: E402
from tinychatengine_b200.runtime import Context  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    name = sys.argv[1] if len(sys.argv) > 1 else "tiny-gqa"
    g = GEOMETRIES[name]
    if len(sys.argv) > 2:  # optional: keep only the first N layers (full-width geometry, short model)
        from tinychatengine_b200.llama import LlamaGeometry

        g = LlamaGeometry(g.name, int(sys.argv[2]), g.num_heads, g.num_kv_heads, g.embed_dim, g.hidden_dim, g.vocab_size, g.rms_eps, g.rope_theta)
    if g.num_kv_heads % world or (g.hidden_dim // world) % 128:
        raise SystemExit(f"{name}: does not split over {world} ranks on 128-channel group boundaries")
    ctx = Context(local)
    W = make_random_weights(g, dev, seed=21, random_zeros=True)
    Wl, gl = shard_weights(W, g, rank, world)
    model = LlamaModel(ctx, gl, max_ctx=256, weights=Wl, tp_rank=rank, tp_size=world)
    model.tp_connect()
    if rank == 0:
        print(f"TP_CHECK kernels per step: {model.kernels_per_step} (1 = persistent kernel with NVLink peer stores)", flush=True)
    ref = None
    if rank == 0:
        os.environ["TCE_PERSISTENT"] = "0"  # the single-GPU reference runs one kernel per op
        ref = LlamaModel(ctx, g, max_ctx=256, weights=W)
    lg_local = torch.empty(gl.vocab_size, dtype=torch.float32)
    lg_ref = torch.empty(g.vocab_size, dtype=torch.float32)
    tok, worst, ok = 3, 0.0, True
    for pos in range(40):
        nxt = model.decode_host(tok, pos, lg_local)
        shards = [torch.empty(gl.vocab_size, dtype=torch.float32, device=dev) for _ in range(world)]
        dist.all_gather(shards, lg_local.to(dev))
        full = torch.cat(shards).cpu()
        nxt_all = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(world)]
        dist.all_gather(nxt_all, torch.tensor([nxt], dtype=torch.int32, device=dev))
        agree = all(int(t) == nxt for t in nxt_all)
        if rank == 0:
            if not agree:
                ok = False
                print(f"pos {pos}: ranks disagree on the greedy token: {[int(t) for t in nxt_all]}", flush=True)
            nref = ref.decode_host(tok, pos, lg_ref)
            err = (full - lg_ref).abs().max().item() / lg_ref.abs().max().item()
            worst = max(worst, err)
            if err > 5e-3 or int(torch.argmax(full)) != nxt:
                ok = False
                print(f"pos {pos}: rel err {err:.3e} next tp={nxt} ref={nref} argmax(full)={int(torch.argmax(full))}", flush=True)
        tok = (nxt * 7 + pos) % g.vocab_size
    if rank == 0:
        print(f"TP_CHECK {name} world={world}: worst rel err {worst:.3e} -> {'OK' if ok else 'FAIL'}", flush=True)
    model.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
