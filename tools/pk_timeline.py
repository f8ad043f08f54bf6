#!/usr/bin/env python
"""Phase timeline of the persistent decode kernel (TCE_PK_DEBUG=1 makes it stamp %globaltimer per CTA and phase):
for each phase type, when the barrier opened, how long staging / consuming took, and the achieved HBM rate of the phase.
    TCE_PK_DEBUG=1 python tools/pk_timeline.py [--ctx 2048] [--model llama3-8b] [--layers 32]
"""
import argparse
import json
import os
import sys
from pathlib import Path

os.environ.setdefault("TCE_PK_DEBUG", "1")
import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tinychatengine_b200.llama import GEOMETRIES, LlamaGeometry, LlamaModel, _tensor_from_ptr  # noqa: E402
from tinychatengine_b200.runtime import Context  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ctx", type=int, default=2048)
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    g0 = GEOMETRIES[args.model]
    g = LlamaGeometry(g0.name, args.layers, g0.num_heads, g0.num_kv_heads, g0.embed_dim, g0.hidden_dim, g0.vocab_size, g0.rms_eps, g0.rope_theta)
    ctx = Context(0)
    model = LlamaModel(ctx, g, max_ctx=4096, seed=1)
    for l in range(g.num_layers):
        model.kv_cache(l, 0).normal_(0, 0.5)
        model.kv_cache(l, 1).normal_(0, 0.5)
    ncta = ctx.num_sms
    nphase = 5 * g.num_layers + 1
    tp = torch.tensor([17, args.ctx], dtype=torch.int32, device="cuda")
    for _ in range(4):
        model.decode(tp)
    torch.cuda.synchronize()
    ptr = ctx.L.tce_llama_debug_buffer(model.h, 4)
    assert ptr, "no debug buffer: TCE_PK_DEBUG=1 must be set before the model is created"
    raw = _tensor_from_ptr(ptr, (ncta * nphase * 8 * 2,), torch.int32, 0).cpu().numpy().view(np.uint64).reshape(ncta, nphase, 8).astype(np.float64)
    if args.out:
        np.save(args.out.replace(".json", "") + "_raw.npy", raw)
    t0 = raw[:, 0, 0].min()
    T = (raw - t0) / 1e3  # us
    T[raw == 0] = np.nan
    hd = g.head_dim
    E, F, H, KVH = g.embed_dim, g.hidden_dim, g.num_heads, g.num_kv_heads
    bytes_of = {0: (H + 2 * KVH) * hd * E * 0.5 * 1.0390625, 1: 2 * KVH * hd * 2 * (args.ctx + 1), 2: E * H * hd * 0.5 * 1.0390625, 3: 2 * F * E * 0.5 * 1.0390625,
                4: E * F * 0.5 * 1.0390625}
    names = {0: "qkv", 1: "attn", 2: "o_proj", 3: "gate_up", 4: "down"}
    total = np.nanmax(T[:, -1, 3]) if not np.all(np.isnan(T[:, -1, 3])) else np.nanmax(T)
    print(f"kernel span (first barrier-pass stamp -> last arrival): {total:.1f} us over {nphase} phases, ctx {args.ctx}")
    # attention sub-steps (medians over the CTAs that own chunks and over layers): time since the CTA entered the phase
    att = {4: [], 7: [], 5: [], 6: [], 2: []}
    for p in range(1, nphase - 1):
        if p % 5 == 1:
            for k in att:
                d = T[:, p, k] - T[:, p, 0]
                if not np.all(np.isnan(d)):
                    att[k].append(np.nanmedian(d))
    print("attention, since phase entry (median): q roped+bar %.2f  first stage landed %.2f  chunks done %.2f  partial written %.2f  phase left %.2f us" %
          tuple(float(np.median(att[k])) if att[k] else float("nan") for k in (4, 7, 5, 6, 2)))
    acc = {}
    for p in range(1, nphase - 1):
        k = p % 5
        prev_done = np.nanmax(T[:, p - 1, 3])          # last CTA arrived at the previous barrier
        opened = T[:, p, 0]                             # each CTA saw the barrier open
        staged = T[:, p, 1]
        consumed = T[:, p, 2]
        arrived = T[:, p, 3]
        d = {"barrier_us": np.nanmedian(opened) - prev_done, "stage_us": np.nanmedian(staged - opened) if k != 1 else 0.0,
             "consume_us": np.nanmedian(consumed - (staged if k != 1 else opened)), "tail_us": np.nanmax(arrived) - np.nanmedian(consumed),
             "phase_us": np.nanmax(arrived) - prev_done, "skew_us": np.nanmax(consumed) - np.nanmin(consumed),
             "hop_us": (np.nanmedian(staged) if k != 1 else np.nanmedian(T[:, p, 4])) - prev_done, "etail_us": np.nanmedian(arrived - consumed)}
        acc.setdefault(k, []).append(d)
    summ = {}
    for k, lst in sorted(acc.items()):
        m = {key: float(np.median([d[key] for d in lst])) for key in lst[0]}
        m["GBps"] = bytes_of[k] / m["phase_us"] / 1e3
        m["hbm_floor_us"] = bytes_of[k] / 6.5696e6
        summ[names[k]] = m
        print(f"{names[k]:8s} phase {m['phase_us']:6.2f} us (HBM floor {m['hbm_floor_us']:5.2f})  barrier {m['barrier_us']:5.2f}  stage {m['stage_us']:5.2f}  "
              f"consume {m['consume_us']:6.2f}  tail {m['tail_us']:5.2f}  skew {m['skew_us']:5.2f}  hop {m['hop_us']:5.2f}  etail {m['etail_us']:5.2f}  -> {m['GBps']:7.0f} GB/s")
    lm = nphase - 1
    prev_done = np.nanmax(T[:, lm - 1, 3])
    print(f"lm_head  phase {np.nanmax(T[:, lm, 3]) - prev_done:6.2f} us")
    summ["lm_head_us"] = float(np.nanmax(T[:, lm, 3]) - prev_done)
    summ["kernel_us"] = float(total)
    if args.out:
        Path(args.out).write_text(json.dumps(summ, indent=1))
    model.close()
    ctx.close()


if __name__ == "__main__":
    main()
