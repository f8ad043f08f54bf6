#!/usr/bin/env python
"""BASELINE.json config 4: Llama-2-13B AWQ-INT4 prompt processing (2048 tokens) on one B200 through tce_llama_prefill.
Reports ms, tokens/s and linear-layer TFLOP/s (2*n*12.688 G + attention 4*T^2*E*L un-masked count).
    python tools/prefill_bench.py [--model llama2-13b] [--n 2048] [--reps 3]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tinychatengine_b200.llama import GEOMETRIES, LlamaModel  # noqa: E402
from tinychatengine_b200.runtime import Context  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama2-13b")
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    g = GEOMETRIES[args.model]
    ctx = Context(0)
    model = LlamaModel(ctx, g, max_ctx=args.n, seed=1)
    toks = [int(t) for t in torch.randint(0, g.vocab_size, (args.n,))]
    model.prefill(toks, 0, None)  # warm-up: allocations, module load
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.reps):
        t0 = time.perf_counter()
        model.prefill(toks, 0, None)  # synchronous call, host tokens in, greedy token out
        ts.append(time.perf_counter() - t0)
    t = min(ts)
    hd = g.head_dim
    per_layer = (g.num_heads * hd + 2 * g.num_kv_heads * hd) * g.embed_dim + g.num_heads * hd * g.embed_dim + 3 * g.hidden_dim * g.embed_dim
    lin_flops = 2.0 * args.n * per_layer * g.num_layers
    attn_flops = 4.0 * args.n * args.n * g.num_heads * hd * g.num_layers
    print(json.dumps({"op": "llama_prefill", "model": args.model, "n": args.n, "ms": round(t * 1e3, 2), "tok_per_s": round(args.n / t, 1),
                      "linear_tflops": round(lin_flops / t / 1e12, 1), "linear_plus_attn_tflops": round((lin_flops + attn_flops) / t / 1e12, 1),
                      "all_ms": [round(x * 1e3, 2) for x in ts]}))


if __name__ == "__main__":
    main()
