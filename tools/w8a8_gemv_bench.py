#!/usr/bin/env python
"""W8A8 at decode shapes (M = 1, BASELINE config 3: Llama-2-7B shapes): the DP4A kernel is HBM-bound, report achieved GB/s.
Weights rotate over > 2x L2 so that no launch finds its matrix cached; launches replayed from a CUDA graph.
    python tools/w8a8_gemv_bench.py [--reps 20] [--m 1]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tinychatengine_b200.runtime import Context  # noqa: E402

SHAPES = {"qkv/o 4096x4096": (4096, 4096), "fc1 11008x4096": (11008, 4096), "fc2 4096x11008": (4096, 11008), "lm_head 32000x4096": (32000, 4096)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--m", type=int, default=1)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    ctx = Context(0)
    peak = json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
    for name, (n, k) in SHAPES.items():
        nbuf = max(2, int(300e6 // (n * k)) + 1)
        Bs = [torch.randint(-127, 128, (n, k), dtype=torch.int8, device=dev) for _ in range(nbuf)]
        A = torch.randint(-127, 128, (args.m, k), dtype=torch.int8, device=dev)
        b8 = torch.randint(-127, 128, (n,), dtype=torch.int8, device=dev)
        out = torch.empty((args.m, n), dtype=torch.int8, device=dev)
        side = torch.cuda.Stream()
        ctx.set_stream(side)
        with torch.cuda.stream(side):
            for i in range(3):
                ctx.w8a8_matmul(0, A, Bs[i % nbuf], b8, 0.0005, 0.02, out=out)
        side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for i in range(args.reps):
                ctx.w8a8_matmul(0, A, Bs[i % nbuf], b8, 0.0005, 0.02, out=out)
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / args.reps
        gbs = n * k / us / 1e3
        print(json.dumps({"op": "w8a8 dp4a", "shape": name, "M": args.m, "us": round(us, 2), "GB/s": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak, 3)}), flush=True)
        del Bs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
