#!/usr/bin/env python
import os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tinychatengine_b200.llama import GEOMETRIES, LlamaGeometry, LlamaModel
from tinychatengine_b200.runtime import Context
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))
from debug_mega import run

g0 = GEOMETRIES["tiny-mha"]
g = LlamaGeometry(g0.name, 1, g0.num_heads, g0.num_kv_heads, g0.embed_dim, g0.hidden_dim, g0.vocab_size, g0.rms_eps, g0.rope_theta)
for det in ("0", "1"):
    if det == "1":
        os.environ["TCE_DETERMINISTIC"] = "1"
    a = run(g, "1", [3, 77]); b = run(g, "0", [3, 77])
    for k in ("act", "resid"):
        d = (a[1][k] - b[1][k]).abs()
        bad = (d > 1e-3 * b[1][k].abs().max()).nonzero().flatten()
        print(f"det={det} {k}: {len(bad)} bad of {d.numel()}; first {bad[:24].tolist()} last {bad[-8:].tolist()}", flush=True)
        if len(bad):
            i = bad[0].item()
            print("   mega", a[1][k][i:i+8].tolist(), "\n   graph", b[1][k][i:i+8].tolist())
