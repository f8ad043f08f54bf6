#!/usr/bin/env python
"""Compare the persistent decode kernel with the kernel-per-op graph path buffer by buffer (1-layer tiny models)."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tinychatengine_b200.llama import GEOMETRIES, LlamaGeometry, LlamaModel  # noqa: E402
from tinychatengine_b200.runtime import Context  # noqa: E402


def run(geom, mega, steps):
    os.environ["TCE_MEGAKERNEL"] = mega
    ctx = Context(0)
    model = LlamaModel(ctx, geom, max_ctx=256, seed=7, random_zeros=True)
    lg = torch.empty(geom.vocab_size, dtype=torch.float32)
    out = []
    for pos, tok in enumerate(steps):
        model.decode_host(tok, pos, lg)
        out.append({"logits": lg.clone(), **{n: model.debug_buffer(i).float().cpu().clone() for i, n in enumerate(["resid", "qkv", "attn", "act"])}})
    model.close()
    ctx.close()
    return out


def main():
    for base in ("tiny-mha", "tiny-gqa"):
        g0 = GEOMETRIES[base]
        for layers in (1, 2):
            g = LlamaGeometry(g0.name, layers, g0.num_heads, g0.num_kv_heads, g0.embed_dim, g0.hidden_dim, g0.vocab_size, g0.rms_eps, g0.rope_theta)
            a = run(g, "1", [3, 77, 5])
            b = run(g, "0", [3, 77, 5])
            for pos in range(3):
                msg = []
                for k in ("qkv", "attn", "act", "resid", "logits"):
                    d = (a[pos][k] - b[pos][k]).abs().max().item() / max(b[pos][k].abs().max().item(), 1e-9)
                    msg.append(f"{k}={d:.2e}")
                print(f"{base} layers={layers} pos={pos}: " + " ".join(msg), flush=True)


if __name__ == "__main__":
    main()
