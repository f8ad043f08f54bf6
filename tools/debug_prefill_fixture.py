"""debug: prefill K rows on the reference-module fixture inputs"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import capi  # noqa: E402
from tinychatengine_b200.runtime import Context  # noqa: E402

g = np.load(Path(__file__).resolve().parents[1] / "tests/golden/llama_attention_module.npz")
H, KVH, prefill, steps, max_sq = (int(g[k]) for k in ("H", "KVH", "prefill", "steps", "max_sq"))
HD = 128
hidden = g["hidden"]
E = hidden.shape[1]
cosb, sinb = capi.rope_tables(max_sq, HD, float(g["theta"]))
dev = torch.device("cuda", 0)
ctx = Context(0)
dcos, dsin = torch.from_numpy(cosb).to(dev), torch.from_numpy(sinb).to(dev)
qkv_all = np.concatenate([hidden[:, g["sel_q"]], hidden[:, g["sel_k"]], hidden[:, g["sel_v"]]], axis=1).astype(np.float16)
for variant in ("held", "temp"):
    kc = torch.zeros((KVH, max_sq, HD), dtype=torch.float16, device=dev)
    vc = torch.zeros_like(kc)
    outs = torch.zeros((prefill + steps, E), dtype=torch.float16, device=dev)
    if variant == "held":
        dq = torch.from_numpy(qkv_all[:prefill].copy()).to(dev)
        ctx.attn_prefill(dq, kc, vc, dcos, dsin, outs[:prefill], float(g["alpha"]), prefill, 0, H, KVH, HD, max_sq)
    else:
        ctx.attn_prefill(torch.from_numpy(qkv_all[:prefill]).to(dev), kc, vc, dcos, dsin, outs[:prefill], float(g["alpha"]), prefill, 0, H, KVH, HD, max_sq)
    torch.cuda.synchronize()
    K = kc[:, :prefill].float().cpu().numpy()
    ref = g["final_k"][:, :prefill]
    print(variant, "err per pos", np.abs(K - ref).max(axis=(0, 2)))
    print(" gpu K[0,0,:8]", K[0, 0, :8], "\n ref K[0,0,:8]", ref[0, 0, :8], "\n in  k[0,:8]  ", qkv_all[0, H * HD:H * HD + 8])
    # does GPU row 0 equal some other token's k / some other slice of the input row?
    for i in range(prefill):
        for off in range(0, (H + 2 * KVH) * HD, HD):
            if np.abs(K[0, 0] - qkv_all[i, off:off + HD].astype(np.float32)).max() < 1e-3:
                print("  gpu K[kvh0,pos0] == input token", i, "offset", off)
    print(" V err", np.abs(vc[:, :prefill].float().cpu().numpy() - g["final_v"][:, :prefill]).max(axis=(0, 2)))
