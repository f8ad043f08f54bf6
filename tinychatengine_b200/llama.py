"""Llama decode runner over the C ABI (tce_llama_*): synthetic-weight builder + step API.

Geometry table = reference llm/include/model.h:71-83.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib
from .runtime import Context, random_w4


@dataclass
class LlamaGeometry:
    name: str
    num_layers: int
    num_heads: int
    num_kv_heads: int
    embed_dim: int
    hidden_dim: int
    vocab_size: int
    rms_eps: float = 1e-5
    rope_theta: float = 10000.0
    head_dim: int = 128


GEOMETRIES = {
    "llama3-8b": LlamaGeometry("llama3-8b", 32, 32, 8, 4096, 14336, 128256, 1e-5, 500000.0),
    "llama2-7b": LlamaGeometry("llama2-7b", 32, 32, 32, 4096, 11008, 32000, 1e-6, 10000.0),
    "llama2-13b": LlamaGeometry("llama2-13b", 40, 40, 40, 5120, 13824, 32000, 1e-6, 10000.0),
    # tiny shapes for tests (same structure, GQA 4:1 like Llama-3)
    "tiny-gqa": LlamaGeometry("tiny-gqa", 2, 8, 2, 1024, 2816, 2048, 1e-5, 500000.0),
    "tiny-mha": LlamaGeometry("tiny-mha", 2, 4, 4, 512, 1408, 1024, 1e-6, 10000.0),
}


def weight_bytes_per_token(g: LlamaGeometry) -> int:
    """Algorithmic HBM bytes one decode step must read from the packed weights (nibbles + fp16 scales + 4-bit
    zeros, unpadded), SURVEY.md 8(d): 3.899 GB for Llama-3-8B."""
    hd = g.head_dim
    mats = []
    for _ in range(g.num_layers):
        mats += [(g.num_heads * hd, g.embed_dim), (g.num_kv_heads * hd, g.embed_dim), (g.num_kv_heads * hd, g.embed_dim),
                 (g.embed_dim, g.num_heads * hd), (g.hidden_dim, g.embed_dim), (g.hidden_dim, g.embed_dim), (g.embed_dim, g.hidden_dim)]
    mats.append((g.vocab_size, g.embed_dim))
    total = 0
    for oc, ic in mats:
        total += oc * ic // 2 + oc * (ic // 128) * 2 + oc * (ic // 128) // 2
    return total


def kv_bytes_per_token(g: LlamaGeometry, ctx_len: int) -> int:
    """fp16 K+V read for `ctx_len` cached positions plus the one-row append, all layers."""
    per_pos = 2 * g.num_kv_heads * g.head_dim * 2 * g.num_layers
    return per_pos * ctx_len + per_pos


def make_random_weights(geom: LlamaGeometry, dev, seed: int = 1234, random_zeros: bool = False, embed_rows: int | None = None):
    """Synthetic AWQ-INT4 Llama weights (QM_CUDA layout) as a plain dict of torch tensors on `dev`.  `embed_rows`: rows of the
    embedding table when it differs from geom.vocab_size (a tensor-parallel rank holds a vocabulary SHARD of lm_head but looks up
    GLOBAL token ids, so its table has all the rows)."""
    g = geom
    hd = g.head_dim
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    W = {"layers": []}
    for l in range(g.num_layers):
        s = seed * 1000 + l * 16
        L = {"q": random_w4(g.num_heads * hd, g.embed_dim, dev, s + 1, 0.02, random_zeros),
             "k": random_w4(g.num_kv_heads * hd, g.embed_dim, dev, s + 2, 0.02, random_zeros),
             "v": random_w4(g.num_kv_heads * hd, g.embed_dim, dev, s + 3, 0.02, random_zeros),
             "o": random_w4(g.embed_dim, g.num_heads * hd, dev, s + 4, 0.02, random_zeros),
             "gate": random_w4(g.hidden_dim, g.embed_dim, dev, s + 5, 0.02, random_zeros),
             "up": random_w4(g.hidden_dim, g.embed_dim, dev, s + 6, 0.02, random_zeros),
             "down": random_w4(g.embed_dim, g.hidden_dim, dev, s + 7, 0.02, random_zeros),
             "input_norm": (1.0 + 0.02 * torch.randn(g.embed_dim, device=dev, generator=gen)).float(),
             "post_norm": (1.0 + 0.02 * torch.randn(g.embed_dim, device=dev, generator=gen)).float()}
        W["layers"].append(L)
    W["embed"] = (torch.randn((embed_rows or g.vocab_size, g.embed_dim), device=dev, generator=gen) * 0.5).to(torch.float16)
    W["final_norm"] = (1.0 + 0.02 * torch.randn(g.embed_dim, device=dev, generator=gen)).float()
    W["lm_head"] = random_w4(g.vocab_size, g.embed_dim, dev, seed * 1000 + 999983, 0.02, random_zeros)
    return W


def shard_w4_rows(t, rank: int, P: int):
    """Output-channel (row) shard of a QM_CUDA tensor triple: contiguous rows [rank*OC/P, (rank+1)*OC/P)."""
    w, z, s = t
    n = w.shape[0] // P
    return tuple(x[rank * n:(rank + 1) * n].contiguous() for x in (w, z, s))


def shard_w4_cols(t, ic: int, rank: int, P: int, group: int = 128):
    """Input-channel shard (row-parallel linear): columns [rank*IC/P, (rank+1)*IC/P) repacked as a QM_CUDA tensor of
    its own: the packed words are a contiguous slice per row, scales/zeros are re-padded to the shard's zeros_width."""
    from .formats import zeros_width

    w, z, s = t
    icl = ic // P
    assert icl % group == 0
    ngl = icl // group
    zwl = zeros_width(icl, group)
    wl = w[:, rank * icl // 8:(rank + 1) * icl // 8].contiguous()
    sl = torch.zeros((w.shape[0], zwl * 8), dtype=s.dtype, device=s.device)
    sl[:, :ngl] = s[:, rank * ngl:(rank + 1) * ngl]
    zi = z.to(torch.int64) & 0xFFFFFFFF
    nib = torch.stack([(zi >> (4 * i)) & 0xF for i in range(8)], dim=2).reshape(z.shape[0], -1)  # [OC, zw*8]
    nl = torch.full((z.shape[0], zwl * 8), 8, dtype=torch.int64, device=z.device)
    nl[:, :ngl] = nib[:, rank * ngl:(rank + 1) * ngl]
    packed = torch.zeros((z.shape[0], zwl), dtype=torch.int64, device=z.device)
    for i in range(8):
        packed |= nl[:, i::8] << (4 * i)
    packed = torch.where(packed >= 2 ** 31, packed - 2 ** 32, packed).to(torch.int32)
    return wl, packed.contiguous(), sl.contiguous()


def shard_weights(W, geom: LlamaGeometry, rank: int, P: int):
    """Megatron-style shard: q/k/v/gate/up/lm_head by rows, o/down by input channels."""
    hd = geom.head_dim
    gl = LlamaGeometry(geom.name, geom.num_layers, geom.num_heads // P, geom.num_kv_heads // P, geom.embed_dim, geom.hidden_dim // P,
                       geom.vocab_size // P, geom.rms_eps, geom.rope_theta, hd)
    Wl = {"layers": [], "embed": W["embed"], "final_norm": W["final_norm"], "lm_head": shard_w4_rows(W["lm_head"], rank, P)}
    for L in W["layers"]:
        Wl["layers"].append({"q": shard_w4_rows(L["q"], rank, P), "k": shard_w4_rows(L["k"], rank, P), "v": shard_w4_rows(L["v"], rank, P),
                             "o": shard_w4_cols(L["o"], geom.num_heads * hd, rank, P), "gate": shard_w4_rows(L["gate"], rank, P),
                             "up": shard_w4_rows(L["up"], rank, P), "down": shard_w4_cols(L["down"], geom.hidden_dim, rank, P),
                             "input_norm": L["input_norm"], "post_norm": L["post_norm"]})
    return Wl, gl


class LlamaModel:
    """Llama weights (synthetic by default) + tce_llama handle.  Holds the torch tensors alive (the C side borrows
    pointers).  `geom` describes the LOCAL shard when tp_size > 1 (see shard_weights)."""

    def __init__(self, ctx: Context, geom: LlamaGeometry, max_ctx: int = 4096, seed: int = 1234, random_zeros: bool = False, weights=None,
                 tp_rank: int = 0, tp_size: int = 1):
        self.ctx, self.geom, self.max_ctx = ctx, geom, max_ctx
        self.tp_rank, self.tp_size = tp_rank, tp_size
        dev = torch.device("cuda", ctx.device)
        g = geom
        hd = g.head_dim
        W = weights if weights is not None else make_random_weights(geom, dev, seed, random_zeros, embed_rows=geom.vocab_size * max(1, tp_size))
        self.W = W
        self.tensors = []

        def w4(t, oc, ic):
            assert t[0].shape == (oc, ic // 8), (t[0].shape, oc, ic)
            self.tensors.append(t)
            return _lib.W4Tensor(t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), oc, ic)

        def norm(t):
            self.tensors.append(t)
            return t.data_ptr()

        self.layers = (_lib.LlamaLayer * g.num_layers)()
        for l in range(g.num_layers):
            L, T = self.layers[l], W["layers"][l]
            L.q = w4(T["q"], g.num_heads * hd, g.embed_dim)
            L.k = w4(T["k"], g.num_kv_heads * hd, g.embed_dim)
            L.v = w4(T["v"], g.num_kv_heads * hd, g.embed_dim)
            L.o = w4(T["o"], g.embed_dim, g.num_heads * hd)
            L.gate = w4(T["gate"], g.hidden_dim, g.embed_dim)
            L.up = w4(T["up"], g.hidden_dim, g.embed_dim)
            L.down = w4(T["down"], g.embed_dim, g.hidden_dim)
            L.input_norm = norm(T["input_norm"])
            L.post_norm = norm(T["post_norm"])
        self.embed = W["embed"]
        self.final_norm = W["final_norm"]
        self.weights = _lib.LlamaWeights()
        self.weights.embed_f16 = self.embed.data_ptr()
        self.weights.layers = C.cast(self.layers, C.POINTER(_lib.LlamaLayer))
        self.weights.final_norm = self.final_norm.data_ptr()
        self.weights.lm_head = w4(W["lm_head"], g.vocab_size, g.embed_dim)
        self.weights.rope_cos = None
        self.weights.rope_sin = None
        self.cfg = _lib.LlamaConfig(g.num_layers, g.num_heads, g.num_kv_heads, hd, g.embed_dim, g.hidden_dim, g.vocab_size, max_ctx,
                                    g.rms_eps, g.rope_theta, 0.0, tp_rank, tp_size)
        h = C.c_void_p()
        _lib.check(ctx.L.tce_llama_create(ctx.h, C.byref(self.cfg), C.byref(self.weights), C.byref(h)), "tce_llama_create")
        self.h = h
        self.kernels_per_step = ctx.L.tce_llama_kernels_per_step(h)
        torch.cuda.synchronize(dev)

    @classmethod
    def load_dir(cls, ctx: Context, path, geom: LlamaGeometry, max_ctx: int = 4096):
        """Model built by the C++ loader from a parameter tree in the reference's on-disk layout (tce_llama_load_dir); the library owns the
        device copies.  `geom` plays the role of the reference's model_config."""
        self = cls.__new__(cls)
        self.ctx, self.geom, self.max_ctx = ctx, geom, max_ctx
        self.tp_rank, self.tp_size = 0, 1
        self.W, self.tensors = None, []
        g = geom
        self.cfg = _lib.LlamaConfig(g.num_layers, g.num_heads, g.num_kv_heads, g.head_dim, g.embed_dim, g.hidden_dim, g.vocab_size, max_ctx,
                                    g.rms_eps, g.rope_theta, 0.0, 0, 1)
        h = C.c_void_p()
        _lib.check(ctx.L.tce_llama_load_dir(ctx.h, str(path).encode(), C.byref(self.cfg), C.byref(h)), "tce_llama_load_dir")
        self.h = h
        self.kernels_per_step = ctx.L.tce_llama_kernels_per_step(h)
        return self

    def save_dir(self, path, rotary: bool = True):
        """Write this model's (synthetic) weights as a parameter tree in the reference's QM_CUDA on-disk layout (tests, examples):
        the inverse of load_dir.  q|k|v are written merged under self_attn/qkv_proj, as llm/tools/llama_qkv_merger.py leaves them."""
        from pathlib import Path

        import numpy as np

        from . import formats

        root = Path(path)
        W = self.W

        def f32(t, p):
            p.parent.mkdir(parents=True, exist_ok=True)
            t.detach().float().cpu().numpy().astype(np.float32).tofile(p)

        def w4(t, p):
            formats.save_qm_cuda_dir(p, t[0].cpu().numpy().view(np.uint32), t[1].cpu().numpy().view(np.uint32), t[2].cpu().numpy())

        f32(W["embed"], root / "decoder" / "embed_tokens" / "weight.bin")
        f32(W["final_norm"], root / "decoder" / "norm" / "weight.bin")
        for l, T in enumerate(W["layers"]):
            lp = root / "decoder" / f"layer{l}"
            f32(T["input_norm"], lp / "input_layernorm" / "weight.bin")
            f32(T["post_norm"], lp / "post_attention_layernorm" / "weight.bin")
            merged = tuple(torch.cat([T[n][i] for n in ("q", "k", "v")], dim=0) for i in range(3))
            w4(merged, lp / "self_attn" / "qkv_proj")
            w4(T["o"], lp / "self_attn" / "o_proj")
            for n in ("gate", "up", "down"):
                w4(T[n], lp / f"{n}_proj")
        w4(W["lm_head"], root / "lm_head")
        if rotary:
            # rotary_emb tables and the qk scale as the reference trees carry them (fp16; llm/tools/rotary_emb_exporter.py, read by
            # RotaryPosEmb_cuda's constructor and Int4llamaAttention.cu:82): a tree with tables overrides rope_theta on load
            g = self.geom
            hd = g.head_dim
            inv = 1.0 / (g.rope_theta ** (np.arange(0, hd, 2, dtype=np.float64) / hd))
            ang = np.arange(self.max_ctx)[:, None] * inv[None, :]
            emb = np.concatenate([ang, ang], axis=1)
            for l in range(g.num_layers):
                sa = root / "decoder" / f"layer{l}" / "self_attn"
                (sa / "rotary_emb").mkdir(parents=True, exist_ok=True)
                (sa / "qk_bmm").mkdir(parents=True, exist_ok=True)
                np.cos(emb).astype(np.float16).tofile(sa / "rotary_emb" / "cos_cached_half.bin")
                np.sin(emb).astype(np.float16).tofile(sa / "rotary_emb" / "sin_cached_half.bin")
                np.array([1.0 / np.sqrt(hd)], dtype=np.float16).tofile(sa / "qk_bmm" / "alpha_half.bin")

    def tp_connect(self, group=None):
        """Exchange the IPC handles of the peer-visible buffers over torch.distributed and map the peers."""
        import torch.distributed as dist

        mine = (C.c_ubyte * 64)()
        _lib.check(self.ctx.L.tce_llama_tp_handle(self.h, C.cast(mine, C.c_void_p)), "tce_llama_tp_handle")
        t = torch.tensor(list(mine), dtype=torch.uint8)
        backend = dist.get_backend(group)
        if backend == "nccl":
            t = t.cuda(self.ctx.device)
        out = [torch.empty_like(t) for _ in range(self.tp_size)]
        dist.all_gather(out, t, group=group)
        blob = b"".join(bytes(o.cpu().tolist()) for o in out)
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        _lib.check(self.ctx.L.tce_llama_tp_connect(self.h, C.cast(buf, C.c_void_p)), "tce_llama_tp_connect")
        self.kernels_per_step = self.ctx.L.tce_llama_kernels_per_step(self.h)  # the step's form is only known once the peers are mapped
        dist.barrier(group=group)

    def layer_tensors(self, l: int):
        """(q, k, v, o, gate, up, down) each as (w, zeros, scales) torch tensors, plus the two norm gammas."""
        per = 9
        base = l * per
        t = self.tensors[base: base + per]
        return {"q": t[0], "k": t[1], "v": t[2], "o": t[3], "gate": t[4], "up": t[5], "down": t[6], "input_norm": t[7], "post_norm": t[8]}

    def decode(self, tokpos_dev: torch.Tensor):
        _lib.check(self.ctx.L.tce_llama_decode(self.h, C.c_void_p(tokpos_dev.data_ptr())), "tce_llama_decode")

    def decode_host(self, token: int, pos: int, logits_host=None):
        nxt = C.c_int(-1)
        p = None if logits_host is None else C.c_void_p(logits_host.data_ptr())
        _lib.check(self.ctx.L.tce_llama_decode_host(self.h, int(token), int(pos), p, C.byref(nxt)), "tce_llama_decode_host")
        return nxt.value

    def generate(self, first_token: int, pos0: int, n_predict: int, *, history=(), eos_id: int = -1, top_k=40, top_p=0.95, temp=0.8, repeat_penalty=1.1,
                 frequency_penalty=0.0, presence_penalty=0.0, repeat_last_n=64, seed=0):
        """Device generate loop (decode + sample per token, only the ids come back); defaults are the reference's opt_params."""
        import numpy as np

        cfg = _lib.Sampling(int(top_k), float(top_p), float(temp), float(repeat_penalty), float(frequency_penalty), float(presence_penalty),
                            int(repeat_last_n), int(seed))
        hist = np.ascontiguousarray(np.asarray(list(history), dtype=np.int32))
        out = np.zeros(max(1, n_predict), dtype=np.int32)
        n = C.c_int(0)
        _lib.check(self.ctx.L.tce_llama_generate(self.h, int(first_token), int(pos0), int(n_predict), C.byref(cfg),
                                                 hist.ctypes.data_as(C.c_void_p) if hist.size else None, int(hist.size), int(eos_id),
                                                 out.ctypes.data_as(C.c_void_p), C.byref(n)), "tce_llama_generate")
        return out[:n.value].tolist()

    def prefill(self, tokens, pos0: int = 0, logits_host=None) -> int:
        """Prompt processing: all `tokens` (host ints) at positions pos0.. in one pass; returns the greedy next token."""
        arr = (C.c_int * len(tokens))(*[int(t) for t in tokens])
        nxt = C.c_int(-1)
        p = None if logits_host is None else C.c_void_p(logits_host.data_ptr())
        _lib.check(self.ctx.L.tce_llama_prefill(self.h, arr, len(tokens), int(pos0), p, C.byref(nxt)), "tce_llama_prefill")
        return nxt.value

    def logits(self) -> torch.Tensor:
        """View of the device logits buffer (float32 [vocab])."""
        ptr = self.ctx.L.tce_llama_logits(self.h)
        return _tensor_from_ptr(ptr, (self.geom.vocab_size,), torch.float32, self.ctx.device)

    def kv_cache(self, layer: int, which: int) -> torch.Tensor:
        ptr = self.ctx.L.tce_llama_kv_cache(self.h, layer, which)
        return _tensor_from_ptr(ptr, (self.geom.num_kv_heads, self.max_ctx, self.geom.head_dim), torch.float16, self.ctx.device)

    def debug_buffer(self, which: int) -> torch.Tensor:
        g = self.geom
        shape, dt = {0: ((g.embed_dim,), torch.float32), 1: (((g.num_heads + 2 * g.num_kv_heads) * g.head_dim,), torch.float16),
                     2: ((g.num_heads * g.head_dim,), torch.float16), 3: ((g.hidden_dim,), torch.float16)}[which]
        return _tensor_from_ptr(self.ctx.L.tce_llama_debug_buffer(self.h, which), shape, dt, self.ctx.device)

    def close(self):
        if self.h:
            self.ctx.L.tce_llama_destroy(self.h)
            self.h = None


class _CudaArrayHolder:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3}


def _tensor_from_ptr(ptr: int, shape, dtype, device: int) -> torch.Tensor:
    typestr = {torch.float32: "<f4", torch.float16: "<f2", torch.int32: "<i4", torch.int8: "|i1"}[dtype]
    return torch.as_tensor(_CudaArrayHolder(ptr, tuple(shape), typestr), device=torch.device("cuda", device))
