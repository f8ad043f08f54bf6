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


class LlamaModel:
    """Synthetic-weight Llama + tce_llama handle.  Holds the torch tensors alive (the C side borrows pointers)."""

    def __init__(self, ctx: Context, geom: LlamaGeometry, max_ctx: int = 4096, seed: int = 1234, random_zeros: bool = False):
        self.ctx, self.geom, self.max_ctx = ctx, geom, max_ctx
        dev = torch.device("cuda", ctx.device)
        g = geom
        hd = g.head_dim
        self.tensors = []
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)

        def w4(oc, ic, s):
            t = random_w4(oc, ic, dev, s, scale=0.02, random_zeros=random_zeros)
            self.tensors.append(t)
            return _lib.W4Tensor(t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), oc, ic)

        def norm():
            t = (1.0 + 0.02 * torch.randn(g.embed_dim, device=dev, generator=gen)).float()
            self.tensors.append(t)
            return t.data_ptr()

        self.layers = (_lib.LlamaLayer * g.num_layers)()
        for l in range(g.num_layers):
            L = self.layers[l]
            s = seed * 1000 + l * 16
            L.q = w4(g.num_heads * hd, g.embed_dim, s + 1)
            L.k = w4(g.num_kv_heads * hd, g.embed_dim, s + 2)
            L.v = w4(g.num_kv_heads * hd, g.embed_dim, s + 3)
            L.o = w4(g.embed_dim, g.num_heads * hd, s + 4)
            L.gate = w4(g.hidden_dim, g.embed_dim, s + 5)
            L.up = w4(g.hidden_dim, g.embed_dim, s + 6)
            L.down = w4(g.embed_dim, g.hidden_dim, s + 7)
            L.input_norm = norm()
            L.post_norm = norm()
        self.embed = (torch.randn((g.vocab_size, g.embed_dim), device=dev, generator=gen) * 0.5).to(torch.float16)
        self.final_norm = (1.0 + 0.02 * torch.randn(g.embed_dim, device=dev, generator=gen)).float()
        self.weights = _lib.LlamaWeights()
        self.weights.embed_f16 = self.embed.data_ptr()
        self.weights.layers = C.cast(self.layers, C.POINTER(_lib.LlamaLayer))
        self.weights.final_norm = self.final_norm.data_ptr()
        self.weights.lm_head = w4(g.vocab_size, g.embed_dim, seed * 1000 + 999983)
        self.weights.rope_cos = None
        self.weights.rope_sin = None
        self.cfg = _lib.LlamaConfig(g.num_layers, g.num_heads, g.num_kv_heads, hd, g.embed_dim, g.hidden_dim, g.vocab_size, max_ctx,
                                    g.rms_eps, g.rope_theta, 0.0, 0, 1)
        h = C.c_void_p()
        _lib.check(ctx.L.tce_llama_create(ctx.h, C.byref(self.cfg), C.byref(self.weights), C.byref(h)), "tce_llama_create")
        self.h = h
        self.kernels_per_step = ctx.L.tce_llama_kernels_per_step(h)
        torch.cuda.synchronize(dev)

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
