"""numpy port of the reference's offline INT4 quantizer formats (llm/tools/quantize_methods.py).

TEST INFRASTRUCTURE ONLY.  Checked byte-for-byte against the reference Python module imported from
/root/reference in tests/golden/make_golden.py (fixtures committed under tests/golden/).

Common rule for every method (quantize_methods.py:393-413): per block of ``qk`` consecutive weights,
``d = (signed value of largest magnitude) / -8``; ``id = 1/d`` (0 if d == 0);
``q = clip(x*id + 8.5, 0, 15)`` truncated to int.  All fp32.
"""
from __future__ import annotations

import numpy as np


def make_divisible(c: int, divisor: int) -> int:
    return (c + divisor - 1) // divisor


def calculate_zeros_width(in_features: int, group_size: int = 128, pack_num: int = 8) -> int:
    """quantize_methods.py:9-21"""
    if group_size >= 128:
        size_multiplier = 1
    elif group_size == 64:
        size_multiplier = 2
    elif group_size == 32:
        size_multiplier = 4
    else:
        raise NotImplementedError
    base_width = make_divisible(in_features // group_size, pack_num)
    return make_divisible(base_width, size_multiplier) * size_multiplier


def _quantize_blocks(w: np.ndarray, qk: int):
    """w fp32 [OC, IC] row-major -> (q int32 [OC*IC/qk, qk], d fp32 [nb])."""
    x = np.ascontiguousarray(w, np.float32).reshape(-1, qk)
    idx = np.argmax(np.abs(x), axis=1)
    max_vals = x[np.arange(x.shape[0]), idx]
    d = (max_vals / -8).astype(np.float32)
    with np.errstate(divide="ignore"):
        idv = (1.0 / d).astype(np.float32)
    idv[d == 0] = 0.0
    q = ((x * idv[:, None]) + 8.5).clip(0, 15).astype(np.int32)
    return q, d


def quantize_q4_6(w: np.ndarray, qk: int = 128):
    """QM_CUDA GEMV format (quantize_row_q4_6, quantize_methods.py:370-442).

    Returns (qs uint32 [OC, IC/8] sequential nibbles, scales fp16 [OC, zw*8] zero padded,
    zeros uint32 [OC, zw] every nibble = 8)."""
    oc, ic = w.shape
    q, d = _quantize_blocks(w, qk)
    xi = q.reshape(oc, ic).astype(np.uint32)
    qs = np.zeros((oc, ic // 8), np.uint32)
    for i in range(8):
        qs |= (xi[:, i::8] & 0xF) << np.uint32(4 * i)
    zw = calculate_zeros_width(ic, qk)
    scales = np.zeros((oc, zw * 8), np.float32)
    scales[:, : ic // qk] = d.reshape(oc, ic // qk)
    zeros = np.full((oc, zw), 0x88888888, np.uint32)
    return qs, scales.astype(np.float16), zeros


def quantize_q4_3(w: np.ndarray):
    """QM_x86 format (quantize_row_q4_3, quantize_methods.py:188-242): block 32, byte e of a 64-weight pair of
    blocks = w[e] | w[32+e] << 4.  Returns (qs uint8 [OC, IC/2], scales fp32 [OC, IC/32])."""
    oc, ic = w.shape
    q, d = _quantize_blocks(w, 32)
    xi = q.astype(np.uint8).reshape(-1, 64)
    qs = (xi[:, :32] | (xi[:, 32:] << 4)).astype(np.uint8)
    return qs.reshape(oc, ic // 2), d.reshape(oc, ic // 32)


def quantize_q4_0_sequential(w: np.ndarray, qk: int = 128):
    """Sequential-nibble bytes [OC, IC/2] (lo nibble = even k) + fp32 scales [OC, IC/qk] with scalar zero 8:
    the layout the generic branch of naive_mat_mul_int4 reads (kernels/matmul_int4.cc:106-127)."""
    oc, ic = w.shape
    q, d = _quantize_blocks(w, qk)
    xi = q.reshape(oc, ic).astype(np.uint8)
    qs = (xi[:, 0::2] | (xi[:, 1::2] << 4)).astype(np.uint8)
    return qs, d.reshape(oc, ic // qk)


def quantize_q4_5(w: np.ndarray, qk: int = 128):
    """AWQ-GEMM format (quantize_row_q4_5, quantize_methods.py:299-368): qs int32 [IC, OC/8] nibble order
    0 2 4 6 1 3 5 7, scales fp16 [IC/qk, OC], zero fixed 8."""
    oc, ic = w.shape
    q, d = _quantize_blocks(w, qk)
    xi = q.reshape(oc, ic).T.astype(np.uint32)  # [IC, OC]
    qs = np.zeros((ic, oc // 8), np.uint32)
    order = [0, 2, 4, 6, 1, 3, 5, 7]
    for pos, src in enumerate(order):
        qs |= (xi[:, src::8] & 0xF) << np.uint32(4 * pos)
    scales = d.reshape(oc, ic // qk).T.astype(np.float16)
    return qs.view(np.int32), np.ascontiguousarray(scales)


def qmcuda_to_sequential_bytes(qs: np.ndarray) -> np.ndarray:
    """uint32 [OC, IC/8] (nibble i = ic 8w+i) -> uint8 [OC, IC/2] (lo nibble = even ic): same bytes, little endian."""
    return np.ascontiguousarray(qs, np.uint32).view(np.uint8).reshape(qs.shape[0], -1)


def dequant_qmcuda(qs: np.ndarray, scales: np.ndarray, zeros: np.ndarray, group: int = 128) -> np.ndarray:
    """fp32 [OC, IC] = s * (q - z), the arithmetic of gemv_kernel_g128 (kernels/cuda/gemv_cuda.cu:181-183)."""
    oc, wpr = qs.shape
    ic = wpr * 8
    q = np.zeros((oc, ic), np.float32)
    for i in range(8):
        q[:, i::8] = ((qs >> np.uint32(4 * i)) & 0xF).astype(np.float32)
    ng = ic // group
    z = np.zeros((oc, ng), np.float32)
    for g in range(ng):
        z[:, g] = ((zeros[:, g // 8] >> np.uint32(4 * (g % 8))) & 0xF).astype(np.float32)
    s = scales[:, :ng].astype(np.float32)
    return (np.repeat(s, group, axis=1) * (q - np.repeat(z, group, axis=1))).astype(np.float32)
