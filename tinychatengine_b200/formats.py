"""The reference's INT4 data formats on the host side (numpy): QM_CUDA packing, the offline quantizer rule and
the on-disk ``*_int4.bin`` tree.

Format source of truth: reference llm/tools/quantize_methods.py:370-442 (quantize_row_q4_6) and
llm/tools/model_quantizer.py:35-66 (file names / dtypes); padding rule calculate_zeros_width
(llm/src/nn_modules/cuda/utils.cu:162-178).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

GROUP = 128  # QK under QM_CUDA (llm/include/common.h:17-21)


def zeros_width(in_features: int, group: int = GROUP, pack_num: int = 8) -> int:
    mult = 1 if group >= 128 else (2 if group == 64 else 4)
    base = (in_features // group + pack_num - 1) // pack_num
    return (base + mult - 1) // mult * mult


def pack_qm_cuda(q: np.ndarray, scales: np.ndarray, zeros: np.ndarray, group: int = GROUP, pad_zero: int = 8):
    """q uint8 [OC, IC] in 0..15, scales [OC, IC/group], zeros uint8 [OC, IC/group] ->
    (w uint32 [OC, IC/8], zeros_packed uint32 [OC, zw], scales_padded fp16 [OC, zw*8])."""
    oc, ic = q.shape
    ng = ic // group
    zw = zeros_width(ic, group)
    q32 = q.astype(np.uint32)
    w = np.zeros((oc, ic // 8), np.uint32)
    for i in range(8):
        w |= (q32[:, i::8] & 0xF) << np.uint32(4 * i)
    zp = np.full((oc, zw * 8), pad_zero & 0xF, np.uint32)  # the reference fills the padding nibbles with 8 too
    zp[:, :ng] = zeros.astype(np.uint32) & 0xF
    zpk = np.zeros((oc, zw), np.uint32)
    for i in range(8):
        zpk |= zp[:, i::8] << np.uint32(4 * i)
    sp = np.zeros((oc, zw * 8), np.float16)
    sp[:, :ng] = scales.astype(np.float16)
    return w, zpk, sp


def unpack_qm_cuda(w: np.ndarray, zeros_packed: np.ndarray, scales_padded: np.ndarray, group: int = GROUP):
    oc, wpr = w.shape
    ic = wpr * 8
    ng = ic // group
    q = np.zeros((oc, ic), np.uint8)
    for i in range(8):
        q[:, i::8] = ((w >> np.uint32(4 * i)) & 0xF).astype(np.uint8)
    z = np.zeros((oc, ng), np.uint8)
    for g in range(ng):
        z[:, g] = ((zeros_packed[:, g // 8] >> np.uint32(4 * (g % 8))) & 0xF).astype(np.uint8)
    return q, scales_padded[:, :ng].copy(), z


def quantize_qm_cuda(weight: np.ndarray, group: int = GROUP):
    """fp32 [OC, IC] -> QM_CUDA tensors with the reference's rule: d = (value of largest magnitude) / -8,
    q = trunc(clip(x/d + 8.5, 0, 15)), zero point 8 (quantize_methods.py:393-442)."""
    oc, ic = weight.shape
    x = np.ascontiguousarray(weight, np.float32).reshape(-1, group)
    idx = np.argmax(np.abs(x), axis=1)
    d = (x[np.arange(x.shape[0]), idx] / -8).astype(np.float32)
    with np.errstate(divide="ignore"):
        inv = (1.0 / d).astype(np.float32)
    inv[d == 0] = 0.0
    q = ((x * inv[:, None]) + 8.5).clip(0, 15).astype(np.uint8).reshape(oc, ic)
    return pack_qm_cuda(q, d.reshape(oc, ic // group), np.full((oc, ic // group), 8, np.uint8), group)


def load_qm_cuda_dir(path: str | Path, oc: int, ic: int, group: int = GROUP):
    """Read one op directory of the reference model tree (weight_int4.bin / scaling_factor_int4.bin /
    zero_point_int4.bin), as Linear_half_int4's constructor does (llm/include/ops/linear.h:188-210)."""
    path = Path(path)
    zw = zeros_width(ic, group)
    w = np.fromfile(path / "weight_int4.bin", dtype=np.uint32)
    s = np.fromfile(path / "scaling_factor_int4.bin", dtype=np.float16)
    z = np.fromfile(path / "zero_point_int4.bin", dtype=np.uint32)
    if w.size != oc * ic // 8 or s.size != oc * zw * 8 or z.size != oc * zw:
        raise ValueError(f"{path}: file sizes do not match OC={oc} IC={ic} (zeros_w={zw})")
    return w.reshape(oc, ic // 8), z.reshape(oc, zw), s.reshape(oc, zw * 8)


def save_qm_cuda_dir(path: str | Path, w: np.ndarray, zeros: np.ndarray, scales: np.ndarray) -> None:
    path = Path(path)
    path.mkdir(parents=True, exist_ok=True)
    np.ascontiguousarray(w, np.uint32).tofile(path / "weight_int4.bin")
    np.ascontiguousarray(scales, np.float16).tofile(path / "scaling_factor_int4.bin")
    np.ascontiguousarray(zeros, np.uint32).tofile(path / "zero_point_int4.bin")


# ---- QM_x86 (CPU build) flavour: llm/tools/quantize_methods.py:188-243 (quantize_row_q4_3), group 32 ---------------------------------
def quantize_qm_x86(weight: np.ndarray):
    """fp32 [OC, IC] -> (qs uint8 [OC, IC/2], scales fp32 [OC, IC/32]): d = (element of largest magnitude) / -8 per 32-group,
    q = trunc(clip(x/d + 8.5, 0, 15)); byte e of every 64-weight run = q[e] | q[32 + e] << 4; zero point 8."""
    oc, ic = weight.shape
    x = np.ascontiguousarray(weight, np.float32).reshape(-1, 32)
    idx = np.argmax(np.abs(x), axis=1)
    d = (x[np.arange(x.shape[0]), idx] / np.float32(-8)).astype(np.float32)
    with np.errstate(divide="ignore"):
        inv = np.where(d != 0, np.float32(1.0) / d, np.float32(0)).astype(np.float32)
    q = ((x * inv[:, None]) + np.float32(8.5)).clip(0, 15).astype(np.uint8).reshape(-1, 64)
    qs = (q[:, :32] | (q[:, 32:] << 4)).astype(np.uint8)
    return qs.reshape(oc, ic // 2), d.reshape(oc, ic // 32)


def dequantize_qm_x86(qs: np.ndarray, scales: np.ndarray) -> np.ndarray:
    oc = qs.shape[0]
    b = qs.reshape(-1, 32)
    q = np.concatenate([(b & 0xF), (b >> 4)], axis=1).astype(np.float32) - np.float32(8)  # [runs, 64]
    d = scales.reshape(-1, 2)
    w = np.concatenate([q[:, :32] * d[:, :1], q[:, 32:] * d[:, 1:]], axis=1)
    return w.reshape(oc, -1).astype(np.float32)
