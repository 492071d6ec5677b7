"""CPU restatement (numpy) of the reference's quantized-decoder hot path.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and
bench.py's `cpu_baseline` / `--impl reference` legs may import this module; the product
(`ctranslate2_b200`) never does and fails loudly when its CUDA library is missing.

Parity status: PINNED.  Every function below is checked (tests/test_oracle.py, `-m "not gpu"`)
against (a) the golden vectors the reference's own gtests hold for this path (tests/golden/
ref_gtest_vectors.json, extracted from /root/reference/tests/ops_test.cc and layers_test.cc) and
(b) outputs of the UNMODIFIED reference compiled by oracle/Makefile.ref (oracle/_ref), whose
fixtures are committed under tests/golden/ by tools/make_golden.py.

Each function cites the reference file:line it follows (paths relative to /root/reference).
All arithmetic is float32 unless stated; integer paths are exact.
"""
from __future__ import annotations

import json
import math
import os
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

f32 = np.float32

# ops::ActivationType order, include/ctranslate2/ops/activation.h:9-17
ACT_NONE = -1
ACT_RELU, ACT_GELU_TANH, ACT_SWISH, ACT_GELU, ACT_GELU_SIGMOID, ACT_TANH, ACT_SIGMOID = range(7)

_erf = np.vectorize(math.erf, otypes=[np.float64])


def activation(x: np.ndarray, act: int) -> np.ndarray:
    """Epilogue functors: src/cpu/kernels.cc:140-206 (CPU) == src/cuda/helpers.h:244-305 (CUDA)."""
    x = x.astype(f32)
    if act == ACT_NONE:
        return x
    if act == ACT_RELU:
        return np.maximum(x, f32(0))
    if act == ACT_SWISH:
        return (x / (f32(1) + np.exp(-x, dtype=f32))).astype(f32)
    if act == ACT_GELU:
        return (f32(0.5) * x * (f32(1) + _erf(x.astype(np.float64) * 0.7071067811865475).astype(f32))).astype(f32)
    if act == ACT_GELU_TANH:
        u = f32(0.7978845608028654) * (x + f32(0.044715) * x * x * x)
        return (f32(0.5) * x * (f32(1) + np.tanh(u, dtype=f32))).astype(f32)
    if act == ACT_GELU_SIGMOID:
        return (x / (f32(1) + np.exp(f32(-1.702) * x, dtype=f32))).astype(f32)
    if act == ACT_TANH:
        return np.tanh(x, dtype=f32)
    if act == ACT_SIGMOID:
        return (f32(1) / (f32(1) + np.exp(-x, dtype=f32))).astype(f32)
    raise ValueError(f"unknown activation {act}")


# --------------------------------------------------------------------------------------
# Quantize / INT8 GEMM / Dequantize  
# --------------------------------------------------------------------------------------

def quantize_rows(x: np.ndarray, round_before_cast: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """ops::Quantize, int8 arm.  src/ops/quantize.cc:21-50, src/cpu/kernels.cc:577-651 (CPU),
    src/ops/quantize_gpu.cu:57-105 (CUDA).  Per row: amax; scale = amax != 0 ? 127/amax : 1;
    q = int8(nearbyint(x * scale)) (round-half-even; no rounding => C truncation for
    binary_version < 5).  The amax is reduced in the INPUT dtype (exact for any dtype: max of abs)."""
    x2 = np.asarray(x)
    depth = x2.shape[-1]
    rows = x2.reshape(-1, depth).astype(f32)
    amax = np.max(np.abs(rows), axis=1).astype(f32)
    scale = np.where(amax != 0, f32(127) / np.where(amax != 0, amax, f32(1)), f32(1)).astype(f32)
    v = (rows * scale[:, None]).astype(f32)
    q = np.rint(v) if round_before_cast else np.trunc(v)
    return q.astype(np.int8).reshape(x2.shape), scale.reshape(x2.shape[:-1])


def gemm_s8(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """ops::Gemm int8 arm with alpha=1, beta=0, trans_b=true (the only form layers::Dense uses).
    src/ops/gemm.cc:45-107 -> primitives<>::gemm<int8_t,int32_t> (src/cuda/primitives.cu:571-597,
    CPU: src/cpu/primitives.cc Ruy/MKL/oneDNN).  a [M,K] int8, b [N,K] int8 -> c [M,N] int32, exact."""
    return a.astype(np.int32) @ b.astype(np.int32).T


def dequantize_gemm_output(c: np.ndarray, a_scale: np.ndarray, b_scale: np.ndarray,
                           bias: Optional[np.ndarray] = None, act: int = ACT_NONE,
                           flavor: str = "cuda") -> np.ndarray:
    """ops::Dequantize, GEMM-output form.  src/ops/dequantize.cc:46-59.
    flavor="cpu":  y = act(c * (1/a_scale) / b_scale + bias)      src/cpu/kernels.cc:653-686
    flavor="cuda": y = act(c / (a_scale * b_scale) + bias)        src/ops/dequantize_gpu.cu:30-54
    (the two differ in the last ulp; both are within the reference's own fp32 tolerance 1e-5)."""
    c32 = c.astype(f32)
    sa = np.asarray(a_scale, f32).reshape(-1, 1)
    sb = np.asarray(b_scale, f32).reshape(1, -1)
    if flavor == "cpu":
        v = (c32 * (f32(1) / sa)).astype(f32) / sb
    else:
        v = c32 / (sa * sb).astype(f32)
    v = v.astype(f32)
    if bias is not None:
        v = (v + np.asarray(bias, f32).reshape(1, -1)).astype(f32)
    return activation(v, act)


def dense_int8(x: np.ndarray, w_q: np.ndarray, w_scale: np.ndarray, bias: Optional[np.ndarray] = None,
               act: int = ACT_NONE, residual: Optional[np.ndarray] = None, flavor: str = "cuda",
               round_before_cast: bool = True) -> np.ndarray:
    """layers::Dense::operator(), quantized arm.  src/layers/common.cc:353-401:
    Quantize(x) -> Gemm s8 -> Dequantize(+bias, activation) -> Add(residual).  round_before_cast is
    Model::round_before_cast_in_quantization(): binary_version >= 5 (include/ctranslate2/models/model.h:87-89)."""
    shape = x.shape
    x2 = x.reshape(-1, shape[-1])
    xq, xs = quantize_rows(x2, round_before_cast)
    y = dequantize_gemm_output(gemm_s8(xq, w_q), xs, w_scale, bias, act, flavor)
    if residual is not None:
        y = (y + residual.reshape(y.shape).astype(f32)).astype(f32)
    return y.reshape(shape[:-1] + (w_q.shape[0],))


def bias_add(value: np.ndarray, bias: np.ndarray, act: int = ACT_NONE, residual: Optional[np.ndarray] = None,
             axis: int = -1) -> np.ndarray:
    """ops::BiasAdd: act(value + bias [+ residual]), bias broadcast along `axis`.  src/ops/bias_add.cc,
    src/cpu/kernels.cc add_bias_and_activation; goldens tests/ops_test.cc:1398-1432."""
    shape = [1] * value.ndim
    shape[axis] = bias.shape[0]
    y = (value.astype(f32) + bias.astype(f32).reshape(shape)).astype(f32)
    if residual is not None:
        y = (y + residual.astype(f32)).astype(f32)
    return activation(y, act)


def gemm_float(a: np.ndarray, b: np.ndarray, c: Optional[np.ndarray] = None, alpha: float = 1.0, beta: float = 0.0,
               trans_a: bool = False, trans_b: bool = False, bias: Optional[np.ndarray] = None,
               residual: Optional[np.ndarray] = None, act: int = ACT_NONE) -> np.ndarray:
    """ops::Gemm, float arm: C = act(alpha * op(A) op(B) + beta * C + bias + residual).  src/ops/gemm.cc:10-25 (the
    activation comes AFTER bias and residual) and :45-107; goldens tests/ops_test.cc:516-681.  layers::Dense uses
    alpha=1, beta=0, trans_b=true (common.cc:440)."""
    A = a.astype(f32).T if trans_a else a.astype(f32)
    B = b.astype(f32).T if trans_b else b.astype(f32)
    y = (f32(alpha) * (A @ B)).astype(f32)
    if c is not None and beta != 0.0:
        y = (y + f32(beta) * c.astype(f32)).astype(f32)
    if bias is not None:
        return bias_add(y, bias, act, residual)
    if residual is not None:
        y = (y + residual.astype(f32)).astype(f32)
    return activation(y, act)


def tp_shard_rows(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of a dimension split evenly over `world` ranks (models::Model::load splits weights with
    ops::Split into equal parts, src/models/model.cc:662-743; the dimension must be divisible)."""
    assert n_total % world == 0
    per = n_total // world
    return rank * per, (rank + 1) * per


def tp_qkv_rows(num_heads: int, num_heads_kv: int, head_dim: int, rank: int, world: int) -> np.ndarray:
    """Rows of the fused [q; k; v] projection owned by `rank`: its query heads, its key heads, its value heads
    (model.cc:689-722 splits the three parts separately so that every rank keeps whole heads)."""
    q0, q1 = tp_shard_rows(num_heads, rank, world)
    k0, k1 = tp_shard_rows(num_heads_kv, rank, world)
    hq, hk = num_heads * head_dim, num_heads_kv * head_dim
    return np.concatenate([np.arange(q0 * head_dim, q1 * head_dim),
                           hq + np.arange(k0 * head_dim, k1 * head_dim),
                           hq + hk + np.arange(k0 * head_dim, k1 * head_dim)])


def dense_int8_row_parallel(x: np.ndarray, w_q: np.ndarray, w_scale: np.ndarray, world: int,
                            residual: Optional[np.ndarray] = None, flavor: str = "cuda",
                            dtype: str = "float32") -> np.ndarray:
    """Row-parallel (input dimension split) quantized Dense under tensor parallelism, src/layers/common.cc:348-401:
    the input rows are quantized with the amax of the WHOLE row (the reference all-gathers the activations before
    Quantize, :360-387), rank r multiplies its K slice, dequantizes its partial to T, the partials are summed
    (ops::ReduceAll, transformer.cc:45-48 / attention.cc:608-612) and the residual is added once (rank 0, :348-352).
    `dtype` is the activation type the partials are rounded to before the sum."""
    shape = x.shape
    x2 = x.reshape(-1, shape[-1])
    xq, xs = quantize_rows(x2)
    k = x2.shape[1]
    total = np.zeros((x2.shape[0], w_q.shape[0]), f32)
    for r in range(world):
        b, e = tp_shard_rows(k, r, world)
        part = dequantize_gemm_output(gemm_s8(xq[:, b:e], w_q[:, b:e]), xs, w_scale, None, ACT_NONE, flavor)
        if dtype != "float32":
            part = part.astype(np.float16).astype(f32) if dtype == "float16" else part
        total = (total + part).astype(f32)
    if residual is not None:
        total = (total + residual.reshape(total.shape).astype(f32)).astype(f32)
    return total.reshape(shape[:-1] + (w_q.shape[0],))


def quantize_weight(w: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Weight quantization done by the converter / on load: python/ctranslate2/specs/model_spec.py:222-243
    == src/models/model.cc:304-369.  scale[i] = 127/amax(W[i,:]) (amax 0 -> 127), W_q = rint(W*scale)."""
    amax = np.max(np.abs(w), axis=1).astype(f32)
    amax = np.where(amax == 0, f32(127), amax)
    scale = (f32(127) / amax).astype(f32)
    q = np.rint(w.astype(f32) * scale[:, None]).astype(np.int8)
    return q, scale


# --------------------------------------------------------------------------------------
# Norms / rotary / softmax / topk / gather  
# --------------------------------------------------------------------------------------

def rms_norm(x: np.ndarray, gamma: np.ndarray, eps: float = 1e-6, use_residual: bool = False) -> np.ndarray:
    """ops::RMSNorm.  src/ops/rms_norm_gpu.cu:19-63 / src/cpu/kernels.cc rms_norm:
    y = x * rsqrt(mean(x^2) + eps) * gamma   (gamma -> 1 + gamma when use_residual)."""
    x = x.astype(f32)
    ms = np.sum(x * x, axis=-1, dtype=f32, keepdims=True) / f32(x.shape[-1])
    inv = (f32(1) / np.sqrt(ms + f32(eps), dtype=f32)).astype(f32)
    g = gamma.astype(f32) + (f32(1) if use_residual else f32(0))
    return (x * inv * g).astype(f32)


def rotary_tables(num_positions: int, dim: int, base: float = 10000.0, interleave: bool = False,
                  scaling_type: int = -1, scaling_factor: float = 1.0,
                  low_freq_factor: float = 1.0, high_freq_factor: float = 4.0,
                  original_max_position_embeddings: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """layers::RotaryEmbeddings::initialize.  src/layers/attention_layer.cc:252-343 (None / Linear=0 /
    Llama3=2 scaling; Su is out of scope).  Returns (sin, cos) [num_positions, dim] float32."""
    i = np.arange(dim // 2, dtype=f32)
    inv_freq = (f32(1) / np.power(f32(base), (i * f32(2)) / f32(dim), dtype=f32)).astype(f32)
    if scaling_type == 2:  # Llama3, attention_layer.cc:271-294
        old_len = f32(original_max_position_embeddings)
        low_wavelen = old_len / f32(low_freq_factor)
        high_wavelen = old_len / f32(high_freq_factor)
        new = inv_freq.copy()
        for j in range(inv_freq.size):
            wavelen = f32(2.0 * math.pi) / inv_freq[j]
            if wavelen < high_wavelen:
                pass
            elif wavelen > low_wavelen:
                new[j] = inv_freq[j] / f32(scaling_factor)
            else:
                smooth = (old_len / wavelen - f32(low_freq_factor)) / (f32(high_freq_factor) - f32(low_freq_factor))
                new[j] = (f32(1) - smooth) * inv_freq[j] / f32(scaling_factor) + smooth * inv_freq[j]
        inv_freq = new.astype(f32)
    t = np.arange(num_positions, dtype=f32)
    if scaling_type == 0:  # Linear
        t = (t / f32(scaling_factor)).astype(f32)
    freqs = (t[:, None] * inv_freq[None, :]).astype(f32)
    if interleave:
        emb = np.repeat(freqs, 2, axis=1)
    else:
        emb = np.concatenate([freqs, freqs], axis=1)
    return np.sin(emb, dtype=f32), np.cos(emb, dtype=f32)


def rotary(x: np.ndarray, sin: np.ndarray, cos: np.ndarray, interleave: bool = False) -> np.ndarray:
    """ops::Rotary.  src/ops/rotary_cpu.cc:8-37 == src/ops/rotary_gpu.cu:27-85.
    x [..., T, D]; sin/cos [T, ndims] (ndims <= D; trailing dims are copied).
    interleave:      y[i] = x[i]*c[i] + (i even ? -x[i+1] : x[i-1]) * s[i]
    non-interleaved: y[i] = x[i]*c[i] + (i < n/2 ? -x[i+n/2] : x[i-n/2]) * s[i]"""
    x = x.astype(f32)
    nd = sin.shape[-1]
    xr = x[..., :nd]
    rot = np.empty_like(xr)
    if interleave:
        rot[..., 0::2] = -xr[..., 1::2]
        rot[..., 1::2] = xr[..., 0::2]
    else:
        h = nd // 2
        rot[..., :h] = -xr[..., h:]
        rot[..., h:] = xr[..., :h]
    y = x.copy()
    y[..., :nd] = (xr * cos.astype(f32) + rot * sin.astype(f32)).astype(f32)
    return y


def softmax(x: np.ndarray, lengths: Optional[np.ndarray] = None, log: bool = False) -> np.ndarray:
    """ops::SoftMax / LogSoftMax over the last axis with optional per-row valid lengths.
    src/ops/softmax.cc:28-47, src/ops/softmax_gpu.cu:190-256, src/cpu/kernels.cc softmax:
    only the first lengths[row] columns participate; the masked tail of the output is 0."""
    x2 = x.reshape(-1, x.shape[-1]).astype(f32)
    out = np.zeros_like(x2)
    for r in range(x2.shape[0]):
        n = x2.shape[1] if lengths is None else int(np.asarray(lengths).reshape(-1)[r])
        if n == 0:
            continue
        row = x2[r, :n]
        m = row.max()
        e = np.exp(row - m, dtype=f32)
        s = e.sum(dtype=f32)
        out[r, :n] = (row - m - np.log(s, dtype=f32)) if log else e / s
    return out.reshape(x.shape)


def topk(x: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """ops::TopK over the last axis.  src/ops/topk.cc:14-22, CPU src/ops/topk_cpu.cc:12-55
    (k=1: std::max_element => lowest index wins ties; k>1: descending values).  The product kernel
    defines ties as lowest-index-first for every k; values are returned unchanged."""
    x2 = x.reshape(-1, x.shape[-1])
    order = np.argsort(-x2.astype(np.float64), axis=1, kind="stable")[:, :k]
    vals = np.take_along_axis(x2, order, axis=1)
    return vals.reshape(x.shape[:-1] + (k,)), order.astype(np.int32).reshape(x.shape[:-1] + (k,))


def gather_rows(data: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """ops::Gather(axis=0, batch_dims=0).  src/ops/gather.cc:49-86 — pure copy."""
    return data[np.asarray(ids).astype(np.int64)]


# --------------------------------------------------------------------------------------
# AWQ-INT4  (SURVEY §8 a7).  The reference has NO CPU implementation and NO tests for AWQ;
# this restates src/ops/awq/dequantize.cuh:14-77 (nibble order) + dequantize_gpu.cu:8-62 (GEMM
# layout) and src/ops/awq/gemv_gpu.cu:289-422 (GEMV layout).  Parity for AWQ is pinned only by
# self-consistency between the two layouts and the pack/unpack round trip.
# --------------------------------------------------------------------------------------

AWQ_ORDER = np.array([0, 4, 1, 5, 2, 6, 3, 7])  # output column 8c+i lives in nibble AWQ_ORDER[i]


def awq_pack_gemm(w_int: np.ndarray, zeros_int: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Pack AWQ_GEMM layout: w_int [K,N] in 0..15 -> qweight int32 [K,N/8]; zeros [K/G,N] -> [K/G,N/8]."""
    def pack(m):
        r, n = m.shape
        m = m.reshape(r, n // 8, 8).astype(np.uint32)
        out = np.zeros((r, n // 8), np.uint32)
        for i in range(8):
            out |= (m[:, :, i] & 0xF) << np.uint32(4 * AWQ_ORDER[i])
        return out.view(np.int32)
    return pack(w_int), pack(zeros_int)


def awq_unpack_gemm(q: np.ndarray) -> np.ndarray:
    u = q.view(np.uint32)
    r, c = u.shape
    out = np.zeros((r, c, 8), np.int32)
    for i in range(8):
        out[:, :, i] = (u >> np.uint32(4 * AWQ_ORDER[i])) & 0xF
    return out.reshape(r, c * 8)


def awq_dequantize_gemm(qweight: np.ndarray, scales: np.ndarray, qzeros: np.ndarray) -> np.ndarray:
    """ops::DequantizeAwq (AWQ_GEMM layout) -> W [K,N] float32: (nibble - zero) * scale, group = K / scales.rows.
    src/ops/awq/dequantize_gpu.cu:8-62 (computed there in fp16: sub.f16x2 then fma.rn.f16x2)."""
    w = awq_unpack_gemm(qweight).astype(f32)
    z = awq_unpack_gemm(qzeros).astype(f32)
    g = w.shape[0] // scales.shape[0]
    return ((w - np.repeat(z, g, axis=0)) * np.repeat(scales.astype(f32), g, axis=0)).astype(f32)


def awq_gemm(x: np.ndarray, qweight: np.ndarray, scales: np.ndarray, qzeros: np.ndarray) -> np.ndarray:
    """ops::GemmAwq: y[M,N] = x[M,K] @ deq(W)[K,N].  src/ops/awq/gemm.cc:8-33."""
    return (x.astype(f32) @ awq_dequantize_gemm(qweight, scales, qzeros)).astype(f32)


def awq_gemv_widths(ic: int, group: int) -> Tuple[int, int]:
    """zeros_w / sf_w padding rules of the AWQ_GEMV layout.  src/ops/awq/gemv_gpu.cu:305-307, 380-382."""
    div = lambda c, d: (c + d - 1) // d
    if group == 64:
        zw = div(div(ic // 64, 8), 2) * 2
    else:
        zw = div(ic // group, 8)
    return zw, zw * 8


def awq_pack_gemv(w_int: np.ndarray, zeros_int: np.ndarray, scales: np.ndarray, group: int):
    """Pack AWQ_GEMV layout: w_int [OC,IC] -> [OC,IC/8] (nibble i = input channel 8w+i, sequential);
    zeros [OC,IC/G] -> [OC,zeros_w]; scales [OC,IC/G] -> fp16 [OC,sf_w]."""
    oc, ic = w_int.shape
    zw, sw = awq_gemv_widths(ic, group)
    m = w_int.reshape(oc, ic // 8, 8).astype(np.uint32)
    qw = np.zeros((oc, ic // 8), np.uint32)
    for i in range(8):
        qw |= (m[:, :, i] & 0xF) << np.uint32(4 * i)
    ng = ic // group
    zpad = np.zeros((oc, zw * 8), np.uint32)
    zpad[:, :ng] = zeros_int
    zp = zpad.reshape(oc, zw, 8)
    qz = np.zeros((oc, zw), np.uint32)
    for i in range(8):
        qz |= (zp[:, :, i] & 0xF) << np.uint32(4 * i)
    sc = np.zeros((oc, sw), np.float16)
    sc[:, :ng] = scales
    return qw.view(np.int32), qz.view(np.int32), sc


def awq_gemv(x: np.ndarray, qweight: np.ndarray, scales: np.ndarray, qzeros: np.ndarray, group: int) -> np.ndarray:
    """ops::GemvAwq: y[M,OC] = x[M,IC] @ deq(W)^T with W in AWQ_GEMV layout; fp32 scale*(nibble-zero), fp32 FMA.
    src/ops/awq/gemv_gpu.cu:289-422."""
    u = qweight.view(np.uint32)
    oc, w8 = u.shape
    ic = w8 * 8
    w = np.zeros((oc, w8, 8), f32)
    for i in range(8):
        w[:, :, i] = ((u >> np.uint32(4 * i)) & 0xF).astype(f32)
    w = w.reshape(oc, ic)
    ng = ic // group
    uz = qzeros.view(np.uint32)
    z = np.zeros((oc, uz.shape[1], 8), f32)
    for i in range(8):
        z[:, :, i] = ((uz >> np.uint32(4 * i)) & 0xF).astype(f32)
    z = z.reshape(oc, -1)[:, :ng]
    s = scales.astype(f32)[:, :ng]
    deq = (np.repeat(s, group, axis=1) * (w - np.repeat(z, group, axis=1))).astype(f32)
    return (x.astype(f32) @ deq.T).astype(f32)


# --------------------------------------------------------------------------------------
# model.bin reader (src/models/model.cc:561-660; writer python/ctranslate2/specs/model_spec.py:382-414)
# --------------------------------------------------------------------------------------

_DTYPES = {0: np.float32, 1: np.int8, 2: np.int16, 3: np.int32, 4: np.float16, 5: np.uint16}  # 5 = bfloat16 bits


def read_model_bin(path: str) -> Tuple[str, int, Dict[str, np.ndarray], Dict[str, str]]:
    import struct
    variables: Dict[str, np.ndarray] = {}
    aliases: Dict[str, str] = {}
    with open(path, "rb") as f:
        def rd(fmt):
            size = struct.calcsize(fmt)
            return struct.unpack(fmt, f.read(size))[0]

        def rd_str():
            n = rd("H")
            s = f.read(n)
            return s[:-1].decode("utf-8")

        version = rd("I")
        spec = rd_str() if version >= 2 else ""
        revision = rd("I") if version >= 2 else 1
        nvars = rd("I")
        for _ in range(nvars):
            name = rd_str()
            rank = rd("B")
            dims = [rd("I") for _ in range(rank)]
            if version >= 4:
                type_id = rd("B")
                nbytes = rd("I")
                dt = _DTYPES[type_id]
            else:
                item = rd("B")
                nbytes = rd("I") * item
                dt = {4: np.float32, 2: np.int16, 1: np.int8}[item]
            buf = f.read(nbytes)
            arr = np.frombuffer(buf, dtype=dt).reshape(dims).copy()
            if version >= 4 and type_id == 5:      # bfloat16 bit patterns -> float32
                arr = (arr.astype(np.uint32) << np.uint32(16)).view(np.float32)
            variables[name] = arr
        if version >= 3:
            for _ in range(rd("I")):
                alias = rd_str()
                aliases[alias] = rd_str()
    for a, t in aliases.items():
        variables[a] = variables[t]
    return spec, revision, variables, aliases


# --------------------------------------------------------------------------------------
# Llama-class decoder and greedy search (a17)
# --------------------------------------------------------------------------------------

@dataclass
class DecoderWeights:
    """The variables layers::TransformerDecoder reads for a pre-norm / RMSNorm / SwiGLU / RoPE decoder
    (src/layers/transformer.cc:473-535, attention_layer.cc:112-142)."""
    v: Dict[str, np.ndarray]
    num_layers: int
    num_heads: int
    num_heads_kv: int
    head_dim: int
    eps: float
    rotary_base: float
    rotary_interleave: bool
    rotary_scaling_type: int = -1
    rotary_scaling_factor: float = 1.0
    rotary_low_freq_factor: float = 1.0
    rotary_high_freq_factor: float = 4.0
    original_max_position_embeddings: int = 0
    flavor: str = "cpu"   # dequantize arithmetic flavor
    awq_layout: int = 0   # config.json quantization_type: 1 = AWQ_GEMM, 2 = AWQ_GEMV (src/models/model.cc:636-637)
    awq_group: int = 128

    @staticmethod
    def from_dir(model_dir: str, flavor: str = "cpu") -> "DecoderWeights":
        import json, os
        _, _, v, _ = read_model_bin(os.path.join(model_dir, "model.bin"))
        cfg = {}
        cpath = os.path.join(model_dir, "config.json")
        if os.path.exists(cpath):
            cfg = json.load(open(cpath))
        L = 0
        while f"decoder/layer_{L}/self_attention/linear_0/weight" in v:
            L += 1
        H = int(v["decoder/num_heads"])
        a = "decoder/layer_0/self_attention/"
        Hkv = int(v.get(a + "num_heads_kv", np.array(H)))
        d_model = v["decoder/embeddings/weight"].shape[1]
        head_dim = int(v[a + "head_dim"]) if a + "head_dim" in v else d_model // H
        get = lambda k, d: (v[a + k].item() if a + k in v else d)
        return DecoderWeights(
            v=v, num_layers=L, num_heads=H, num_heads_kv=Hkv, head_dim=head_dim,
            eps=float(cfg.get("layer_norm_epsilon") or 1e-6),   # model.cc / common.cc:455-462 default 1e-6
            rotary_base=float(get("rotary_base", 10000.0)),
            rotary_interleave=bool(get("rotary_interleave", True)),
            rotary_scaling_type=int(get("rotary_scaling_type", -1)),
            rotary_scaling_factor=float(get("rotary_scaling_factor", 1.0)),
            rotary_low_freq_factor=float(get("rotary_low_freq_factor", 1.0)),
            rotary_high_freq_factor=float(get("rotary_high_freq_factor", 4.0)),
            original_max_position_embeddings=int(get("original_max_position_embeddings", 0)),
            flavor=flavor, awq_layout=int(cfg.get("quantization_type") or 0),
            awq_group=int(cfg.get("quantization_group_size") or 128))


def ensure_compute_type(w: "DecoderWeights", weights: str, float_dtype=np.float32) -> "DecoderWeights":
    """Model::set_compute_type + ensure_dtype (src/models/model.cc:178-234, 304-369) for the variables of a decoder:
    every variable whose name ends in "weight" (is_quantizable, :288-290) is brought to the weight type of the requested
    compute type — weights="int8": a float matrix is quantized row-wise on its float32 value with ops::Quantize
    (scale = 127/amax or 1, q = rint(w*scale)) and "<name>_scale" is registered; weights="float": an int8 matrix becomes
    float_dtype(float32(q) * (1/scale)) (the CPU Dequantize, dequantize_cpu.cc:12-21, the load happens on the CPU) and its
    scale is removed.  AWQ variables (int32) are left alone (model.cc:750-757)."""
    import dataclasses
    v = dict(w.v)
    for name in list(v):
        if name not in v:
            continue
        a = v[name]
        if not name.endswith("weight") or a.ndim != 2 or a.dtype == np.int32:
            continue
        if weights == "int8" and a.dtype != np.int8:
            q, sc = quantize_rows(a.astype(f32))
            v[name], v[name + "_scale"] = q, sc
        elif weights == "float" and a.dtype == np.int8:
            r = (f32(1) / v[name + "_scale"].astype(f32)).astype(f32)
            v[name] = (a.astype(f32) * r[:, None]).astype(f32).astype(float_dtype)
            del v[name + "_scale"]
    return dataclasses.replace(w, v=v)


def layer_norm(x: np.ndarray, gamma: np.ndarray, beta: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """ops::LayerNorm over the last axis.  src/cpu/kernels.cc:463-495: mean = sum/n, var = max(sum(x^2)/n - mean^2, 0),
    y = (x - mean) / sqrt(var + eps) * gamma + beta  (eps 1e-5 when the layer has a beta, layers/common.cc:449-453)."""
    x = x.astype(f32)
    n = f32(x.shape[-1])
    mean = (x.sum(-1, dtype=f32) / n).astype(f32)
    var = np.maximum((x * x).sum(-1, dtype=f32) / n - mean * mean, f32(0)).astype(f32)
    rstd = (f32(1) / np.sqrt(var + f32(eps), dtype=f32)).astype(f32)
    return ((x - mean[..., None]) * rstd[..., None] * gamma.astype(f32) + beta.astype(f32)).astype(f32)


def sinusoidal_position_encoding(max_time: int, depth: int) -> np.ndarray:
    """layers::SinusoidalPositionEncoder.  src/layers/common.cc:204-229: positions start at 1, [sin | cos] halves,
    timescale_j = exp(-j * log(10000) / (depth/2 - 1))."""
    inc = f32(math.log(10000.0)) / f32(depth // 2 - 1)
    timescales = np.exp(np.arange(depth // 2, dtype=f32) * -inc, dtype=f32)
    scaled = (np.arange(1, max_time + 1, dtype=f32)[:, None] * timescales[None, :]).astype(f32)
    return np.concatenate([np.sin(scaled, dtype=f32), np.cos(scaled, dtype=f32)], axis=1)


class Seq2SeqOracle:
    """fp32 restatement of the encoder-decoder Transformer behind ctranslate2::Translator (the golden model of
    tests/translator_test.cc:53-96): TransformerEncoder (src/layers/transformer.cc:405-471), TransformerDecoder with
    cross-attention (:621-871, attention.cc:371-440), pre-norm LayerNorm, ReLU FFN, sinusoidal positions, embeddings scaled by
    sqrt(d) (:382-402), INT8 Dense with bias, beam search (beam_search above).  `v` = variables of model.bin."""

    def __init__(self, variables: Dict[str, np.ndarray], num_heads: int = 8, flavor: str = "cpu",
                 binary_version: int = 6, compute_type: str = "int8", eps: float = 1e-5):
        self.v = variables
        self.flavor = flavor
        self.round_before_cast = binary_version >= 5
        # compute_type "float32": Model::ensure_dtype dequantizes the saved int8 weights (q / scale, model.cc:330-341) and
        # layers::Dense runs its float arm -- no activation quantization, hence no rounding flips: the strict structural pin
        self.compute_type = compute_type
        self._float_w: Dict[str, np.ndarray] = {}
        self.num_heads = int(variables.get("encoder/num_heads", np.int16(num_heads)))
        self.enc_emb = "encoder/embeddings_0" if "encoder/embeddings_0/weight" in variables else "encoder/embeddings"
        self.d = variables["decoder/embeddings/weight"].shape[1]
        self.enc_layers = 0
        while f"encoder/layer_{self.enc_layers}/ffn/linear_0/weight" in variables:
            self.enc_layers += 1
        self.dec_layers = 0
        while f"decoder/layer_{self.dec_layers}/ffn/linear_0/weight" in variables:
            self.dec_layers += 1
        self.pos = sinusoidal_position_encoding(512, self.d)
        # attributes of TransformerSpec (scoped since spec revision 5, models/transformer.cc:67-79)
        def attr(scope, name, default):
            for key in (f"{scope}/{name}", name):
                if key in variables:
                    return variables[key]
            return default
        self.pre_norm = {s: bool(attr(s, "pre_norm", True)) for s in ("encoder", "decoder")}
        self.act = {s: int(attr(s, "activation", ACT_RELU)) for s in ("encoder", "decoder")}
        self.zero_first = bool(variables.get("decoder/start_from_zero_embedding", False))
        self.eps = eps

        def emb_scale(scope):                                    # build_embeddings_scale, transformer.cc:380-402
            sc = variables.get(scope + "/scale_embeddings", variables.get(scope + "/embeddings/multiply_by_sqrt_depth"))
            if sc is None or (sc.dtype == np.int8 and bool(sc)):
                return f32(math.sqrt(self.d))
            if sc.dtype != np.int8 and float(sc) != 1.0:
                return f32(sc)
            return None
        self.emb_scale = {s: emb_scale(s) for s in ("encoder", "decoder")}
        if "decoder/position_encodings/encodings" in variables:  # PositionEmbedding instead of the sinusoidal encoder
            self.pos = variables["decoder/position_encodings/encodings"].astype(f32)

    @classmethod
    def from_dir(cls, model_dir: str, flavor: str = "cpu", compute_type: str = "int8") -> "Seq2SeqOracle":
        _, _, variables, _ = read_model_bin(model_dir + "/model.bin")
        with open(model_dir + "/model.bin", "rb") as f:
            binary_version = struct.unpack("<I", f.read(4))[0]
        eps = 1e-5
        cfg = os.path.join(model_dir, "config.json")
        if os.path.exists(cfg):
            with open(cfg) as f:
                e = json.load(f).get("layer_norm_epsilon")
            eps = eps if e is None else float(e)
        return cls(variables, flavor=flavor, binary_version=binary_version, compute_type=compute_type, eps=eps)

    # -- layers -----------------------------------------------------------------------
    def _dense(self, prefix, x, act=ACT_NONE, residual=None):
        v = self.v
        if self.compute_type == "float32":
            w = self._float_w.get(prefix)
            if w is None:
                w = v[prefix + "/weight"].astype(f32)
                if prefix + "/weight_scale" in v:                      # int8: per-row scales; int16: one scale per layer
                    sc = v[prefix + "/weight_scale"].astype(f32)
                    w = (w / (sc[:, None] if sc.ndim == 1 else sc)).astype(f32)
                self._float_w[prefix] = w
            return gemm_float(x.reshape(-1, x.shape[-1]), w, trans_b=True, bias=v.get(prefix + "/bias"),
                              residual=None if residual is None else residual.reshape(-1, w.shape[0]),
                              act=act).reshape(x.shape[:-1] + (w.shape[0],))
        return dense_int8(x, v[prefix + "/weight"], v[prefix + "/weight_scale"], v.get(prefix + "/bias"), act, residual,
                          self.flavor, self.round_before_cast)

    def _embed(self, scope, ids):
        v = self.v
        emb = self.enc_emb if scope == "encoder" else "decoder/embeddings"
        x = gather_rows(v[emb + "/weight"], ids).astype(f32)
        if emb + "/weight_scale" in v:                           # Embeddings::operator(), common.cc:64-81
            sc = v[emb + "/weight_scale"].astype(f32)
            x = (x / (gather_rows(sc, ids)[..., None] if sc.ndim == 1 else sc)).astype(f32)
        sc = self.emb_scale[scope]
        return x if sc is None else (x * sc).astype(f32)

    def _ln(self, prefix, x):
        return layer_norm(x, self.v[prefix + "/gamma"], self.v[prefix + "/beta"], self.eps)

    def _sublayer(self, scope, prefix, x, fn):
        """pre-norm: x + f(LN(x)); post-norm: LN(x + f(x)) (attention.cc:497-500, 602-614; transformer.cc:21-51)."""
        if self.pre_norm[scope]:
            return fn(self._ln(prefix + "/layer_norm", x), x)
        return self._ln(prefix + "/layer_norm", fn(x, x))

    def _attend(self, q, k, v_, lens_rows):
        """q [B,T,d], k/v [B,S,d] -> context [B,T,d]; softmax over the first lens_rows keys of each (b, h, t) row."""
        B, T, _ = q.shape
        S = k.shape[1]
        H, D = self.num_heads, self.d // self.num_heads
        qh = q.reshape(B, T, H, D).transpose(0, 2, 1, 3)
        kh = k.reshape(B, S, H, D).transpose(0, 2, 1, 3)
        vh = v_.reshape(B, S, H, D).transpose(0, 2, 1, 3)
        scores = (np.einsum("bhtd,bhsd->bhts", qh, kh) * f32(1.0 / math.sqrt(D))).astype(f32)
        probs = softmax(scores.reshape(-1, S), lens_rows).reshape(B, H, T, S)
        ctx = np.einsum("bhts,bhsd->bhtd", probs, vh).astype(f32)
        return ctx.transpose(0, 2, 1, 3).reshape(B, T, self.d)

    # -- encoder ----------------------------------------------------------------------
    def encode(self, src: np.ndarray, lengths: np.ndarray) -> np.ndarray:
        """src [B,S] ids (padding ignored through `lengths`) -> memory [B,S,d]."""
        B, S = src.shape
        x = (self._embed("encoder", src) + self.pos[:S][None]).astype(f32)
        lens_rows = np.repeat(lengths, self.num_heads * S)
        for l in range(self.enc_layers):
            p = f"encoder/layer_{l}/"

            def attn(h, res):
                q, k, v_ = np.split(self._dense(p + "self_attention/linear_0", h), 3, axis=-1)
                return self._dense(p + "self_attention/linear_1", self._attend(q, k, v_, lens_rows), residual=res)

            def ffn(h, res):
                return self._dense(p + "ffn/linear_1", self._dense(p + "ffn/linear_0", h, act=self.act["encoder"]), residual=res)

            x = self._sublayer("encoder", p + "self_attention", x, attn)
            x = self._sublayer("encoder", p + "ffn", x, ffn)
        return self._ln("encoder/layer_norm", x) if "encoder/layer_norm/gamma" in self.v else x

    # -- decoder ----------------------------------------------------------------------
    def start(self, memory: np.ndarray, lengths: np.ndarray, beam_size: int):
        """Decoder state for B * beam_size rows: empty self-attention caches, projected memory keys / values per layer."""
        memory = np.repeat(memory, beam_size, axis=0)
        self.mem_lengths = np.repeat(lengths, beam_size)
        self.self_k = [np.zeros((memory.shape[0], 0, self.d), f32) for _ in range(self.dec_layers)]
        self.self_v = [np.zeros((memory.shape[0], 0, self.d), f32) for _ in range(self.dec_layers)]
        self.mem_k, self.mem_v = [], []
        for l in range(self.dec_layers):
            kv = self._dense(f"decoder/layer_{l}/attention/linear_1", memory)
            k, v_ = np.split(kv, 2, axis=-1)
            self.mem_k.append(k)
            self.mem_v.append(v_)

    def reorder(self, index: np.ndarray):
        self.self_k = [k[index] for k in self.self_k]
        self.self_v = [v_[index] for v_ in self.self_v]
        self.mem_k = [k[index] for k in self.mem_k]
        self.mem_v = [v_[index] for v_ in self.mem_v]
        self.mem_lengths = self.mem_lengths[index]

    def step(self, ids: np.ndarray, step: int) -> np.ndarray:
        """One target position for every row: ids [N] at position `step` -> logits [N, V]."""
        N = ids.shape[0]
        H = self.num_heads
        x = self._embed("decoder", ids.reshape(N, 1))
        if self.zero_first and step == 0:                        # start_from_zero_embedding (transformer.cc:637-640)
            x = np.zeros_like(x)
        x = (x + self.pos[step:step + 1][None]).astype(f32)
        for l in range(self.dec_layers):
            p = f"decoder/layer_{l}/"

            def self_attn(h, res):
                q, k, v_ = np.split(self._dense(p + "self_attention/linear_0", h), 3, axis=-1)
                self.self_k[l] = np.concatenate([self.self_k[l], k], axis=1)
                self.self_v[l] = np.concatenate([self.self_v[l], v_], axis=1)
                S = self.self_k[l].shape[1]
                ctx = self._attend(q, self.self_k[l], self.self_v[l], np.full(N * H, S))
                return self._dense(p + "self_attention/linear_1", ctx, residual=res)

            def cross_attn(h, res):
                q = self._dense(p + "attention/linear_0", h)
                ctx = self._attend(q, self.mem_k[l], self.mem_v[l], np.repeat(self.mem_lengths, H))
                return self._dense(p + "attention/linear_2", ctx, residual=res)

            def ffn(h, res):
                return self._dense(p + "ffn/linear_1", self._dense(p + "ffn/linear_0", h, act=self.act["decoder"]), residual=res)

            x = self._sublayer("decoder", p + "self_attention", x, self_attn)
            x = self._sublayer("decoder", p + "attention", x, cross_attn)
            x = self._sublayer("decoder", p + "ffn", x, ffn)
        if "decoder/layer_norm/gamma" in self.v:
            x = self._ln("decoder/layer_norm", x)
        return self._dense("decoder/projection", x)[:, 0, :]

    # -- Translator::translate_batch ---------------------------------------------------
    def translate(self, source_ids: Sequence[Sequence[int]], beam_size: int = 2, num_hypotheses: int = 1,
                  max_length: int = 256, min_length: int = 1, length_penalty: float = 1.0, bos: int = 1, eos: int = 2):
        """TranslationOptions defaults (include/ctranslate2/translation.h): beam 2, length_penalty 1, min_decoding_length 1."""
        B = len(source_ids)
        lengths = np.array([len(r) for r in source_ids])
        S = int(lengths.max())
        src = np.zeros((B, S), np.int64)
        for b, r in enumerate(source_ids):
            src[b, :len(r)] = r
        memory = self.encode(src, lengths)
        self.start(memory, lengths, beam_size)
        V = self.v["decoder/projection/weight"].shape[0]
        return beam_search(self.step, self.reorder, np.full(B, bos), V, beam_size, max_length, min_length, [eos],
                           length_penalty, num_hypotheses)


def conv1d(x: np.ndarray, w: np.ndarray, bias: Optional[np.ndarray], stride: int, padding: int) -> np.ndarray:
    """ops::Conv1D (src/ops/conv1d.cc, conv1d_cpu.cc:128-220 im2col + GEMM; conv1d_gpu.cu cuDNN): x [B, Cin, T], w [Cout, Cin, K]
    -> [B, Cout, (T + 2 * padding - K) / stride + 1]."""
    B, Cin, T = x.shape
    Cout, _, K = w.shape
    xp = np.zeros((B, Cin, T + 2 * padding), f32)
    xp[:, :, padding:padding + T] = x
    Tout = (T + 2 * padding - K) // stride + 1
    cols = np.stack([xp[:, :, k:k + stride * Tout:stride] for k in range(K)], axis=2)      # [B, Cin, K, Tout]
    y = np.einsum("bckt,ock->bot", cols, w.astype(f32)).astype(f32)
    return y if bias is None else (y + bias.astype(f32)[None, :, None]).astype(f32)


class WhisperOracle(Seq2SeqOracle):
    """fp32 restatement of models::WhisperReplica (src/models/whisper.cc:106-390) without timestamp rules: WhisperEncoder
    (src/layers/whisper.cc:8-63: conv1 + GELU, conv2 stride 2 + GELU, stored positions, pre-norm GELU layers, LayerNorm), the
    TransformerDecoder with cross-attention and stored positions, forward_prompt on the task tokens, then the search with
    SuppressTokens / SuppressTokensBegin and hypotheses that exclude the end token."""

    def __init__(self, variables, config: Optional[dict] = None, **kw):
        super().__init__(variables, **kw)
        self.config = config or {}
        self.pre_norm["encoder"], self.act["encoder"] = True, ACT_GELU
        vocab = self.v["decoder/embeddings/weight"].shape[0]
        self.vocab = vocab

    @classmethod
    def from_dir(cls, model_dir: str, flavor: str = "cpu", compute_type: str = "int8") -> "WhisperOracle":
        _, _, variables, aliases = read_model_bin(model_dir + "/model.bin")
        for alias, target in aliases.items():
            variables[alias] = variables[target]
            if target + "_scale" in variables:
                variables[alias + "_scale"] = variables[target + "_scale"]
        with open(os.path.join(model_dir, "config.json")) as f:
            config = json.load(f)
        with open(os.path.join(model_dir, "vocabulary.json")) as f:
            tokens = json.load(f)
        o = cls(variables, config=config, flavor=flavor, binary_version=6, compute_type=compute_type)
        o.sot, o.eot = tokens.index("<|startoftranscript|>"), tokens.index("<|endoftext|>")
        o.no_timestamps = tokens.index("<|notimestamps|>")
        o.no_speech = tokens.index("<|nospeech|>") if "<|nospeech|>" in tokens else tokens.index("<|nocaptions|>")
        return o

    def encode_features(self, features: np.ndarray) -> np.ndarray:
        """features [B, n_mels, T] -> [B, T / 2, d]."""
        v = self.v
        x = activation(conv1d(features.astype(f32), v["encoder/conv1/weight"], v["encoder/conv1/bias"], 1, 1), ACT_GELU)
        x = activation(conv1d(x, v["encoder/conv2/weight"], v["encoder/conv2/bias"], 2, 1), ACT_GELU)
        x = np.ascontiguousarray(x.transpose(0, 2, 1))
        B, S, _ = x.shape
        x = (x + v["encoder/position_encodings/encodings"].astype(f32)[:S][None]).astype(f32)
        lens_rows = np.full(B * self.num_heads * S, S)
        for l in range(self.enc_layers):
            p = f"encoder/layer_{l}/"

            def attn(h, res):
                q, k, v_ = np.split(self._dense(p + "self_attention/linear_0", h), 3, axis=-1)
                return self._dense(p + "self_attention/linear_1", self._attend(q, k, v_, lens_rows), residual=res)

            def ffn(h, res):
                return self._dense(p + "ffn/linear_1", self._dense(p + "ffn/linear_0", h, act=ACT_GELU), residual=res)

            x = self._sublayer("encoder", p + "self_attention", x, attn)
            x = self._sublayer("encoder", p + "ffn", x, ffn)
        return self._ln("encoder/layer_norm", x)

    def generate(self, features: np.ndarray, prompts: np.ndarray, beam_size: int = 5, patience: float = 1.0,
                 num_hypotheses: int = 1, length_penalty: float = 1.0, max_length: int = 448, suppress_blank: bool = True,
                 suppress_default: bool = True, max_initial_timestamp_index: int = 50):
        """prompts [B, P]: <|startoftranscript|> + task tokens (no text after them); ApplyTimestampRules (whisper.cc:742-860)
        unless the last one is <|notimestamps|>.  Returns (per entry [(tokens, score), ...] best first, no_speech_probs [B])."""
        prompts = np.asarray(prompts)
        B, P = prompts.shape
        assert P >= 2
        timestamps = prompts[0, -1] != self.no_timestamps
        ts_begin, ts_end = self.no_timestamps + 1, self.vocab - 1
        memory = self.encode_features(features)
        self.start(memory, np.full(B, memory.shape[1]), beam_size)
        no_speech = np.zeros(B, f32)
        for t in range(P - 1):                                   # WhisperDecoder::forward_prompt on prompt[:-1]
            logits = self.step(np.repeat(prompts[:, t], beam_size), t)
            if (prompts[:, t] == self.sot).all():
                no_speech = softmax(logits[::beam_size])[:, self.no_speech]
        start_step = P - 1
        steps = min(max_length // 2, max_length - start_step)
        disable = list(self.config.get("suppress_ids", [])) if suppress_default else []
        disable_begin = list(self.config.get("suppress_ids_begin", [])) if suppress_blank else []
        lowest = np.finfo(f32).min

        state = {"seq": [[] for _ in range(B * beam_size)]}

        def hook(step, logits):                                  # SuppressTokens, SuppressTokensBegin (decoding_utils.cc:152-188)
            for t in disable:
                logits[:, t] = lowest
            if step == 0:
                for t in disable_begin:
                    logits[:, t] = lowest
            if not timestamps:
                return
            check = []
            for n in range(logits.shape[0]):                     # ApplyTimestampRules::apply, sample_begin = 0
                seq = state["seq"][n]
                logits[n, self.no_timestamps] = lowest
                if step == 0:
                    logits[n, :ts_begin] = lowest
                    logits[n, ts_begin + max_initial_timestamp_index + 1:ts_end + 1] = lowest
                else:
                    last = seq[step - 1]
                    if last >= ts_begin:
                        penult = seq[step - 2] if step - 1 > 0 else last
                        if penult >= ts_begin:
                            logits[n, ts_begin:ts_end + 1] = lowest
                        else:
                            logits[n, :self.eot] = lowest
                            logits[n, ts_begin:last] = lowest
                            check.append(n)
                    else:
                        check.append(n)
                        for t in range(step - 1, -1, -1):
                            if seq[t] >= ts_begin:
                                logits[n, ts_begin:seq[t] + 1] = lowest
                                break
            if check:
                with np.errstate(over="ignore"):
                    lp = softmax(logits, log=True)
                for n in check:
                    ts = lp[n, ts_begin:ts_end + 1]
                    mx = ts.max()
                    if f32(mx + np.log(np.exp(ts - mx, dtype=f32).sum(dtype=f32))) > lp[n, :ts_begin].max():
                        logits[n, :ts_begin] = lowest

        def step_fn(ids, s):
            if s > 0:                                            # the search gathered the beams: ids are the new last tokens
                state["seq"] = [state["seq"][g] + [int(t)] for g, t in zip(state["gather"], ids)]
            return self.step(ids, start_step + s)

        def reorder(index):
            state["gather"] = [int(i) for i in index]
            self.reorder(index)

        res = beam_search(step_fn, reorder, prompts[:, -1], self.vocab, beam_size, steps, 0, [self.eot], length_penalty,
                          num_hypotheses, patience, include_eos_in_hypotheses=False, logits_hook=hook)
        return res, no_speech


def apply_logits_processors(logits: np.ndarray, sampled: Sequence[int], repetition_penalty: float = 1.0,
                            no_repeat_ngram_size: int = 0, suppress_sequences: Sequence[Sequence[int]] = (),
                            disable_ids: Sequence[int] = ()) -> None:
    """The LogitsProcessor chain of one row, in the order decoding.cc:1099-1112 builds it, on `logits` [V] in place.
    `sampled` = the row of alive_seq: every token chosen by the loop so far (empty at step 0).  src/decoding_utils.cc:
      RepetitionPenalty (:45-67, cpu/primitives.cc:411-428): scores of previous tokens are gathered FIRST, then each is
        rewritten from its gathered value (score < 0 ? score * p : score / p) — a token seen twice is penalised once;
      NoRepeatNgram (:75-107): every token that completed an earlier occurrence of the current (n-1)-gram suffix is disabled;
      SuppressTokens (:170-177, disable_unk) and SuppressSequences (:110-150): single tokens always, the last token of a
        longer sequence when the row ends with the rest.
    Disabled entries become the lowest float (DisableTokens, decoding_utils.h:39) after all processors ran."""
    lowest = np.finfo(f32).min
    disabled = set()
    if repetition_penalty != 1.0 and len(sampled) > 0:
        ids = np.asarray(sampled, np.int64)
        prev = logits[ids].copy()
        logits[ids] = np.where(prev < 0, prev * f32(repetition_penalty), prev / f32(repetition_penalty)).astype(f32)
    n = int(no_repeat_ngram_size)
    if n > 0 and len(sampled) >= n:
        suffix = list(sampled[len(sampled) - n + 1:]) if n > 1 else []
        for p in range(0, len(sampled) - n + 1):
            if list(sampled[p:p + n - 1]) == suffix:
                disabled.add(int(sampled[p + n - 1]))
    for t in disable_ids:
        disabled.add(int(t))
    for seq in suppress_sequences:
        seq = [int(t) for t in seq]
        if len(seq) == 0:
            continue
        if len(seq) == 1:
            disabled.add(seq[0])
        elif len(sampled) >= len(seq) - 1 and len(sampled) > 0 and list(sampled[len(sampled) - (len(seq) - 1):]) == seq[:-1]:
            disabled.add(seq[-1])
    for t in disabled:
        logits[t] = lowest


def beam_search(step_fn, reorder_fn, start_ids: np.ndarray, vocab: int, beam_size: int, max_length: int,
                min_length: int = 0, end_ids: Sequence[int] = (), length_penalty: float = 1.0, num_hypotheses: int = 1,
                patience: float = 1.0, include_eos_in_hypotheses: bool = True, logits_hook=None):
    """BeamSearch::search (src/decoding.cc:425-720).  No prefix bias, no coverage penalty; hypotheses keep their end token
    while they are scored (include_eos_in_hypotheses = true, decoding.h:154) and lose it in the returned tokens.
    step_fn(ids [B*beam], step) -> logits [B*beam, vocab] (advances the decoder state); reorder_fn(index [B*beam]) gathers the
    state rows (Decoder::update_state).  Returns per batch entry a list of (tokens, score), best first
    (finalize_result / sort_hypotheses, :189-254).
      * 2 * beam_size candidates per step from TopK over the flattened [beam * vocab] cumulative log-probabilities;
      * only beam 0 is live at step 0 (initialize_beam_scores: the others start at the lowest float, :84-93);
      * a candidate among the first beam_size that ends (end token, or last step) is registered as a hypothesis and its
        slot is refilled with the next non-end candidate of the secondary list (:617-655);
      * an entry is finished at the last step, or — with no length penalty — when its top beam ended and it has
        num_hypotheses hypotheses, or else when it has round(beam_size * patience) hypotheses (:657-663)."""
    B, V = len(start_ids), vocab
    end_ids = list(end_ids)
    ids = np.repeat(np.asarray(start_ids), beam_size).astype(np.int64)               # [B * beam]
    lowest = np.finfo(f32).min
    scores = np.tile(np.array([0.0] + [lowest] * (beam_size - 1), f32), B)           # initialize_beam_scores
    alive = [[[] for _ in range(beam_size)] for _ in range(B)]
    hyps: List[List[Tuple[List[int], float]]] = [[] for _ in range(B)]
    finished = [False] * B
    top_done = [False] * B
    ncand = 2 * beam_size
    max_candidates = int(math.floor(beam_size * patience + 0.5))                      # std::round: half away from zero
    early_exit = length_penalty == 0
    for step in range(max_length):
        logits = np.array(step_fn(ids, step), f32)                                    # [B*beam, V]
        if step < min_length:
            for e in end_ids:
                logits[:, e] = lowest                                                 # apply_min_length + DisableTokens
        if logits_hook is not None:
            logits_hook(step, logits)                                                 # LogitsProcessor chain (SuppressTokens ...)
        with np.errstate(over="ignore"):
            lp = (softmax(logits, log=True) + scores[:, None]).astype(f32).reshape(B, beam_size * V)
        cand_scores, cand_ids = topk(lp, ncand)
        origin, word = cand_ids // V, cand_ids % V
        is_last = step + 1 == max_length
        new_ids = np.zeros((B, beam_size), np.int64)
        new_scores = np.zeros((B, beam_size), f32)
        gather = np.zeros((B, beam_size), np.int64)
        for i in range(B):
            seqs = [alive[i][int(origin[i, j])] + [int(word[i, j])] for j in range(ncand)]
            secondary = beam_size
            active = []
            for k in range(beam_size):
                nxt = k
                if not finished[i] and (int(word[i, k]) in end_ids or is_last):
                    if k == 0:
                        top_done[i] = True
                    drop = int(word[i, k]) in end_ids and not include_eos_in_hypotheses   # decoding.cc:601-603
                    hyps[i].append((seqs[k][:step] if drop else seqs[k][:step + 1], float(cand_scores[i, k])))
                    for j in range(secondary, ncand):
                        if int(word[i, j]) not in end_ids:
                            nxt, secondary = j, j + 1
                            break
                active.append(nxt)
            if not finished[i]:
                if is_last:
                    finished[i] = True
                elif early_exit:
                    finished[i] = top_done[i] and len(hyps[i]) >= num_hypotheses
                else:
                    finished[i] = len(hyps[i]) >= max_candidates
            alive[i] = [seqs[a] for a in active]
            new_ids[i] = word[i, active]
            new_scores[i] = cand_scores[i, active]
            gather[i] = i * beam_size + origin[i, active]
        if all(finished):
            break
        reorder_fn(gather.reshape(-1))
        ids, scores = new_ids.reshape(-1), new_scores.reshape(-1).astype(f32)
    out = []
    for i in range(B):
        with np.errstate(divide="ignore"):
            final = [(t, float(f32(sc) / f32(f32(len(t)) ** f32(length_penalty)))) for t, sc in hyps[i]]
        order = sorted(range(len(final)), key=lambda j: -final[j][1])                  # std::sort, descending score
        best = []
        for j in order[:num_hypotheses]:
            t = list(final[j][0])
            while t and t[-1] in end_ids:
                t.pop()
            best.append((t, final[j][1]))
        out.append(best)
    return out


class LlamaOracle:
    """fp32 restatement of TransformerDecoder::decode (src/layers/transformer.cc:621-871) for the
    Llama family: Embeddings (common.cc:64-81) -> L x [MultiHeadAttention (attention.cc:442-615) ->
    FeedForwardNetwork (transformer.cc:21-51)] -> output RMSNorm -> projection Dense."""

    def __init__(self, w: DecoderWeights):
        self.w = w
        self._sin = None
        self._cos = None
        self.reset(0)

    # -- state ------------------------------------------------------------------------
    def reset(self, batch: int):
        w = self.w
        self.k_cache = [np.zeros((batch, w.num_heads_kv, 0, w.head_dim), f32) for _ in range(w.num_layers)]
        self.v_cache = [np.zeros((batch, w.num_heads_kv, 0, w.head_dim), f32) for _ in range(w.num_layers)]

    def _tables(self, n):
        w = self.w
        if self._sin is None or self._sin.shape[0] < n:
            self._sin, self._cos = rotary_tables(
                max(n, 64), w.head_dim, w.rotary_base, w.rotary_interleave, w.rotary_scaling_type,
                w.rotary_scaling_factor, w.rotary_low_freq_factor, w.rotary_high_freq_factor,
                w.original_max_position_embeddings)
        return self._sin, self._cos

    # -- layers -----------------------------------------------------------------------
    def _dense(self, prefix: str, x: np.ndarray, act: int = ACT_NONE, residual=None) -> np.ndarray:
        v = self.w.v
        wq = v[prefix + "/weight"]
        bias = v.get(prefix + "/bias")
        if wq.dtype == np.int8:
            return dense_int8(x, wq, v[prefix + "/weight_scale"], bias, act, residual, self.w.flavor)
        if wq.dtype == np.int32:                  # AWQ arms, common.cc:402-438 (weights dequantized to fp16)
            sc, zr = v[prefix + "/weight_scale"], v[prefix + "/weight_zero"]
            x2 = x.reshape(-1, x.shape[-1])
            if self.w.awq_layout == 1:
                deq = awq_dequantize_gemm(wq, sc, zr).astype(np.float16).astype(f32)          # [K, N]
                y = x2.astype(f32) @ deq
            else:
                y = awq_gemv(x2, wq, sc, zr, self.w.awq_group)
            y = y.reshape(x.shape[:-1] + (y.shape[-1],))
            # apply_bias_and_activation (gemm.cc:10-25): act(y + bias + residual).  No layer of the decoder passes an
            # activation together with a residual, so this equals the product's fused epilogue (activation, then residual).
            if bias is not None:
                y = y + bias.astype(f32)
            if residual is not None:
                y = y + residual
            return activation(y.astype(f32), act)
        return gemm_float(x.reshape(-1, x.shape[-1]), wq, trans_b=True, bias=bias,          # float arm, common.cc:440
                          residual=None if residual is None else residual.reshape(-1, wq.shape[0]),
                          act=act).reshape(x.shape[:-1] + (wq.shape[0],))

    def _embed(self, ids: np.ndarray) -> np.ndarray:
        v = self.w.v
        e = v["decoder/embeddings/weight"]
        rows = gather_rows(e, ids)
        if e.dtype == np.int8:   # common.cc:64-81: gather int8 rows + scales, Dequantize: x / scale
            sc = gather_rows(v["decoder/embeddings/weight_scale"], ids)
            rows = (rows.astype(f32) / sc[..., None].astype(f32)).astype(f32)
        return rows.astype(f32)

    def forward(self, ids: np.ndarray, offset: int = 0, all_logits: bool = True) -> np.ndarray:
        """ids [B,T] at positions offset..offset+T-1 (all rows share `offset`, as in the reference,
        flash_attention_gpu.cu:269).  Appends K/V to the cache.  Returns logits [B,T,V] (or [B,1,V])."""
        w = self.w
        B, T = ids.shape
        H, Hkv, D = w.num_heads, w.num_heads_kv, w.head_dim
        sin, cos = self._tables(offset + T)
        sin, cos = sin[offset:offset + T], cos[offset:offset + T]
        x = self._embed(ids)                                                    # [B,T,d]
        for l in range(w.num_layers):
            p = f"decoder/layer_{l}/"
            a = p + "self_attention/"
            h = rms_norm(x, w.v[a + "layer_norm/gamma"], w.eps)
            qkv = self._dense(a + "linear_0", h)                                # [B,T,(H+2Hkv)D]
            q = qkv[..., :H * D].reshape(B, T, H, D).transpose(0, 2, 1, 3)
            k = qkv[..., H * D:(H + Hkv) * D].reshape(B, T, Hkv, D).transpose(0, 2, 1, 3)
            vv = qkv[..., (H + Hkv) * D:].reshape(B, T, Hkv, D).transpose(0, 2, 1, 3)
            q = rotary(q, sin, cos, w.rotary_interleave)
            k = rotary(k, sin, cos, w.rotary_interleave)
            self.k_cache[l] = np.concatenate([self.k_cache[l], k], axis=2)
            self.v_cache[l] = np.concatenate([self.v_cache[l], vv], axis=2)
            K, V = self.k_cache[l], self.v_cache[l]                             # [B,Hkv,S,D]
            S = K.shape[2]
            g = H // Hkv
            Kr = np.repeat(K, g, axis=1)                                        # replicate_heads, attention.cc:291-295
            Vr = np.repeat(V, g, axis=1)
            scores = (np.einsum("bhtd,bhsd->bhts", q, Kr) * f32(1.0 / math.sqrt(D))).astype(f32)
            # causal mask: query t (absolute offset+t) sees keys 0..offset+t  (attention_layer.cc:152-174)
            lens = np.minimum(np.arange(T) + offset + 1, S)
            lens_rows = np.broadcast_to(lens, (B, H, T)).reshape(-1)
            probs = softmax(scores.reshape(-1, S), lens_rows).reshape(B, H, T, S)
            ctx = np.einsum("bhts,bhsd->bhtd", probs, Vr).astype(f32)
            ctx = ctx.transpose(0, 2, 1, 3).reshape(B, T, H * D)
            x = self._dense(a + "linear_1", ctx, residual=x)
            f = p + "ffn/"
            h = rms_norm(x, w.v[f + "layer_norm/gamma"], w.eps)
            gate = self._dense(f + "linear_0", h, act=ACT_SWISH)
            up = self._dense(f + "linear_0_noact", h)
            x = self._dense(f + "linear_1", (gate * up).astype(f32), residual=x)
        if not all_logits:
            x = x[:, -1:, :]
        x = rms_norm(x, w.v["decoder/layer_norm/gamma"], w.eps)
        return self._dense("decoder/projection", x)

    # -- beam search ------------------------------------------------------------------
    def generate_beam(self, prompts: np.ndarray, beam_size: int, max_length: int, min_length: int = 0,
                      end_ids: Sequence[int] = (), length_penalty: float = 1.0, num_hypotheses: int = 1,
                      patience: float = 1.0):
        """Generator::generate_batch with beam_size > 1: the prompt pass of language_model.cc:217-238 (all but the last prompt
        token), state replicated beam_size times, then beam_search() below from the last prompt token."""
        B, P = prompts.shape
        V = self.w.v["decoder/projection/weight"].shape[0]
        self.reset(B)
        if P > 1:
            self.forward(prompts[:, :P - 1], 0, all_logits=False)
        self.k_cache = [np.repeat(k, beam_size, axis=0) for k in self.k_cache]        # replicate_state
        self.v_cache = [np.repeat(v, beam_size, axis=0) for v in self.v_cache]

        def step_fn(ids, step):
            return self.forward(ids.reshape(-1, 1), P - 1 + step, all_logits=False)[:, 0, :]

        def reorder_fn(g):                                                            # Decoder::update_state
            self.k_cache = [k[g] for k in self.k_cache]
            self.v_cache = [v[g] for v in self.v_cache]

        return beam_search(step_fn, reorder_fn, prompts[:, P - 1], V, beam_size, max_length, min_length, end_ids,
                           length_penalty, num_hypotheses, patience)

    # -- greedy search ----------------------------------------------------------------
    def score(self, sequences: Sequence[Sequence[int]], offset: int = 0) -> List[List[float]]:
        """Generator::score_batch.  src/scoring.cc:6-66: inputs = sequence[:-1], outputs = sequence[1:], one full-sequence
        forward, LogSoftMax, Gather of the output ids; results start at `offset`; sequences of fewer than two tokens score
        nothing (language_model.cc:136-140).  Rows are scored one by one here (padding never reaches a valid position of a
        causal decoder, so the batch composition does not matter)."""
        out: List[List[float]] = []
        for seq in sequences:
            seq = [int(t) for t in seq]
            if len(seq) < 2:
                out.append([])
                continue
            self.reset(1)
            logits = self.forward(np.array([seq[:-1]]), 0)[0]                  # [T, V]
            lp = softmax(logits, log=True)
            out.append([float(lp[t, seq[t + 1]]) for t in range(offset, len(seq) - 1)])
        return out

    def generate(self, prompts: np.ndarray, max_length: int, min_length: int = 0,
                 end_ids: Sequence[int] = (), return_scores: bool = False, length_penalty: float = 1.0,
                 repetition_penalty: float = 1.0, no_repeat_ngram_size: int = 0,
                 suppress_sequences: Sequence[Sequence[int]] = (), disable_ids: Sequence[int] = ()):
        """Generator::generate_batch, greedy, include_prompt_in_result=false.
        src/models/language_model.cc:217-238 (prefill of P-1 tokens) + GreedySearch::search
        (src/decoding.cc:732-974): argmax = TopK k=1 lowest-index ties; EOS forbidden until min_length;
        a finished row stops (its tokens are not reported further).  The logits processors of GenerationOptions
        (decoding.cc:1099-1112, apply_logits_processors below) see the tokens sampled so far in the loop — NOT the prompt."""
        B, P = prompts.shape
        processors = repetition_penalty != 1.0 or no_repeat_ngram_size > 0 or len(suppress_sequences) > 0 or len(disable_ids) > 0
        sampled: List[List[int]] = [[] for _ in range(B)]      # alive_seq, decoding.cc:886-893 (includes an end token)
        self.reset(B)
        if P > 1:
            self.forward(prompts[:, :P - 1], 0, all_logits=False)
        cur = prompts[:, P - 1:P].copy()
        out: List[List[int]] = [[] for _ in range(B)]
        done = [False] * B
        ended_by_eos = [False] * B
        scores = np.zeros(B, np.float64)
        for step in range(max_length):
            logits = self.forward(cur, P - 1 + step, all_logits=False)[:, 0, :]
            if step < min_length:
                for e in end_ids:
                    logits[:, e] = np.finfo(f32).min     # DisableTokens, decoding.cc:852-856
            if processors:
                for b in range(B):
                    if not done[b]:
                        apply_logits_processors(logits[b], sampled[b], repetition_penalty, no_repeat_ngram_size,
                                                suppress_sequences, disable_ids)
            _, idx = topk(logits, 1)
            nxt = idx[:, 0]
            if return_scores:
                # decoding.cc:875-880: LogSoftMax over the processed logits, score += log-prob of the sampled token
                # (the end token's too: the addition precedes the is_finished test, :919-923)
                lp = softmax(logits, log=True)
            for b in range(B):
                if done[b]:
                    continue
                tok = int(nxt[b])
                sampled[b].append(tok)
                if return_scores:
                    scores[b] += float(lp[b, tok])
                if tok in end_ids:
                    done[b] = True
                    ended_by_eos[b] = True
                else:
                    out[b].append(tok)
                    if len(out[b]) >= max_length:
                        done[b] = True
            cur = nxt.reshape(B, 1).astype(np.int64)
            if all(done):
                break
        if return_scores:
            # finalize_hypothesis_score, decoding.cc:189-203: score / length^length_penalty.  The Generator decodes with
            # include_eos_in_hypotheses = true (decoding.h:154) and strips the end token afterwards
            # (language_model.cc:253-257), so the length that normalises the score counts the end token.
            final = [float(scores[b] / ((len(out[b]) + (1 if ended_by_eos[b] else 0)) ** length_penalty))
                     for b in range(B)]
            return out, np.array(final, np.float32)
        return out
