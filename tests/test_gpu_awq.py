"""-m gpu: AWQ-INT4 kernels vs the oracle's restatement of the reference's dequantize.cuh / gemv_gpu.cu formulas.
(The reference has neither a CPU implementation nor tests for AWQ; the oracle is pinned by layout round trips.)"""
import numpy as np
import pytest
import torch

from ctranslate2_b200 import ops
from oracle import ct2_oracle as O
from gpu_util import DEV, dev, gpu, to_np


def make_awq(n, k, g, seed):
    r = np.random.default_rng(seed)
    w_int = r.integers(0, 16, size=(k, n))
    z_int = r.integers(0, 16, size=(k // g, n))
    scales = r.uniform(0.002, 0.02, size=(k // g, n)).astype(np.float16)
    deq = ((w_int - np.repeat(z_int, g, 0)).astype(np.float32) * np.repeat(scales.astype(np.float32), g, 0))
    deq = deq.astype(np.float16).astype(np.float32)          # (q - z) * s rounded once to fp16, as the reference does
    return w_int, z_int, scales, deq                          # deq [k, n]


def pack(w_int, z_int, scales, g, layout):
    if layout == ops.AWQ_GEMM:
        qw, qz = O.awq_pack_gemm(w_int, z_int)
        return qw, scales, qz
    qw, qz, sc = O.awq_pack_gemv(w_int.T.copy(), z_int.T.copy(), scales.T.copy(), g)
    return qw, sc, qz


@gpu
@pytest.mark.parametrize("layout", [ops.AWQ_GEMM, ops.AWQ_GEMV])
@pytest.mark.parametrize("nkg", [(256, 512, 128), (1024, 4096, 128), (384, 1024, 64)])
def test_dequantize_awq_exact(layout, nkg):
    n, k, g = nkg
    w_int, z_int, scales, deq = make_awq(n, k, g, n + layout)
    qw, sc, qz = pack(w_int, z_int, scales, g, layout)
    w = ops.dequantize_awq(dev(qw), dev(sc), dev(qz), layout, g)
    np.testing.assert_array_equal(to_np(w), deq)               # [k, n], bit-exact fp16


@gpu
@pytest.mark.parametrize("layout", [ops.AWQ_GEMM, ops.AWQ_GEMV])
@pytest.mark.parametrize("mnkg", [(1, 256, 512, 128), (7, 1024, 4096, 128), (32, 4096, 4096, 128), (64, 384, 1024, 64),
                                  (200, 512, 1024, 128)])
def test_dense_awq(layout, mnkg):
    m, n, k, g = mnkg
    w_int, z_int, scales, deq = make_awq(n, k, g, m + n)
    qw, sc, qz = pack(w_int, z_int, scales, g, layout)
    wt = ops.AwqWeight(dev(qw), dev(sc), dev(qz), layout, g)
    r = np.random.default_rng(m)
    x = r.standard_normal((m, k)).astype(np.float16)
    bias = r.standard_normal(n).astype(np.float16)
    res = r.standard_normal((m, n)).astype(np.float16)
    ref = x.astype(np.float64) @ deq.astype(np.float64)
    y = ops.dense_awq(dev(x), wt)
    np.testing.assert_allclose(to_np(y), ref, rtol=1e-2, atol=1e-2 * np.abs(ref).max())
    y = ops.dense_awq(dev(x), wt, bias=dev(bias), residual=dev(res), activation_type=ops.ActivationType.Swish)
    v = ref.astype(np.float32) + bias.astype(np.float32)
    ref2 = O.activation(v, O.ACT_SWISH) + res.astype(np.float32)
    np.testing.assert_allclose(to_np(y), ref2, rtol=1e-2, atol=1e-2 * max(1.0, np.abs(ref2).max()))
    # deterministic: split tiles are reduced in a fixed order
    assert torch.equal(ops.dense_awq(dev(x), wt), ops.dense_awq(dev(x), wt))
    # oracle formulas of both reference arms agree with it as well
    if layout == ops.AWQ_GEMM:
        np.testing.assert_allclose(O.awq_gemm(x, qw, scales, qz), ref, rtol=2e-3, atol=2e-3 * np.abs(ref).max())
    else:
        np.testing.assert_allclose(O.awq_gemv(x, qw, sc, qz, g), ref, rtol=2e-3, atol=2e-3 * np.abs(ref).max())


@gpu
@pytest.mark.parametrize("m", [1, 32, 100])
def test_dense_awq_glu(m):
    n, k, g = 1024, 512, 128
    wg = make_awq(n, k, g, 1)
    wu = make_awq(n, k, g, 2)
    g_w = ops.AwqWeight(*[dev(a) for a in pack(*wg[:3], g, ops.AWQ_GEMM)], ops.AWQ_GEMM, g)
    u_w = ops.AwqWeight(*[dev(a) for a in pack(*wu[:3], g, ops.AWQ_GEMV)], ops.AWQ_GEMV, g)
    x = np.random.default_rng(m).standard_normal((m, k)).astype(np.float16)
    h = ops.dense_awq_glu(dev(x), g_w, u_w)
    gate = O.activation((x.astype(np.float64) @ wg[3].astype(np.float64)).astype(np.float32), O.ACT_SWISH)
    up = (x.astype(np.float64) @ wu[3].astype(np.float64)).astype(np.float32)
    ref = gate * up
    np.testing.assert_allclose(to_np(h), ref, rtol=2e-2, atol=2e-2 * np.abs(ref).max())


@gpu
@pytest.mark.parametrize("cs", [1, 2, 3, 4])
@pytest.mark.parametrize("rows", [128, 96])
def test_awq_decode_plans(cs, rows, monkeypatch):
    """awq_decode.cu: every cluster size (split-K through DSMEM) and a reduced tile height, pinned with
    CT2B200_GEMM_CS / CT2B200_GEMM_ROWS; also against the general kernel (CT2B200_AWQ_DECODE=0)."""
    n, k, g = 1000 - 1000 % 8, 2048, 128
    w_int, z_int, scales, deq = make_awq(n, k, g, cs * 10 + rows)
    qw, sc, qz = pack(w_int, z_int, scales, g, ops.AWQ_GEMM)
    wt = ops.AwqWeight(dev(qw), dev(sc), dev(qz), ops.AWQ_GEMM, g)
    wu = make_awq(n, k, g, 3)
    u_w = ops.AwqWeight(*[dev(a) for a in pack(*wu[:3], g, ops.AWQ_GEMM)], ops.AWQ_GEMM, g)
    r = np.random.default_rng(cs)
    for m in (1, 17, 33, 64):
        x = r.standard_normal((m, k)).astype(np.float16)
        res = r.standard_normal((m, n)).astype(np.float16)
        monkeypatch.setenv("CT2B200_AWQ_DECODE", "0")
        y_general = to_np(ops.dense_awq(dev(x), wt, residual=dev(res)))
        monkeypatch.setenv("CT2B200_AWQ_DECODE", "1")
        monkeypatch.setenv("CT2B200_GEMM_CS", str(cs))
        monkeypatch.setenv("CT2B200_GEMM_ROWS", str(rows))
        y = to_np(ops.dense_awq(dev(x), wt, residual=dev(res)))
        h = to_np(ops.dense_awq_glu(dev(x), wt, u_w))
        monkeypatch.delenv("CT2B200_GEMM_CS")
        monkeypatch.delenv("CT2B200_GEMM_ROWS")
        ref = x.astype(np.float64) @ deq.astype(np.float64) + res.astype(np.float64)
        np.testing.assert_allclose(y, ref, rtol=1e-2, atol=1e-2 * max(1.0, np.abs(ref).max()))
        np.testing.assert_allclose(y, y_general, rtol=4e-3, atol=4e-3 * max(1.0, np.abs(ref).max()))
        gate = O.activation((x.astype(np.float64) @ deq.astype(np.float64)).astype(np.float32), O.ACT_SWISH)
        up = (x.astype(np.float64) @ wu[3].astype(np.float64)).astype(np.float32)
        np.testing.assert_allclose(h, gate * up, rtol=2e-2, atol=2e-2 * np.abs(gate * up).max())
