"""-m gpu: op-level parity of the CUDA kernels (through the C-ABI) against the oracle and the reference's
golden vectors.  Integer / index results are bit-exact; floats within the reference's own tolerances."""
import json
import os

import numpy as np
import pytest
import torch

from ctranslate2_b200 import ops
from oracle import ct2_oracle as O
from gpu_util import DEV, TDT, TOL, dev, gpu, round_through, to_np

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
GT = json.load(open(os.path.join(GOLDEN, "ref_gtest_vectors.json")))
RND = np.load(os.path.join(GOLDEN, "ref_ops_random.npz"))
DTYPES = ["float32", "float16", "bfloat16"]
IMPLS = [ops.GEMM_TCGEN05, ops.GEMM_MMA_SYNC]


def vec(test, name, dtype=np.float32, nth=0):
    items = [i for i in GT[test] if i["name"] == name]
    return np.array(items[nth]["values"], dtype=dtype).reshape(items[nth]["shape"])


# ---------------- Quantize ----------------
@gpu
@pytest.mark.parametrize("dt", DTYPES)
def test_quantize_gtest_golden(dt):
    for test in ("QuantizeINT8", "QuantizeINT8ZeroRow"):
        a = vec(test, "a")
        q, s = ops.Quantize(True)(dev(a, TDT[dt]))
        np.testing.assert_array_equal(to_np(q), vec(test, "expected_qa", np.int8, 0))
        np.testing.assert_allclose(to_np(s), vec(test, "expected_scale"), rtol=1e-6)
        q, _ = ops.Quantize(False)(dev(a, TDT[dt]))
        np.testing.assert_array_equal(to_np(q), vec(test, "expected_qa", np.int8, 2 if test == "QuantizeINT8" else 1))


@gpu
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(1, 8), (5, 96), (3, 4096), (33, 14336), (7, 1001), (0, 64)])
def test_quantize_bit_exact(dt, shape):
    r = np.random.default_rng(sum(shape))
    x = round_through(r.standard_normal(shape) * 3, dt)
    if shape[0] > 2:
        x[2] = 0
    q, s = ops.Quantize()(dev(x, TDT[dt]))
    qo, so = O.quantize_rows(x)
    np.testing.assert_array_equal(to_np(q), qo)
    np.testing.assert_array_equal(to_np(s), so)


@gpu
def test_quantize_reference_fixture():
    q, s = ops.Quantize()(dev(RND["q_x"]))
    np.testing.assert_array_equal(to_np(q), RND["q_q"])
    np.testing.assert_array_equal(to_np(s), RND["q_s"])


# ---------------- INT8 GEMM ----------------
@gpu
@pytest.mark.parametrize("impl", IMPLS)
def test_gemm_int8_gtest_golden(impl):
    a, b = vec("GemmInt8", "a", np.int8), vec("GemmInt8", "b", np.int8)
    a16 = np.zeros((3, 16), np.int8); a16[:, :8] = a          # k must be a multiple of 16: zero-pad K
    b16 = np.zeros((4, 16), np.int8); b16[:, :8] = b.T
    c = ops.Gemm(impl=impl)(dev(a16), dev(b16))
    np.testing.assert_array_equal(to_np(c), vec("GemmInt8", "expected", np.int32))


@gpu
@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("mnk", [(1, 128, 128), (1, 4096, 4096), (5, 24, 96), (16, 6144, 4096), (17, 300, 1040),
                                 (32, 4096, 14336), (33, 1000, 512), (64, 2048, 4096), (100, 200, 304),
                                 (128, 256, 128), (300, 1000, 2048), (1024, 6144, 4096)])
def test_gemm_int8_exact(impl, mnk):
    m, n, k = mnk
    g = torch.Generator(device=DEV).manual_seed(m * 7 + n)
    a = torch.randint(-127, 128, (m, k), device=DEV, dtype=torch.int8, generator=g)
    b = torch.randint(-127, 128, (n, k), device=DEV, dtype=torch.int8, generator=g)
    c = ops.Gemm(impl=impl)(a, b)
    ref = (a.double() @ b.double().T).to(torch.int32)      # exact: |sum| < 2^53
    assert torch.equal(c, ref), f"max abs diff {(c - ref).abs().max().item()}"
    # the split-K scratch must be left clean: a second call gives the same answer
    assert torch.equal(ops.Gemm(impl=impl)(a, b), ref)


@gpu
def test_gemm_int8_reference_fixture():
    for impl in IMPLS:
        c = ops.Gemm(impl=impl)(dev(RND["g_a"]), dev(RND["g_b"]))
        np.testing.assert_array_equal(to_np(c), RND["g_c"])


# ---------------- Dequantize / fused Dense ----------------
@gpu
@pytest.mark.parametrize("dt", DTYPES)
def test_dequantize_gemm_output(dt):
    c, sa, sb, bias = RND["g_c"], RND["dq_sa"], RND["dq_sb"], round_through(RND["dq_bias"], dt)
    for act in (-1, 0, 1, 2, 3, 4, 5, 6):
        y = ops.Dequantize(None if act < 0 else act)(dev(c), dev(sa), dev(sb), dev(bias, TDT[dt]), dtype=TDT[dt])
        ref = O.dequantize_gemm_output(c, sa, sb, bias, act, "cuda")
        tol = max(TOL[dt], 3e-5)
        np.testing.assert_allclose(to_np(y), ref, rtol=tol, atol=tol * max(1.0, np.abs(ref).max()))
        if dt == "float32":   # against the unmodified reference's output
            np.testing.assert_allclose(to_np(y), RND["dq_y_act%d" % act], rtol=3e-5, atol=3e-5)


@gpu
@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("mnk", [(1, 256, 512), (8, 1000, 1024), (32, 4096, 4096), (200, 512, 768), (100, 520, 2048)])
def test_dense_int8_fused(impl, dt, mnk):
    """Dense = Quantize -> Gemm -> Dequantize(+bias, act) -> Add(residual), one launch, vs the oracle."""
    m, n, k = mnk
    r = np.random.default_rng(m + n)
    x = round_through(r.standard_normal((m, k)), dt)
    w = (r.standard_normal((n, k)) * 0.05).astype(np.float32)
    wq, ws = O.quantize_weight(w)
    bias = round_through(r.standard_normal(n) * 0.1, dt)
    res = round_through(r.standard_normal((m, n)), dt)
    xq, xs = ops.Quantize()(dev(x, TDT[dt]))
    for act, use_bias, use_res in ((-1, False, False), (ops.ActivationType.Swish, False, False), (-1, True, True),
                                   (ops.ActivationType.GELU, True, False)):
        y = ops.dense_int8(xq, xs, dev(wq), dev(ws), dev(bias, TDT[dt]) if use_bias else None,
                           dev(res, TDT[dt]) if use_res else None, None if act < 0 else act, TDT[dt], impl)
        ref = O.dense_int8(x, wq, ws, bias if use_bias else None, act, res if use_res else None, "cuda")
        tol = TOL[dt] if dt != "float32" else 2e-5
        np.testing.assert_allclose(to_np(y), ref, rtol=tol, atol=tol * max(1.0, float(np.abs(ref).max())))


@gpu
@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("dt", ["float16", "float32"])
@pytest.mark.parametrize("mnk", [(1, 512, 256), (32, 1024, 512), (300, 768, 512)])
def test_dense_int8_glu(impl, dt, mnk):
    m, n, k = mnk
    r = np.random.default_rng(n)
    x = round_through(r.standard_normal((m, k)), dt)
    wg, sg = O.quantize_weight((r.standard_normal((n, k)) * 0.05).astype(np.float32))
    wu, su = O.quantize_weight((r.standard_normal((n, k)) * 0.05).astype(np.float32))
    xq, xs = ops.Quantize()(dev(x, TDT[dt]))
    h = ops.dense_int8_glu(xq, xs, dev(wg), dev(sg), dev(wu), dev(su), ops.ActivationType.Swish, TDT[dt], impl)
    gate = round_through(O.dense_int8(x, wg, sg, None, O.ACT_SWISH), dt)
    up = round_through(O.dense_int8(x, wu, su), dt)
    ref = gate * up
    tol = TOL[dt] if dt != "float32" else 2e-5
    np.testing.assert_allclose(to_np(h), ref, rtol=tol, atol=tol * max(1.0, float(np.abs(ref).max())))


@gpu
@pytest.mark.parametrize("dt", ["float16", "bfloat16"])
@pytest.mark.parametrize("mnk", [(1, 256, 512), (40, 1000, 1024), (300, 512, 264)])
def test_gemm_f16_tcgen05(dt, mnk):
    m, n, k = mnk
    r = np.random.default_rng(k)
    a = round_through(r.standard_normal((m, k)), dt)
    b = round_through(r.standard_normal((n, k)) * 0.05, dt)
    bias = round_through(r.standard_normal(n), dt)
    c = ops.Gemm()(dev(a, TDT[dt]), dev(b, TDT[dt]), bias=dev(bias, TDT[dt]))
    ref = a.astype(np.float64) @ b.astype(np.float64).T + bias
    np.testing.assert_allclose(to_np(c), ref, rtol=TOL[dt], atol=TOL[dt] * float(np.abs(ref).max()))


@gpu
@pytest.mark.parametrize("cs", [1, 2, 3, 4])
@pytest.mark.parametrize("rows", [128, 104, 97, 64])
def test_decode_gemm_plans(cs, rows, monkeypatch):
    """gemm_decode.cu: every cluster size (split-K through DSMEM) and tile height the planner can pick, pinned with
    CT2B200_GEMM_CS / CT2B200_GEMM_ROWS; fp32 output => any lost or doubled partial shows up at 2e-5."""
    monkeypatch.setenv("CT2B200_GEMM_CS", str(cs))
    monkeypatch.setenv("CT2B200_GEMM_ROWS", str(rows))
    r = np.random.default_rng(cs * 1000 + rows)
    n, k = 1000, 1536                                    # ragged last tile; 12 K blocks
    w = (r.standard_normal((n, k)) * 0.05).astype(np.float32)
    wq, ws = O.quantize_weight(w)
    wu, su = O.quantize_weight((r.standard_normal((n, k)) * 0.05).astype(np.float32))
    for m in (1, 16, 17, 33, 64):
        for dt in ("float32", "float16"):
            x = round_through(r.standard_normal((m, k)), dt)
            res = round_through(r.standard_normal((m, n)), dt)
            xq, xs = ops.Quantize()(dev(x, TDT[dt]))
            tol = TOL[dt] if dt != "float32" else 2e-5
            y = ops.dense_int8(xq, xs, dev(wq), dev(ws), None, dev(res, TDT[dt]), None, TDT[dt], ops.GEMM_TCGEN05)
            ref = O.dense_int8(x, wq, ws, None, -1, res, "cuda")
            np.testing.assert_allclose(to_np(y), ref, rtol=tol, atol=tol * max(1.0, float(np.abs(ref).max())))
            h = ops.dense_int8_glu(xq, xs, dev(wq), dev(ws), dev(wu), dev(su), ops.ActivationType.Swish, TDT[dt],
                                   ops.GEMM_TCGEN05)
            gate = round_through(O.dense_int8(x, wq, ws, None, O.ACT_SWISH), dt)
            ref = gate * round_through(O.dense_int8(x, wu, su), dt)
            np.testing.assert_allclose(to_np(h), ref, rtol=tol, atol=tol * max(1.0, float(np.abs(ref).max())))
        a = round_through(r.standard_normal((m, k)), "float16")
        b = round_through(r.standard_normal((n, k)) * 0.05, "float16")
        c = ops.Gemm()(dev(a, TDT["float16"]), dev(b, TDT["float16"]))
        ref = a.astype(np.float64) @ b.astype(np.float64).T
        np.testing.assert_allclose(to_np(c), ref, rtol=TOL["float16"], atol=TOL["float16"] * float(np.abs(ref).max()))


@gpu
@pytest.mark.parametrize("dt", ["float32", "float16", "bfloat16"])
@pytest.mark.parametrize("mnk", [(65, 256, 128), (1000, 1000, 1024), (640, 776, 2048)])
def test_prefill_gemm_tiles(dt, mnk):
    """gemm_prefill.cu (m > 64): ragged tiles in m and n, several tiles per CTA, residual in place, GLU, f16 GEMM."""
    m, n, k = mnk
    r = np.random.default_rng(m + k)
    x = round_through(r.standard_normal((m, k)), dt)
    wq, ws = O.quantize_weight((r.standard_normal((n, k)) * 0.05).astype(np.float32))
    wu, su = O.quantize_weight((r.standard_normal((n, k)) * 0.05).astype(np.float32))
    bias = round_through(r.standard_normal(n) * 0.1, dt)
    res = round_through(r.standard_normal((m, n)), dt)
    xq, xs = ops.Quantize()(dev(x, TDT[dt]))
    tol = TOL[dt] if dt != "float32" else 2e-5
    for act, use_bias, use_res in ((-1, False, True), (ops.ActivationType.GELU, True, False)):
        y = ops.dense_int8(xq, xs, dev(wq), dev(ws), dev(bias, TDT[dt]) if use_bias else None,
                           dev(res, TDT[dt]) if use_res else None, None if act < 0 else act, TDT[dt], ops.GEMM_TCGEN05)
        ref = O.dense_int8(x, wq, ws, bias if use_bias else None, act, res if use_res else None, "cuda")
        np.testing.assert_allclose(to_np(y), ref, rtol=tol, atol=tol * max(1.0, float(np.abs(ref).max())))
    if dt != "bfloat16":
        h = ops.dense_int8_glu(xq, xs, dev(wq), dev(ws), dev(wu), dev(su), ops.ActivationType.Swish, TDT[dt], ops.GEMM_TCGEN05)
        gate = round_through(O.dense_int8(x, wq, ws, None, O.ACT_SWISH), dt)
        ref = gate * round_through(O.dense_int8(x, wu, su), dt)
        np.testing.assert_allclose(to_np(h), ref, rtol=tol, atol=tol * max(1.0, float(np.abs(ref).max())))
    if dt != "float32":
        a = round_through(r.standard_normal((m, k)), dt)
        b = round_through(r.standard_normal((n, k)) * 0.05, dt)
        c = ops.Gemm()(dev(a, TDT[dt]), dev(b, TDT[dt]), bias=dev(bias, TDT[dt]))
        ref = a.astype(np.float64) @ b.astype(np.float64).T + bias
        np.testing.assert_allclose(to_np(c), ref, rtol=TOL[dt], atol=TOL[dt] * float(np.abs(ref).max()))


@gpu
def test_decode_gemm_auto_plan_llama_shapes():
    """The shapes of one Llama-3-8B decode layer through the planner's own choice (no pinning), batch 1 and 32."""
    r = np.random.default_rng(8)
    for n, k in ((6144, 4096), (4096, 4096), (4096, 14336)):
        w = (r.standard_normal((n, k)) * 0.05).astype(np.float32)
        wq, ws = O.quantize_weight(w)
        for m in (1, 32):
            x = r.standard_normal((m, k)).astype(np.float32)
            xq, xs = ops.Quantize()(dev(x, TDT["float32"]))
            y = ops.dense_int8(xq, xs, dev(wq), dev(ws), None, None, None, TDT["float32"], ops.GEMM_TCGEN05)
            ref = O.dense_int8(x, wq, ws, None, -1, None, "cuda")
            np.testing.assert_allclose(to_np(y), ref, rtol=2e-5, atol=2e-5 * float(np.abs(ref).max()))


# ---------------- RMSNorm / Rotary / SoftMax / TopK / Gather / Embeddings ----------------
@gpu
@pytest.mark.parametrize("dt", DTYPES)
def test_rms_norm(dt):
    y = ops.RMSNorm()(dev(vec("RMSNorm", "gamma"), TDT[dt]), dev(vec("RMSNorm", "x"), TDT[dt]))
    np.testing.assert_allclose(to_np(y), vec("RMSNorm", "expected"), atol=TOL[dt] * 4.2)
    r = np.random.default_rng(1)
    x, g = round_through(r.standard_normal((9, 4096)), dt), round_through(1 + 0.1 * r.standard_normal(4096), dt)
    y = ops.RMSNorm(1e-5)(dev(g, TDT[dt]), dev(x, TDT[dt]))
    ref = O.rms_norm(x, g, 1e-5)
    np.testing.assert_allclose(to_np(y), ref, rtol=TOL[dt], atol=TOL[dt])
    # fused RMSNorm+Quantize == Quantize(T(RMSNorm)) bit-exactly
    q, s = ops.RMSNorm(1e-5).quantize(dev(g, TDT[dt]), dev(x, TDT[dt]))
    q2, s2 = ops.Quantize()(y)
    assert torch.equal(q, q2) and torch.equal(s, s2)


@gpu
def test_rms_norm_reference_fixture():
    y = ops.RMSNorm(1e-5)(dev(RND["rn_gamma"]), dev(RND["q_x"]))
    np.testing.assert_allclose(to_np(y), RND["rn_y"], rtol=1e-5, atol=1e-6)


@gpu
@pytest.mark.parametrize("dt", DTYPES)
def test_rotary(dt):
    x, exp = vec("RotaryEmbedding", "input"), vec("RotaryEmbedding", "expected")
    sin, cos = O.rotary_tables(4, 6, 10000.0, interleave=True)
    y = ops.Rotary(0, True)(dev(x, TDT[dt]), dev(sin[2:4], TDT[dt]), dev(cos[2:4], TDT[dt]))
    np.testing.assert_allclose(to_np(y), exp, atol=TOL[dt] * 1.5)
    for key, inter in (("ro_y_interleave", True), ("ro_y_half", False)):
        y = ops.Rotary(0, inter)(dev(RND["ro_x"], TDT[dt]), dev(RND["ro_sin"], TDT[dt]), dev(RND["ro_cos"], TDT[dt]))
        np.testing.assert_allclose(to_np(y), RND[key], atol=TOL[dt] * 4)


@gpu
@pytest.mark.parametrize("dt", DTYPES)
def test_softmax(dt):
    np.testing.assert_allclose(to_np(ops.SoftMax()(dev(vec("SoftMax", "x"), TDT[dt]))), vec("SoftMax", "expected"), atol=TOL[dt])
    np.testing.assert_allclose(to_np(ops.LogSoftMax()(dev(vec("LogSoftMax", "x"), TDT[dt]))), vec("LogSoftMax", "expected"),
                               atol=TOL[dt] * 10)
    y = ops.SoftMax()(dev(vec("MaskedSoftMax", "x"), TDT[dt]), dev(vec("MaskedSoftMax", "lengths", np.int32)))
    np.testing.assert_allclose(to_np(y), vec("MaskedSoftMax", "expected"), atol=TOL[dt])
    if dt == "float32":
        np.testing.assert_allclose(to_np(ops.SoftMax()(dev(RND["sm_x"]), dev(RND["sm_len"]))), RND["sm_y_len"], atol=1e-6)
        np.testing.assert_allclose(to_np(ops.LogSoftMax()(dev(RND["sm_x"]))), RND["sm_logy"], atol=1e-5)


@gpu
def test_topk_golden_and_ties():
    for test in ("TopK", "TopKVariableDepth"):
        v, i = ops.TopK(3)(dev(vec(test, "input")))
        np.testing.assert_array_equal(to_np(i), vec(test, "expected_indices", np.int32))
        np.testing.assert_allclose(to_np(v), vec(test, "expected_values"))
    v, i = ops.TopK(1)(dev(RND["tk_x"]))
    np.testing.assert_array_equal(to_np(i), RND["tk_i1"])
    np.testing.assert_array_equal(to_np(v), RND["tk_v1"])
    v, i = ops.TopK(4)(dev(RND["tk_x"]))
    np.testing.assert_array_equal(to_np(v), RND["tk_v4"])
    assert to_np(i)[1].tolist()[:2] == [17, 500]          # exact tie: lowest index first (documented rule)


@gpu
@pytest.mark.parametrize("dt", DTYPES)
def test_topk_large_vocab_bit_exact(dt):
    r = np.random.default_rng(5)
    x = round_through(r.standard_normal((6, 128256)), dt)   # fp16 logits => exact ties are realistic
    x[0, 100] = x[0, 99999] = x[0].max() + 1
    for k in (1, 8):
        v, i = ops.TopK(k)(dev(x, TDT[dt]))
        vo, io = O.topk(x, k)
        np.testing.assert_array_equal(to_np(i), io)
        np.testing.assert_array_equal(to_np(v), vo)


@gpu
def test_gather_and_embeddings():
    for t in ("GatherData1D", "GatherData2D", "GatherData3D"):
        out = ops.Gather()(dev(vec(t, "data")), dev(vec(t, "ids", np.int32)))
        np.testing.assert_array_equal(to_np(out), vec(t, "expected"))
    np.testing.assert_array_equal(to_np(ops.Gather()(dev(RND["ga_d"]), dev(RND["ga_i"]))), RND["ga_y"])
    r = np.random.default_rng(2)
    w, s = O.quantize_weight(r.standard_normal((300, 4096)).astype(np.float32))
    ids = r.integers(0, 300, (4, 5)).astype(np.int32)
    for dt in DTYPES:
        y = ops.embedding_int8(dev(w), dev(s), dev(ids), TDT[dt])
        ref = w[ids].astype(np.float32) / s[ids][..., None]
        np.testing.assert_allclose(to_np(y), ref, rtol=TOL[dt], atol=1e-6)


@gpu
def test_invalid_arguments_raise():
    with pytest.raises(ValueError):
        ops.Gemm()(torch.zeros((2, 24), dtype=torch.int8, device=DEV), torch.zeros((2, 24), dtype=torch.int8, device=DEV))
    with pytest.raises(ValueError):
        ops.TopK(100)(torch.zeros((2, 8), device=DEV))
    with pytest.raises(ValueError):
        ops.Quantize()(torch.zeros((2, 8)))   # CPU tensor: no CPU path


@gpu
@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
@pytest.mark.parametrize("mnk", [(1, 512, 1024), (5, 6144, 4096), (32, 4096, 4096), (64, 4096, 14336), (33, 1000, 2048)])
def test_dense_int8_rows_fused_is_bit_identical(dtype, mnk):
    """ct2b200_dense_s8_rows / _glu_rows ([RMSNorm +] Quantize + Dense from T rows) give exactly the bits of the separate
    ops::Quantize / RMSNorm + Quantize calls followed by the fused Dense — int8 rows, scales and outputs, call after call."""
    m, n, k = mnk
    r = np.random.default_rng(m + n)
    x = dev(r.standard_normal((m, k)).astype(np.float32) * 3, TDT[dtype])
    gamma = dev(r.uniform(0.5, 1.5, size=k).astype(np.float32), TDT[dtype])
    res = dev(r.standard_normal((m, n)).astype(np.float32), TDT[dtype])
    w = dev(r.integers(-127, 128, size=(n, k)).astype(np.int8))
    ws = dev(r.uniform(500, 4000, size=n).astype(np.float32))
    w2 = dev(r.integers(-127, 128, size=(n, k)).astype(np.int8))
    for rep in range(3):
        for g in (None, gamma):
            if g is None:
                xq0, xs0 = ops.Quantize()(x)
            else:
                xq0, xs0 = ops.RMSNorm(1e-5).quantize(g, x)
            y0 = ops.dense_int8(xq0, xs0, w, ws, residual=res, dtype=TDT[dtype])
            y1, xq1, xs1 = ops.dense_int8_rows(x, w, ws, gamma=g, eps=1e-5, residual=res)
            assert torch.equal(xq0, xq1) and torch.equal(xs0, xs1)
            assert torch.equal(y0, y1)
            h0 = ops.dense_int8_glu(xq0, xs0, w, ws, w2, ws, dtype=TDT[dtype])
            h1, xq2, xs2 = ops.dense_int8_glu_rows(x, w, ws, w2, ws, gamma=g, eps=1e-5)
            assert torch.equal(xq0, xq2) and torch.equal(xs0, xs2)
            assert torch.equal(h0, h1)
