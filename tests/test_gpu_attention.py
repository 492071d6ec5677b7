"""-m gpu: the attention kernels (C-ABI) vs the oracle's restatement of MultiHeadAttention
(rotary -> cache append -> softmax(q k^T / sqrt(d)) v over an un-replicated GQA cache)."""
import math

import numpy as np
import pytest
import torch

from ctranslate2_b200 import ops
from oracle import ct2_oracle as O
from gpu_util import DEV, TDT, TOL, dev, gpu, round_through, to_np


def ref_attention(q, k, v, G):
    """q [B,H,T,D] (rotated), k/v [B,Hkv,S,D]; causal with the T queries at the END of the S keys."""
    B, H, T, D = q.shape
    S = k.shape[2]
    kr, vr = np.repeat(k, G, axis=1), np.repeat(v, G, axis=1)
    sc = np.einsum("bhtd,bhsd->bhts", q, kr) / math.sqrt(D)
    lens = np.broadcast_to((S - T) + np.arange(T) + 1, (B, H, T)).reshape(-1)
    p = O.softmax(sc.reshape(-1, S).astype(np.float32), lens).reshape(B, H, T, S)
    return np.einsum("bhts,bhsd->bhtd", p, vr)


@gpu
@pytest.mark.parametrize("dt", ["float32", "float16", "bfloat16"])
@pytest.mark.parametrize("cfg", [dict(B=1, H=32, Hkv=8, D=128, lens=[700]), dict(B=3, H=8, Hkv=2, D=128, lens=[0, 5, 130]),
                                 dict(B=2, H=4, Hkv=4, D=64, lens=[63, 64]), dict(B=4, H=4, Hkv=2, D=32, lens=[1, 2, 3, 300]),
                                 dict(B=2, H=8, Hkv=1, D=128, lens=[17, 257]),
                                 # many ragged rows: (row, head) pairs start and end in the middle of a CTA's unit range
                                 dict(B=40, H=32, Hkv=8, D=128, lens="ragged"),
                                 dict(B=5, H=16, Hkv=2, D=64, lens=[1023, 0, 511, 64, 65]),
                                 dict(B=33, H=8, Hkv=8, D=128, lens="ragged")])
@pytest.mark.parametrize("kernel", ["auto", "persistent", "split"])
def test_attention_decode(dt, cfg, kernel, monkeypatch):
    """kernel: the launcher's own choice, or CT2B200_ATTN_DECODE pinned to the persistent / the split-KV kernel."""
    if kernel != "auto":
        if dt == "float32" or cfg["D"] == 32:
            pytest.skip("fp32 / head_dim 32 always take the SIMT kernel")
        monkeypatch.setenv("CT2B200_ATTN_DECODE", kernel)
    B, H, Hkv, D, lens = cfg["B"], cfg["H"], cfg["Hkv"], cfg["D"], cfg["lens"]
    if lens == "ragged":
        lens = np.random.default_rng(B).integers(0, 1000, size=B).tolist()
    G, max_len = H // Hkv, 1024
    r = np.random.default_rng(B * 31 + H)
    qkv = round_through(r.standard_normal((B, (H + 2 * Hkv) * D)), dt)
    kc = round_through(r.standard_normal((B, Hkv, max_len, D)), dt)
    vc = round_through(r.standard_normal((B, Hkv, max_len, D)), dt)
    sin, cos = O.rotary_tables(max_len, D, 500000.0, interleave=False)
    kc_d, vc_d = dev(kc, TDT[dt]), dev(vc, TDT[dt])
    out = ops.attention_decode(dev(qkv, TDT[dt]), kc_d, vc_d, dev(sin), dev(cos), dev(np.array(lens, np.int32)), H, Hkv, D)
    out = to_np(out).reshape(B, H, D)
    tol = TOL[dt] if dt != "float32" else 2e-5
    for b in range(B):
        pos = lens[b]
        q = qkv[b, :H * D].reshape(1, H, 1, D)
        k = qkv[b, H * D:(H + Hkv) * D].reshape(1, Hkv, 1, D)
        v = qkv[b, (H + Hkv) * D:].reshape(1, Hkv, 1, D)
        qr = O.rotary(q, sin[pos:pos + 1], cos[pos:pos + 1], False)
        kr = round_through(O.rotary(k, sin[pos:pos + 1], cos[pos:pos + 1], False), dt)
        K = np.concatenate([kc[b:b + 1, :, :pos], kr], axis=2)
        V = np.concatenate([vc[b:b + 1, :, :pos], v], axis=2)
        ref = ref_attention(qr, K, V, G)[0, :, 0]
        np.testing.assert_allclose(out[b], ref, rtol=tol, atol=tol * 2)
        # the cache was appended in place at position `pos` (bit-exact copy for V, rotated K within tolerance)
        np.testing.assert_array_equal(to_np(vc_d[b, :, pos]), v[0, :, 0])
        np.testing.assert_allclose(to_np(kc_d[b, :, pos]), kr[0, :, 0], rtol=tol, atol=tol)
        # nothing else was touched
        np.testing.assert_array_equal(to_np(kc_d[b, :, :pos]), kc[b, :, :pos])


@gpu
@pytest.mark.parametrize("dt", ["float32", "float16"])
@pytest.mark.parametrize("cfg", [dict(B=2, T=9, off=0, H=8, Hkv=2, D=128), dict(B=1, T=70, off=33, H=4, Hkv=4, D=64),
                                 dict(B=3, T=5, off=2, H=4, Hkv=2, D=32)])
def test_attention_prefill(dt, cfg):
    B, T, off, H, Hkv, D = (cfg[k] for k in ("B", "T", "off", "H", "Hkv", "D"))
    G, max_len = H // Hkv, 256
    r = np.random.default_rng(T)
    qkv = round_through(r.standard_normal((B * T, (H + 2 * Hkv) * D)), dt)
    kc = round_through(r.standard_normal((B, Hkv, max_len, D)), dt)
    vc = round_through(r.standard_normal((B, Hkv, max_len, D)), dt)
    sin, cos = O.rotary_tables(max_len, D, 10000.0, interleave=False)
    kc_d, vc_d = dev(kc, TDT[dt]), dev(vc, TDT[dt])
    out = ops.attention_prefill(dev(qkv, TDT[dt]), kc_d, vc_d, dev(sin), dev(cos), B, T, off, H, Hkv, D)
    out = to_np(out).reshape(B, T, H, D).transpose(0, 2, 1, 3)
    x = qkv.reshape(B, T, -1)
    q = x[..., :H * D].reshape(B, T, H, D).transpose(0, 2, 1, 3)
    k = x[..., H * D:(H + Hkv) * D].reshape(B, T, Hkv, D).transpose(0, 2, 1, 3)
    v = x[..., (H + Hkv) * D:].reshape(B, T, Hkv, D).transpose(0, 2, 1, 3)
    qr = round_through(O.rotary(q, sin[off:off + T], cos[off:off + T], False), dt)
    kr = round_through(O.rotary(k, sin[off:off + T], cos[off:off + T], False), dt)
    K = np.concatenate([kc[:, :, :off], kr], axis=2)
    V = np.concatenate([vc[:, :, :off], v], axis=2)
    ref = ref_attention(qr, K, V, G)
    tol = TOL[dt] if dt != "float32" else 2e-5
    np.testing.assert_allclose(out, ref, rtol=tol, atol=tol * 2)
    np.testing.assert_array_equal(to_np(vc_d[:, :, off:off + T]), v)


@gpu
@pytest.mark.parametrize("dt", ["float16", "bfloat16"])
@pytest.mark.parametrize("kernel", ["persistent", "split"])
def test_attention_decode_workspace_reuse(dt, kernel, monkeypatch):
    """Successive launches on ONE workspace (as the 32 layers of a decode step do): the ready flags of the shared
    (row, head) pairs are cleared by the combining CTA, so later launches see a clean slate; results stay exact
    when the same cache is attended again at the next position."""
    monkeypatch.setenv("CT2B200_ATTN_DECODE", kernel)
    B, H, Hkv, D, max_len = 2, 8, 2, 128, 2048
    G = H // Hkv
    r = np.random.default_rng(77)
    kc = round_through(r.standard_normal((B, Hkv, max_len, D)), dt)
    vc = round_through(r.standard_normal((B, Hkv, max_len, D)), dt)
    sin, cos = O.rotary_tables(max_len, D, 500000.0, interleave=False)
    kc_d, vc_d = dev(kc, TDT[dt]), dev(vc, TDT[dt])
    nbytes = ops.attention_decode_workspace_bytes(B, H, D, max_len)
    ws = torch.zeros(nbytes, dtype=torch.uint8, device=DEV)
    lens = np.array([1500, 900], np.int32)
    for step in range(3):
        qkv = round_through(r.standard_normal((B, (H + 2 * Hkv) * D)), dt)
        out = to_np(ops.attention_decode(dev(qkv, TDT[dt]), kc_d, vc_d, dev(sin), dev(cos), dev(lens), H, Hkv, D,
                                         workspace=ws)).reshape(B, H, D)
        kc_h, vc_h = to_np(kc_d), to_np(vc_d)                # caches after the append
        for b in range(B):
            pos = int(lens[b])
            q = qkv[b, :H * D].reshape(1, H, 1, D)
            qr = O.rotary(q, sin[pos:pos + 1], cos[pos:pos + 1], False)
            ref = ref_attention(qr, kc_h[b:b + 1, :, :pos + 1], vc_h[b:b + 1, :, :pos + 1], G)[0, :, 0]
            np.testing.assert_allclose(out[b], ref, rtol=TOL[dt], atol=TOL[dt] * 2)
        lens = lens + 1
