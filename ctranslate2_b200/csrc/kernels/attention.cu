// attention.cu — layers::MultiHeadAttention between the QKV Dense and the output Dense
// (reference src/layers/attention.cc:485-602: Split -> [replicate_heads] -> Rotary -> Concat(cache) ->
// MatMul/SoftMax/MatMul), re-designed around an UN-replicated, preallocated GQA cache
// [slot, Hkv, max_len, D]: nothing is copied or tiled per step, K/V are appended in place.
// This file holds the launch policy and the SIMT kernels that cover every shape / dtype (fp32 activations, head_dim 32,
// ragged prefill); fp16 / bf16 with head_dim 64 / 128 go to the tensor-core kernels:
//   decode:  attention_mma.cu (split-KV grid, TMA-staged tiles) or attention_decode.cu (persistent, work-balanced);
//   prefill: attention_mma.cu (causal flash attention on mma.sync) after rope_append_kernel (here).
//  * attention_decode_kernel: SIMT split-KV "flash decoding": grid = (splits, Hkv, batch); a CTA serves the G = H/Hkv
//    query heads of one KV head over a slice of the keys, fp32 online softmax, ticketed combine; rotary of q and of the
//    new k, and the K/V append, happen in the same kernel.
//  * rope_append_kernel + attention_prefill_simple_kernel: T new tokens, causal.
#include <cstdlib>
#include <algorithm>

#include "../common.cuh"
#include "kernels.h"

namespace ct2b200 {

namespace {

constexpr int kDecThreads = 128;

// rotate element i of a head vector x[0..D) (fp32 math; reference src/ops/rotary_gpu.cu:27-85)
template <typename T>
__device__ __forceinline__ float rope_at(const T* x, const float* sin, const float* cos, int i, int D, bool interleave) {
  float other;
  if (interleave) other = (i & 1) ? to_f32(x[i - 1]) : -to_f32(x[i + 1]);
  else other = (i < D / 2) ? -to_f32(x[i + D / 2]) : to_f32(x[i - D / 2]);
  return to_f32(x[i]) * cos[i] + other * sin[i];
}

// partial results: per (batch, head, split): D floats of un-normalised output, then m (max, log2 domain), l (sum)
__host__ __device__ inline size_t partial_stride(int D) { return static_cast<size_t>(D) + 2; }

template <typename T, int D, int G>
__global__ void __launch_bounds__(kDecThreads)
    attention_decode_kernel(const T* __restrict__ qkv, T* __restrict__ k_cache, T* __restrict__ v_cache,
                            const float* __restrict__ sin_t, const float* __restrict__ cos_t,
                            const int32_t* __restrict__ lens, int H, int Hkv, int64_t max_len, bool interleave,
                            float scale_log2, T* __restrict__ out, float* __restrict__ partials,
                            int32_t* __restrict__ tickets) {
  constexpr int VEC = 16 / sizeof(T);          // elements per 16-byte load
  constexpr int LANES = D / VEC;               // lanes covering one key row
  constexpr int KPW = 32 / LANES;              // keys per warp-wide load
  constexpr int NW = kDecThreads / 32;
  static_assert(D % VEC == 0 && LANES <= 32 && 32 % LANES == 0, "unsupported head_dim");

  __shared__ float s_q[G][D];
  __shared__ __align__(16) T s_knew[D];
  __shared__ float s_m[NW][KPW][G], s_l[NW][KPW][G];
  __shared__ float s_acc[NW][KPW][G][D];
  __shared__ bool s_last;

  griddep_launch();
  griddep_wait();
  const int split = blockIdx.x, nsplit = gridDim.x, kvh = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pos = lens[b];                     // position of the new token; keys 0..pos are attended
  const int nkeys = pos + 1;
  const int64_t row_w = static_cast<int64_t>(H + 2 * Hkv) * D;
  const T* q_in = qkv + b * row_w + static_cast<int64_t>(kvh) * G * D;
  const T* k_in = qkv + b * row_w + static_cast<int64_t>(H) * D + static_cast<int64_t>(kvh) * D;
  const T* v_in = k_in + static_cast<int64_t>(Hkv) * D;
  T* kc = k_cache + (static_cast<int64_t>(b) * Hkv + kvh) * max_len * D;
  T* vc = v_cache + (static_cast<int64_t>(b) * Hkv + kvh) * max_len * D;
  const float* sn = sin_t + static_cast<int64_t>(pos) * D;
  const float* cs = cos_t + static_cast<int64_t>(pos) * D;

  // key slice of this CTA (multiples of KPW*NW keys so that warps stay aligned)
  int per = (nkeys + nsplit - 1) / nsplit;
  per = ((per + KPW * NW - 1) / (KPW * NW)) * (KPW * NW);
  const int s0 = split * per;
  const int s1 = min(nkeys, s0 + per);

  // rotary(q) * scale -> smem (fp32)
  for (int e = tid; e < G * D; e += kDecThreads) {
    const int h = e / D, i = e % D;
    s_q[h][i] = rope_at(q_in + h * D, sn, cs, i, D, interleave) * scale_log2;
  }
  // the CTA whose slice holds `pos` appends rotary(k_new), v_new to the cache
  const bool owner = pos >= s0 && pos < s1;
  if (owner) {
    for (int i = tid; i < D; i += kDecThreads) {
      const T kr = from_f32<T>(rope_at(k_in, sn, cs, i, D, interleave));
      kc[static_cast<int64_t>(pos) * D + i] = kr;
      vc[static_cast<int64_t>(pos) * D + i] = v_in[i];
    }
  }
  __syncthreads();

  // registers: this lane's VEC dims of each of the G query heads, and the matching accumulators
  const int sub = lane / LANES;                // which key of the warp-wide load this lane works on
  const int d0 = (lane % LANES) * VEC;
  float q[G][VEC], acc[G][VEC], m[G], l[G];
#pragma unroll
  for (int h = 0; h < G; ++h) {
    m[h] = -INFINITY;
    l[h] = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      q[h][i] = s_q[h][d0 + i];
      acc[h][i] = 0.f;
    }
  }

  constexpr int UNROLL = 4;
  for (int base = s0 + warp * KPW; base < s1; base += NW * KPW * UNROLL) {
    Vec16<T> kv[UNROLL], vv[UNROLL];
    bool ok[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {       // issue all loads first (memory-level parallelism)
      const int key = base + u * NW * KPW + sub;
      ok[u] = key < s1;
      const int64_t off = static_cast<int64_t>(ok[u] ? key : s0) * D + d0;
      kv[u] = ld16(kc + off);
      vv[u] = ld16(vc + off);
    }
    float kf[UNROLL][VEC], vf[UNROLL][VEC];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        kf[u][i] = to_f32(kv[u].v[i]);
        vf[u][i] = to_f32(vv[u].v[i]);
      }
#pragma unroll
    for (int h = 0; h < G; ++h) {
      // scores of the UNROLL keys of this lane group, then ONE rescale of the running state per block
      float sc[UNROLL], mblk = -INFINITY;
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) s = fmaf(q[h][i], kf[u][i], s);
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        sc[u] = ok[u] ? s : -INFINITY;
        mblk = fmaxf(mblk, sc[u]);
      }
      const float mn = fmaxf(m[h], mblk);
      if (mn == -INFINITY) continue;             // nothing valid yet (uniform within the lane group)
      const float corr = exp2f(m[h] - mn);       // m == -inf -> 0
      float psum = 0.f, pu[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        pu[u] = exp2f(sc[u] - mn);               // -inf -> 0
        psum += pu[u];
      }
      l[h] = l[h] * corr + psum;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float a = acc[h][i] * corr;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) a = fmaf(pu[u], vf[u][i], a);
        acc[h][i] = a;
      }
      m[h] = mn;
    }
  }

  // merge the KPW x NW partial states of this CTA through shared memory
#pragma unroll
  for (int h = 0; h < G; ++h) {
    if (lane % LANES == 0) {
      s_m[warp][sub][h] = m[h];
      s_l[warp][sub][h] = l[h];
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) s_acc[warp][sub][h][d0 + i] = acc[h][i];
  }
  __syncthreads();
  float* part = partials + ((static_cast<int64_t>(b) * H + static_cast<int64_t>(kvh) * G) * nsplit) * partial_stride(D);
  for (int e = tid; e < G * D; e += kDecThreads) {
    const int h = e / D, i = e % D;
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < NW; ++w)
#pragma unroll
      for (int k = 0; k < KPW; ++k) mm = fmaxf(mm, s_m[w][k][h]);
    float ll = 0.f, a = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w)
#pragma unroll
      for (int k = 0; k < KPW; ++k) {
        const float c = s_m[w][k][h] == -INFINITY ? 0.f : exp2f(s_m[w][k][h] - mm);
        ll += s_l[w][k][h] * c;
        a += s_acc[w][k][h][i] * c;
      }
    float* ph = part + (static_cast<int64_t>(h) * nsplit + split) * partial_stride(D);
    ph[i] = a;
    if (i == 0) {
      ph[D] = mm;
      ph[D + 1] = ll;
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(tickets + b * Hkv + kvh, 1) == nsplit - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last CTA of this (batch, kv-head): combine the splits and write the output
  for (int e = tid; e < G * D; e += kDecThreads) {
    const int h = e / D, i = e % D;
    const float* ph = part + static_cast<int64_t>(h) * nsplit * partial_stride(D);
    float mm = -INFINITY;
#pragma unroll 8
    for (int s = 0; s < nsplit; ++s) mm = fmaxf(mm, __ldcg(ph + s * partial_stride(D) + D));
    float ll = 0.f, a = 0.f;
#pragma unroll 8
    for (int s = 0; s < nsplit; ++s) {
      const float ms = __ldcg(ph + s * partial_stride(D) + D);
      const float ls = __ldcg(ph + s * partial_stride(D) + D + 1);
      const float as = __ldcg(ph + s * partial_stride(D) + i);
      const float c = ms == -INFINITY ? 0.f : exp2f(ms - mm);
      ll += ls * c;
      a += as * c;
    }
    out[static_cast<int64_t>(b) * H * D + (static_cast<int64_t>(kvh) * G + h) * D + i] = from_f32<T>(a / ll);
  }
  if (tid == 0) tickets[b * Hkv + kvh] = 0;
}

// ---- prefill: rotary + append for T new tokens, q rotated in place ----
template <typename T>
__global__ void rope_append_kernel(T* __restrict__ qkv, T* __restrict__ k_cache, T* __restrict__ v_cache,
                                   const float* __restrict__ sin_t, const float* __restrict__ cos_t,
                                   const int32_t* __restrict__ lengths, int64_t time, int64_t offset, int H,
                                   int Hkv, int D, int64_t max_len, bool interleave) {
  extern __shared__ float s_rot[];
  const int head = blockIdx.x;                 // 0..H-1 q heads, H..H+Hkv-1 k heads, then v heads
  const int64_t t = blockIdx.y, b = blockIdx.z;
  if (lengths && t >= lengths[b]) return;
  const int64_t row_w = static_cast<int64_t>(H + 2 * Hkv) * D;
  T* x = qkv + (b * time + t) * row_w + static_cast<int64_t>(head) * D;
  const int64_t pos = offset + t;
  if (head < H + Hkv) {
    const float* sn = sin_t + pos * D;
    const float* cs = cos_t + pos * D;
    for (int i = threadIdx.x; i < D; i += blockDim.x) s_rot[i] = rope_at(x, sn, cs, i, D, interleave);
    __syncthreads();
    if (head < H) {
      for (int i = threadIdx.x; i < D; i += blockDim.x) x[i] = from_f32<T>(s_rot[i]);
    } else {
      T* kc = k_cache + ((b * Hkv + (head - H)) * max_len + pos) * D;
      for (int i = threadIdx.x; i < D; i += blockDim.x) kc[i] = from_f32<T>(s_rot[i]);
    }
  } else {
    T* vc = v_cache + ((b * Hkv + (head - H - Hkv)) * max_len + pos) * D;
    for (int i = threadIdx.x; i < D; i += blockDim.x) vc[i] = x[i];
  }
}

// Same operation, one CTA per TOKEN ROW with 16-byte accesses (the per-head grid above launches H+2Hkv tiny blocks per
// token: 393 k blocks of 64 threads for a 32 x 256-token chunk of Llama-3-8B, i.e. scheduling-bound).  A task is one
// 16-byte vector of a q/k head together with its rotation partner (non-interleaved: the vector half a head away;
// interleaved: the pairs sit inside the vector), or one vector of a v head (plain copy).  Same fp32 math and the same
// single rounding as rope_at, so the results are bit-identical.
template <typename T>
__global__ void __launch_bounds__(256)
    rope_append_rows_kernel(T* __restrict__ qkv, T* __restrict__ k_cache, T* __restrict__ v_cache,
                            const float* __restrict__ sin_t, const float* __restrict__ cos_t,
                            const int32_t* __restrict__ lengths, int64_t time, int64_t offset, int H, int Hkv, int D,
                            int64_t max_len, bool interleave) {
  constexpr int N = Vec16<T>::N;
  const int64_t t = blockIdx.x, b = blockIdx.y;
  if (lengths && t >= lengths[b]) return;
  const int64_t row_w = static_cast<int64_t>(H + 2 * Hkv) * D;
  T* row = qkv + (b * time + t) * row_w;
  const int64_t pos = offset + t;
  const float* sn = sin_t + pos * D;
  const float* cs = cos_t + pos * D;
  const int half = D / 2;
  const int vec_per_task = interleave ? 1 : 2;
  const int tasks_per_head = D / (N * vec_per_task);
  const int rot_tasks = (H + Hkv) * tasks_per_head;
  const int copy_tasks = Hkv * (D / N);
  for (int task = threadIdx.x; task < rot_tasks + copy_tasks; task += blockDim.x) {
    if (task >= rot_tasks) {                         // v head: copy into the cache
      const int c = task - rot_tasks;
      const int hv = c / (D / N), i0 = (c % (D / N)) * N;
      const Vec16<T> v = ld16(row + static_cast<int64_t>(H + Hkv + hv) * D + i0);
      st16(v_cache + ((b * Hkv + hv) * max_len + pos) * D + i0, v);
      continue;
    }
    const int head = task / tasks_per_head, i0 = (task % tasks_per_head) * N;
    T* x = row + static_cast<int64_t>(head) * D;
    T* dst = head < H ? x : k_cache + ((b * Hkv + (head - H)) * max_len + pos) * D;
    if (interleave) {
      const Vec16<T> v = ld16(x + i0);
      Vec16<T> o;
#pragma unroll
      for (int j = 0; j < N; j += 2) {
        const float a = to_f32(v.v[j]), c = to_f32(v.v[j + 1]);
        o.v[j] = from_f32<T>(a * cs[i0 + j] + (-c) * sn[i0 + j]);
        o.v[j + 1] = from_f32<T>(c * cs[i0 + j + 1] + a * sn[i0 + j + 1]);
      }
      st16(dst + i0, o);
    } else {
      const Vec16<T> lo = ld16(x + i0), hi = ld16(x + i0 + half);
      Vec16<T> olo, ohi;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const float a = to_f32(lo.v[j]), c = to_f32(hi.v[j]);
        olo.v[j] = from_f32<T>(a * cs[i0 + j] + (-c) * sn[i0 + j]);
        ohi.v[j] = from_f32<T>(c * cs[i0 + half + j] + a * sn[i0 + half + j]);
      }
      st16(dst + i0, olo);
      st16(dst + i0 + half, ohi);
    }
  }
}

// ---- prefill attention, generic SIMT version: one warp per (batch, head, query) row ----
// (kept as the dtype-generic reference path; the tensor-core kernel in attention_mma.cu is the fast one)
template <typename T, int D>
__global__ void __launch_bounds__(128)
    attention_prefill_simple_kernel(const T* __restrict__ qkv, const T* __restrict__ k_cache,
                                    const T* __restrict__ v_cache, const int32_t* __restrict__ lengths,
                                    int64_t time, int64_t offset, int H, int Hkv, int64_t max_len,
                                    float scale_log2, T* __restrict__ out) {
  constexpr int VEC = 16 / sizeof(T), LANES = D / VEC, KPW = 32 / LANES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * 4 + warp;
  const int h = blockIdx.y;
  const int64_t b = blockIdx.z;
  if (t >= time) return;
  const int G = H / Hkv, kvh = h / G;
  const int64_t row_w = static_cast<int64_t>(H + 2 * Hkv) * D;
  T* o = out + ((b * time + t) * H + h) * D;
  const int sub = lane / LANES, d0 = (lane % LANES) * VEC;
  if (lengths && t >= lengths[b]) {            // padding row: defined output (zeros)
    if (sub == 0) for (int i = 0; i < VEC; ++i) o[d0 + i] = from_f32<T>(0.f);
    return;
  }
  const T* qp = qkv + (b * time + t) * row_w + static_cast<int64_t>(h) * D;
  const T* kc = k_cache + (b * Hkv + kvh) * max_len * D;
  const T* vc = v_cache + (b * Hkv + kvh) * max_len * D;
  const int nkeys = static_cast<int>(offset + t + 1);      // causal (reference attention_layer.cc:152-174)
  float q[VEC], acc[VEC], m = -INFINITY, l = 0.f;
  {
    const Vec16<T> qv = ld16(qp + d0);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      q[i] = to_f32(qv.v[i]) * scale_log2;
      acc[i] = 0.f;
    }
  }
  for (int base = 0; base < nkeys; base += KPW) {
    const int key = base + sub;
    const bool ok = key < nkeys;
    const int64_t off = static_cast<int64_t>(ok ? key : 0) * D + d0;
    const Vec16<T> kv = ld16(kc + off), vv = ld16(vc + off);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) s = fmaf(q[i], to_f32(kv.v[i]), s);
#pragma unroll
    for (int o2 = LANES / 2; o2 > 0; o2 >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2);
    if (ok) {
      const float mn = fmaxf(m, s), corr = exp2f(m - mn), p = exp2f(s - mn);
      l = l * corr + p;
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = fmaf(p, to_f32(vv.v[i]), acc[i] * corr);
      m = mn;
    }
  }
  // merge the KPW sub-states of the warp
#pragma unroll
  for (int o2 = LANES; o2 < 32; o2 <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o2), l2 = __shfl_xor_sync(0xffffffffu, l, o2);
    const float mn = fmaxf(m, m2);
    const float c1 = m == -INFINITY ? 0.f : exp2f(m - mn), c2 = m2 == -INFINITY ? 0.f : exp2f(m2 - mn);
    l = l * c1 + l2 * c2;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float a2 = __shfl_xor_sync(0xffffffffu, acc[i], o2);
      acc[i] = acc[i] * c1 + a2 * c2;
    }
    m = mn;
  }
  if (sub == 0) {
    Vec16<T> r;
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.v[i] = from_f32<T>(acc[i] / l);
    st16(o + d0, r);
  }
}

template <typename T, int D>
void launch_decode_g(const void* qkv, void* kc, void* vc, const float* sn, const float* cs, const int32_t* lens,
                     int64_t batch, int H, int Hkv, int64_t max_len, bool interleave, float scale, void* out,
                     float* partials, int32_t* tickets, int splits, cudaStream_t st) {
  const float scale_log2 = scale * 1.4426950408889634f;
  dim3 grid(splits, Hkv, static_cast<unsigned>(batch));
  const int G = H / Hkv;
#define CT2_LAUNCH_G(GV)                                                                                      \
  launch_pdl(attention_decode_kernel<T, D, GV>, grid, dim3(kDecThreads), 0, st, static_cast<const T*>(qkv),   \
             static_cast<T*>(kc), static_cast<T*>(vc), sn, cs, lens, H, Hkv, max_len, interleave, scale_log2, \
             static_cast<T*>(out), partials, tickets)
  switch (G) {
    case 1: CT2_LAUNCH_G(1); break;
    case 2: CT2_LAUNCH_G(2); break;
    case 4: CT2_LAUNCH_G(4); break;
    case 8: CT2_LAUNCH_G(8); break;
    default: throw InvalidArgument("attention_decode: num_heads / num_heads_kv must be 1, 2, 4 or 8");
  }
#undef CT2_LAUNCH_G
  check_launch();
}

}  // namespace

int attention_decode_splits(int64_t batch, int Hkv, int64_t max_len, int sm_count) {
  // Each (batch row, kv head) is cut into `s` slices of whole 64-key tiles.  Cost model: the kernel is bound by the
  // per-SM streaming rate, so a candidate s costs  ceil(CTAs / SMs) * (tiles per slice + fixed per-CTA overhead);
  // the smallest cost wins (ties: fewer slices, which also skips the partial/ticket/combine path when s == 1).
  if (const char* e = std::getenv("CT2B200_ATTN_SPLITS")) {
    const int s = std::atoi(e);
    if (s >= 1 && s <= 16) return s;
  }
  const int64_t ctas = batch * Hkv;
  const int64_t tiles = std::max<int64_t>(1, (max_len + 63) / 64);
  int best = 1;
  double best_cost = 1e30;
  for (int s = 1; s <= 16 && s <= tiles; ++s) {
    const int64_t per = (tiles + s - 1) / s;
    const int64_t rounds = (ctas * s + sm_count - 1) / sm_count;
    const double cost = static_cast<double>(rounds) * (static_cast<double>(per) + 1.5);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  return best;
}

size_t attention_decode_workspace_bytes(int64_t batch, int H, int D, int splits) {
  // tickets (int32 per (batch, kv-head) <= batch*H) first, then the fp32 partials
  const size_t tickets = ((static_cast<size_t>(batch) * H * sizeof(int32_t) + 255) / 256) * 256;
  return tickets + static_cast<size_t>(batch) * H * splits * partial_stride(D) * sizeof(float);
}

void launch_attention_decode(const void* qkv, void* kc, void* vc, const float* sn, const float* cs,
                             const int32_t* lens, int64_t batch, int H, int Hkv, int D, int64_t max_len,
                             bool interleave, float scale, void* out, void* workspace, size_t workspace_bytes,
                             int splits, int dtype, cudaStream_t st) {
  if (batch == 0) return;
  CT2_REQUIRE(H % Hkv == 0, "attention: num_heads must be a multiple of num_heads_kv");
  CT2_REQUIRE(workspace_bytes >= attention_decode_workspace_bytes(batch, H, D, splits),
              "attention_decode: workspace too small");
  int32_t* tickets = static_cast<int32_t*>(workspace);
  const size_t toff = ((static_cast<size_t>(batch) * H * sizeof(int32_t) + 255) / 256) * 256;
  float* partials = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + toff);
  // persistent work-balanced kernel (attention_decode.cu) when the workspace holds its 64 slots per (row, head) behind
  // the 16 of the split-KV kernel (which therefore never runs with more than 16 slices)
  const int slots = static_cast<int>((workspace_bytes - toff) / (static_cast<size_t>(batch) * H * partial_stride(D) * sizeof(float)));
  if (launch_attention_decode_persistent(qkv, kc, vc, sn, cs, lens, batch, H, Hkv, D, max_len, interleave, scale, out,
                                         partials, tickets, slots, dtype, st))
    return;
  if (launch_attention_decode_mma(qkv, kc, vc, sn, cs, lens, batch, H, Hkv, D, max_len, interleave, scale, out, partials,
                                  tickets, splits, dtype, st))
    return;
#define CT2_DEC(DV)                                                                                          \
  CT2_DISPATCH_DTYPE(dtype, (launch_decode_g<T, DV>(qkv, kc, vc, sn, cs, lens, batch, H, Hkv, max_len,       \
                                                     interleave, scale, out, partials, tickets, splits, st)))
  switch (D) {
    case 128: CT2_DEC(128); break;
    case 64: CT2_DEC(64); break;
    case 32: CT2_DEC(32); break;
    default: throw InvalidArgument("attention: head_dim must be 32, 64 or 128");
  }
#undef CT2_DEC
}

void launch_rope_append(void* qkv, void* kc, void* vc, const float* sn, const float* cs, const int32_t* lengths,
                        int64_t batch, int64_t time, int64_t offset, int H, int Hkv, int D, int64_t max_len,
                        bool interleave, int dtype, cudaStream_t st) {
  if (batch * time == 0) return;
  const int vecn = dtype == CT2B200_F32 ? 4 : 8;
  const bool aligned = (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(kc) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(vc) & 15) == 0;
  if (aligned && D % (2 * vecn) == 0 && time <= 2147483647 && batch <= 65535) {
    dim3 grid_rows(static_cast<unsigned>(time), static_cast<unsigned>(batch));
    CT2_DISPATCH_DTYPE(dtype, (rope_append_rows_kernel<T><<<grid_rows, 256, 0, st>>>(
                                  static_cast<T*>(qkv), static_cast<T*>(kc), static_cast<T*>(vc), sn, cs, lengths,
                                  time, offset, H, Hkv, D, max_len, interleave)));
    check_launch();
    return;
  }
  dim3 grid(H + 2 * Hkv, static_cast<unsigned>(time), static_cast<unsigned>(batch));
  CT2_DISPATCH_DTYPE(dtype, (rope_append_kernel<T><<<grid, 64, D * sizeof(float), st>>>(
                                static_cast<T*>(qkv), static_cast<T*>(kc), static_cast<T*>(vc), sn, cs, lengths,
                                time, offset, H, Hkv, D, max_len, interleave)));
  check_launch();
}

void launch_attention_prefill_simple(const void* qkv, const void* kc, const void* vc, const int32_t* lengths,
                                     int64_t batch, int64_t time, int64_t offset, int H, int Hkv, int D,
                                     int64_t max_len, float scale, void* out, int dtype, cudaStream_t st) {
  if (batch * time == 0) return;
  const float scale_log2 = scale * 1.4426950408889634f;
  dim3 grid(div_up(time, 4), H, static_cast<unsigned>(batch));
#define CT2_PRE(DV)                                                                                           \
  CT2_DISPATCH_DTYPE(dtype, (attention_prefill_simple_kernel<T, DV><<<grid, 128, 0, st>>>(                    \
                                static_cast<const T*>(qkv), static_cast<const T*>(kc), static_cast<const T*>(vc), \
                                lengths, time, offset, H, Hkv, max_len, scale_log2, static_cast<T*>(out))))
  switch (D) {
    case 128: CT2_PRE(128); break;
    case 64: CT2_PRE(64); break;
    case 32: CT2_PRE(32); break;
    default: throw InvalidArgument("attention: head_dim must be 32, 64 or 128");
  }
#undef CT2_PRE
  check_launch();
}

}  // namespace ct2b200
