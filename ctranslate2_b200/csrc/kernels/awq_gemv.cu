// awq_gemv.cu — AWQ-INT4 Dense for one or two rows on the CUDA cores: the latency path of the decode step (batch 1).
//
// Replaces ops::GemvAwq (src/ops/awq/gemv_gpu.cu:289-470: one warp per output channel, fp32 FMAs, a second launch for the
// split-K sum at m > 8) + bias / activation / Mul.  At m = 1, 2 the tensor cores have nothing to amortise: the work is
// streaming 0.5 byte per weight and turning it into fp16 once.  One warp owns one output channel; per trip a lane loads 16
// bytes of packed nibbles (32 channels, one quantization group), turns them into the exact integers q - z as fp16 with the
// lop3 magic-number trick of awq_common.cuh and multiplies them with the activation rows in packed half2 math; 16 products
// accumulate in half2, then the partial sum is scaled by the group's scale and folded into fp32 accumulators
// (sum_k s (q - z) x = s sum_k (q - z) x inside a group: one multiply per 16 channels instead of one per channel; the
// reference's gemv applies s per weight in fp32, gemv_gpu.cu:331-352 — same value up to fp16 summation rounding).
// ~1.5 instructions per weight at m = 1, so the kernel is bound by the HBM stream, not by the conversion (the tensor-core kernels of awq.cu / awq_decode.cu pay a shared- or tensor-memory round trip per tile and a
// pipeline hand-over per K block instead).  No shared memory, no barriers; the activations come through L1.
//
// Native layout (ct2b200_awq_repack): wp int32 [n, k/8] (word w of row c = channels 8w .. 8w+7 in nibbles {0,4,1,5,2,6,3,7}),
// sc / zr fp16 [n, k/group].  group = 128 (what AutoAWQ writes and the reference's kernels assume), k % 128 == 0, k <= 16384;
// other shapes and m > 2 go to the tensor-core kernels (awq_decode.cu, awq.cu).
#include "awq_common.cuh"
#include "gemm_decode_common.cuh"
#include "kernels.h"

namespace ct2b200 {
namespace {

constexpr int kWarps = 8;

struct GemvWeight {
  const uint32_t* wp;
  const __half* sc;
  const __half* zr;
};

struct GemvParams {
  int64_t n, k;
  int group;
  // plain Dense epilogue (NB == 1) / SwiGLU (NB == 2)
  const __half* bias;
  const __half* residual;
  __half* y;
  int act;
};

// One warp = one output channel.  A "trip" is 1024 input channels (a lane owns 32 of them = one 16-byte load of nibbles);
// trips are processed four at a time with all their weight loads issued first, so a lane keeps 64 (NB = 2: 128) bytes of the
// HBM stream in flight — with 24+ resident warps per SM that covers the bandwidth-delay product (~40 KB per SM).  The
// group scales / zeros of the row (k / 128 of each) are fetched once with coalesced loads — lane g holds group g + 32 c
// of chunk c — and handed to the lane that needs them by shuffle: trip u of chunk c uses group 32 c + 8 u + lane / 4.
constexpr int kChunkTrips = 4;
constexpr int kMaxChunks = 4;          // k <= 16384

template <int M, int NB>
__global__ void __launch_bounds__(kWarps * 32) awq_gemv_kernel(const __half* __restrict__ x, GemvWeight w0, GemvWeight w1,
                                                               GemvParams p) {
  griddep_launch();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ch = static_cast<int64_t>(blockIdx.x) * kWarps + warp;
  if (ch >= p.n) {
    griddep_wait();
    return;
  }
  const int64_t words = p.k / 8;
  const int ng = static_cast<int>(p.k / 128);
  const int chunks = static_cast<int>((p.k + 4095) / 4096);

  // the weights, scales and zeros never depend on the previous kernel: requested before the dependency wait
  __half sc_r[NB][kMaxChunks], zr_r[NB][kMaxChunks];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const GemvWeight& w = b == 0 ? w0 : w1;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int g = c * 32 + lane;
      sc_r[b][c] = g < ng ? w.sc[ch * ng + g] : __float2half(0.f);
      zr_r[b][c] = g < ng ? w.zr[ch * ng + g] : __float2half(0.f);
    }
  }
  float acc[NB][M];
#pragma unroll
  for (int b = 0; b < NB; ++b)
#pragma unroll
    for (int r = 0; r < M; ++r) acc[b][r] = 0.f;

  uint4 q[NB][kChunkTrips];
  auto load_chunk = [&](int c) {
#pragma unroll
    for (int u = 0; u < kChunkTrips; ++u) {
      const int64_t k0 = (static_cast<int64_t>(c) * kChunkTrips + u) * 1024 + lane * 32;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const GemvWeight& w = b == 0 ? w0 : w1;
        q[b][u] = k0 < p.k ? __ldcs(reinterpret_cast<const uint4*>(w.wp + ch * words + k0 / 8)) : make_uint4(0, 0, 0, 0);
      }
    }
  };
  load_chunk(0);
  griddep_wait();                                  // the activations come from the previous kernel

#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    if (c >= chunks) break;
#pragma unroll
    for (int u = 0; u < kChunkTrips; ++u) {
      const int64_t k0 = (static_cast<int64_t>(c) * kChunkTrips + u) * 1024 + lane * 32;
      const bool live = k0 < p.k;                                        // warp-uniform only for whole trips: predicate per lane
      __half2 zb[NB], zt[NB];
      float sc_f[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const int src = u * 8 + (lane >> 2);
        const __half sc = __shfl_sync(0xffffffffu, sc_r[b][c], src), zp = __shfl_sync(0xffffffffu, zr_r[b][c], src);
        zb[b] = __half2half2(__hadd(__float2half(1024.f), zp));
        zt[b] = __half2half2(__hneg(__hadd(__float2half(64.f), zp)));
        sc_f[b] = live ? __half2float(sc) : 0.f;
      }
      uint4 xv[M][4];
#pragma unroll
      for (int r = 0; r < M; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          xv[r][j] = live ? *(reinterpret_cast<const uint4*>(x + r * p.k + k0) + j) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const uint32_t wq[4] = {q[b][u].x, q[b][u].y, q[b][u].z, q[b][u].w};
#pragma unroll
        for (int half_trip = 0; half_trip < 2; ++half_trip) {      // 16 channels -> one half2 partial sum per row
          __half2 part[M];
#pragma unroll
          for (int r = 0; r < M; ++r) part[r] = __float2half2_rn(0.f);
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int j = half_trip * 2 + jj;
            const uint4 d = awq_unscaled_word(wq[j], zb[b], zt[b]);
            const __half2* dv = reinterpret_cast<const __half2*>(&d);
#pragma unroll
            for (int r = 0; r < M; ++r) {
              const __half2* xh = reinterpret_cast<const __half2*>(&xv[r][j]);
#pragma unroll
              for (int i = 0; i < 4; ++i) part[r] = __hfma2(dv[i], xh[i], part[r]);
            }
          }
#pragma unroll
          for (int r = 0; r < M; ++r) {
            const float2 f = __half22float2(part[r]);
            acc[b][r] = fmaf(f.x + f.y, sc_f[b], acc[b][r]);
          }
        }
      }
    }
    if (c + 1 < chunks) load_chunk(c + 1);
  }
#pragma unroll
  for (int b = 0; b < NB; ++b)
#pragma unroll
    for (int r = 0; r < M; ++r) acc[b][r] = warp_sum(acc[b][r]);
  if (lane == 0) {
    if constexpr (NB == 1) {
      FloatEpilogue e{p.bias, p.residual, p.y, p.act, p.n};
#pragma unroll
      for (int r = 0; r < M; ++r) float_epilogue_store<__half>(e, acc[0][r], r, ch);
    } else {
      FloatGluEpilogue e{p.y, p.act, p.n};
#pragma unroll
      for (int r = 0; r < M; ++r) float_glu_epilogue_store<__half>(e, acc[0][r], acc[1][r], r, ch);
    }
  }
}

template <int NB>
void launch_gemv(const void* x, const AwqNative& a, const AwqNative* b, int64_t m, const GemvParams& p, cudaStream_t st) {
  const GemvWeight w0{static_cast<const uint32_t*>(a.wp), static_cast<const __half*>(a.sc), static_cast<const __half*>(a.zr)};
  const GemvWeight w1 = b ? GemvWeight{static_cast<const uint32_t*>(b->wp), static_cast<const __half*>(b->sc),
                                       static_cast<const __half*>(b->zr)} : w0;
  const dim3 grid(static_cast<unsigned>((a.n + kWarps - 1) / kWarps)), block(kWarps * 32);
  const __half* xh = static_cast<const __half*>(x);
  if (m == 1) launch_pdl(awq_gemv_kernel<1, NB>, grid, block, 0, st, xh, w0, w1, p);
  else launch_pdl(awq_gemv_kernel<2, NB>, grid, block, 0, st, xh, w0, w1, p);
  check_launch();
}

bool covered(const AwqNative& w, int64_t m, const void* x) {
  return m >= 1 && m <= 2 && w.group == 128 && w.k % 128 == 0 && w.k <= 4096 * kMaxChunks &&
         (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w.wp) & 15) == 0;
}

// CT2B200_AWQ_GEMV: 1 = use this kernel for m <= 2 (opt-in until its hardware validation is recorded in profiles/README.md)
bool enabled() { return dec::env_int("CT2B200_AWQ_GEMV", CT2B200_DEFAULT_AWQ_GEMV) != 0; }

}  // namespace

bool dense_awq_gemv(const void* x, const AwqNative& w, const void* bias, const void* residual, int act, int64_t m, void* y,
                    cudaStream_t st) {
  if (!enabled() || !covered(w, m, x)) return false;
  GemvParams p{w.n, w.k, w.group, static_cast<const __half*>(bias), static_cast<const __half*>(residual), static_cast<__half*>(y), act};
  launch_gemv<1>(x, w, nullptr, m, p, st);
  return true;
}

bool dense_awq_glu_gemv(const void* x, const AwqNative& wg, const AwqNative& wu, int act, int64_t m, void* h, cudaStream_t st) {
  if (!enabled() || !covered(wg, m, x) || !covered(wu, m, x) || wg.n != wu.n || wg.k != wu.k || wg.group != wu.group) return false;
  GemvParams p{wg.n, wg.k, wg.group, nullptr, nullptr, static_cast<__half*>(h), act};
  launch_gemv<2>(x, wg, &wu, m, p, st);
  return true;
}

}  // namespace ct2b200
