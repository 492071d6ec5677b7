// gemm_tc.cu — sm_100a tensor-core GEMMs: tcgen05.mma (kind::i8 / kind::f16) with TMA-staged,
// 128B-swizzled operand tiles in shared memory, accumulators in TMEM, and the fused Dense epilogue
// read back with tcgen05.ld.  One kernel template serves
//   * decode (m <= 64, weight-streaming, HBM-bound): "swap-AB" — the weight tile [128 x K] is the
//     UMMA M-side operand and the quantized activations [m x K] the N-side operand, so a 128-row UMMA
//     is fully used by 128 output channels and the tiny batch rides in the N dimension; split-K across
//     CTAs keeps all 148 SMs streaming weights;
//   * prefill (m large, tensor-bound): activations on the M side, weights on the N side (BN up to 256).
// Optional second weight matrix = SwiGLU gate/up fusion (two accumulators, one pass over x).
//
// Replaces cublasGemmEx s8/f16/bf16 (reference src/cuda/primitives.cu:485-597) + Dequantize epilogue
// (src/ops/dequantize_gpu.cu:30-121) + ops::Add/ops::Mul (src/layers/common.cc:392-401, transformer.cc:31-37).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one elected lane),
// warps 2..5 = epilogue (warp w reads TMEM lanes 32*(w%4) ..+31).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>

#include "../common.cuh"
#include "gemm_common.cuh"

namespace ct2b200 {

namespace {

constexpr int kTcThreads = 192;
constexpr int kTileM = 128;              // UMMA M
constexpr int kSwizzleBytes = 128;       // bytes of K per smem row (= one 128B swizzle atom)
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

// ---- PTX wrappers ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if constexpr (KIND == 0) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate));
  } else {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate));
  }
}
// 32 lanes x 32 columns of 32-bit accumulators -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows of 128 bytes,
// 8-row groups 1024 bytes apart (SBO), version 1 (sm_100), layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);        // start address, bits [0,14)
  d |= static_cast<uint64_t>(1) << 16;                           // leading byte offset (unused for SW128 K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                   // stride byte offset, bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                           // descriptor version
  d |= static_cast<uint64_t>(2) << 61;                           // SWIZZLE_128B
  return d;
}

// cute::UMMA::InstrDescriptor
template <int KIND>
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  uint32_t d = 0;
  d |= (KIND == 0 ? 2u : 1u) << 4;                  // c_format: S32 / F32
  const uint32_t fmt = KIND == 0 ? 1u /*S8*/ : (KIND == 1 ? 0u /*F16*/ : 1u /*BF16*/);
  d |= fmt << 7;                                    // a_format
  d |= fmt << 10;                                   // b_format
  d |= static_cast<uint32_t>(n >> 3) << 17;         // n_dim
  d |= static_cast<uint32_t>(kTileM >> 4) << 24;    // m_dim
  return d;                                         // a_major = b_major = K (0), dense, no negate
}

template <int KIND> struct KindTraits;
template <> struct KindTraits<0> { using Acc = int32_t; static constexpr int kElem = 1; };
template <> struct KindTraits<1> { using Acc = float; static constexpr int kElem = 2; };
template <> struct KindTraits<2> { using Acc = float; static constexpr int kElem = 2; };

// Float epilogue of the f16/bf16 GEMM (ops::Gemm::apply_bias_and_activation, reference src/ops/gemm.cc:10-25)
struct FloatEpilogue {
  const void* bias;
  const void* residual;
  void* y;
  int act;
  int64_t ldy;
};
template <typename T>
__device__ __forceinline__ void float_epilogue_store(const FloatEpilogue& e, float acc, int64_t i, int64_t j) {
  float v = round_to<T>(acc);
  if (e.bias) v = round_to<T>(v + to_f32(static_cast<const T*>(e.bias)[j]));
  if (e.act >= 0) v = round_to<T>(apply_act(v, e.act));
  if (e.residual) v = v + to_f32(static_cast<const T*>(e.residual)[i * e.ldy + j]);
  static_cast<T*>(e.y)[i * e.ldy + j] = from_f32<T>(v);
}

struct TcParams {
  int64_t rows_a;      // rows of the M-side operand (n when swapped, m otherwise)
  int64_t rows_b;      // rows of the N-side operand
  int64_t k;           // elements
  int tiles_a;         // M-side tiles of 128 rows
  int tiles_b;         // N-side tiles of BN rows
  int kb_total;        // K blocks (128 bytes of K each) per output tile
  int whole_tiles;     // 1 = CTA ranges are aligned to whole tiles (no scratch needed)
  DenseEpilogue dense;
  GluEpilogue glu;
  FloatEpilogue fl;
  int32_t* ws;
  int32_t* counters;
};

template <int BN, int NB, bool kSwap>
struct TcSmem {
  static constexpr int kA = kTileM * kSwizzleBytes * (kSwap ? NB : 1);     // M-side bytes per stage
  static constexpr int kB = BN * kSwizzleBytes * (kSwap ? 1 : NB);         // N-side bytes per stage
  static constexpr int kStage = kA + kB;
  static constexpr int kStages = (200 * 1024 / kStage) > 8 ? 8 : (200 * 1024 / kStage);
  static constexpr size_t kBytes = static_cast<size_t>(kStages) * kStage + 1024 /*align*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// index of the CTA whose unit range [c*U/P, (c+1)*U/P) contains unit u
__device__ __forceinline__ int cta_of_unit(int64_t u, int64_t U, int64_t P) {
  return static_cast<int>(((u + 1) * P + U - 1) / U - 1);
}

// Persistent "stream-K" GEMM: the work is the list of (output tile, K block) units, tile-major; CTA c of P
// owns the contiguous unit range [c*U/P, (c+1)*U/P), so every SM streams the same number of bytes and the TMA
// ring never drains between tiles.  A tile whose K range is covered by one CTA is finished by that CTA
// straight from TMEM; a tile shared by several CTAs is reduced through the zeroed scratch (integer
// red.global.add => bit-exact, order independent) and finished by the last arriver (ticket).
// Accumulators are double-buffered in TMEM so the epilogue of one segment overlaps the MMAs of the next.
//
// T = output dtype, KIND = 0 s8 / 1 f16 / 2 bf16, BN = UMMA N, NB = weight matrices (2 = GLU),
// kSwap = weights on the M side (decode).
template <typename T, int KIND, int BN, int NB, bool kSwap>
__global__ void __launch_bounds__(kTcThreads, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                   const __grid_constant__ CUtensorMap tm_w2, const TcParams p) {
  using S = TcSmem<BN, NB, kSwap>;
  constexpr int kElem = KindTraits<KIND>::kElem;
  constexpr int BK = kSwizzleBytes / kElem;            // elements of K per stage
  constexpr int kUmmaK = 32 / kElem;                   // elements of K per MMA
  constexpr int kStages = S::kStages;
  constexpr int kAccCols = BN * NB;                    // TMEM columns per accumulator buffer
  constexpr uint32_t kTmemCols = (2 * kAccCols) <= 32 ? 32 : (2 * kAccCols) <= 64 ? 64 : (2 * kAccCols) <= 128 ? 128
                               : (2 * kAccCols) <= 256 ? 256 : 512;
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256 && 2 * kAccCols <= 512, "invalid UMMA N");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * S::kStage);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;       // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  __shared__ int s_last;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t KB = p.kb_total;
  const int64_t U = static_cast<int64_t>(p.tiles_a) * p.tiles_b * KB;
  const int64_t P = gridDim.x;
  const int64_t T_all = static_cast<int64_t>(p.tiles_a) * p.tiles_b;
  const int64_t u_begin = p.whole_tiles ? (blockIdx.x * T_all / P) * KB : blockIdx.x * U / P;
  const int64_t u_end = p.whole_tiles ? ((blockIdx.x + 1) * T_all / P) * KB : (blockIdx.x + 1) * U / P;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tmem_full_bar + b, 1);
      mbar_init(tmem_empty_bar + b, 4);               // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const CUtensorMap* map_a0 = kSwap ? &tm_w : &tm_x;
  const CUtensorMap* map_a1 = &tm_w2;                   // only when kSwap && NB == 2
  const CUtensorMap* map_b0 = kSwap ? &tm_x : &tm_w;
  const CUtensorMap* map_b1 = &tm_w2;                   // only when !kSwap && NB == 2
  const uint64_t pol_a = kSwap ? kEvictFirst : kEvictLast;   // weights stream once; activations are reused
  const uint64_t pol_b = kSwap ? kEvictLast : kEvictFirst;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int it = 0;
      int tile = static_cast<int>(u_begin / KB);
      int kb = static_cast<int>(u_begin - tile * KB);
      int a0 = (tile % p.tiles_a) * kTileM, b0 = (tile / p.tiles_a) * BN;
      for (int64_t u = u_begin; u < u_end; ++u, ++it, ++kb) {
        if (kb == KB) {
          kb = 0;
          ++tile;
          a0 = (tile % p.tiles_a) * kTileM;
          b0 = (tile / p.tiles_a) * BN;
        }
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(empty_bar + s, ph ^ 1);
        mbar_expect_tx(full_bar + s, S::kStage);
        uint8_t* sa = smem + s * S::kStage;
        uint8_t* sb = sa + S::kA;
        const int kc = kb * BK;
        tma_load_2d(sa, map_a0, full_bar + s, kc, a0, pol_a);
        if (kSwap && NB == 2) tma_load_2d(sa + kTileM * kSwizzleBytes, map_a1, full_bar + s, kc, a0, pol_a);
        tma_load_2d(sb, map_b0, full_bar + s, kc, b0, pol_b);
        if (!kSwap && NB == 2) tma_load_2d(sb + BN * kSwizzleBytes, map_b1, full_bar + s, kc, b0, pol_b);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc<KIND>(BN);
      int it = 0, seg = 0;
      for (int64_t u = u_begin; u < u_end; ++seg) {
        const int64_t tile = u / KB;
        const int kb0 = static_cast<int>(u - tile * KB);
        const int kb1 = static_cast<int>(min(KB, static_cast<int64_t>(kb0) + (u_end - u)));
        const int buf = seg & 1;
        mbar_wait(tmem_empty_bar + buf, ((seg >> 1) & 1) ^ 1);       // epilogue drained this buffer
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * kAccCols;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(full_bar + s, ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * S::kStage);
          const uint32_t sb = sa + S::kA;
#pragma unroll
          for (int w = 0; w < NB; ++w) {
            const uint64_t da = make_smem_desc(sa + (kSwap ? w * kTileM * kSwizzleBytes : 0));
            const uint64_t db = make_smem_desc(sb + (kSwap ? 0 : w * BN * kSwizzleBytes));
#pragma unroll
            for (int k = 0; k < BK / kUmmaK; ++k) {
              // advancing K inside the 128B swizzle atom = +32 bytes on the start address (>>4 => +2)
              umma<KIND>(acc + w * BN, da + 2 * k, db + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(empty_bar + s);               // frees the smem stage when these MMAs retire
        }
        umma_commit(tmem_full_bar + buf);           // this segment's accumulators are complete
        u += kb1 - kb0;
      }
    }
  } else {
    // ===== epilogue warps =====
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int et = threadIdx.x - 64;                // 0..127 among the epilogue threads
    const int64_t ldw = kSwap ? p.rows_a : p.rows_b;                 // row pitch of the [m, n] scratch plane
    const int64_t plane = p.rows_a * p.rows_b;
    int seg = 0;
    for (int64_t u = u_begin; u < u_end; ++seg) {
      const int64_t tile = u / KB;
      const int kb0 = static_cast<int>(u - tile * KB);
      const int kb1 = static_cast<int>(min(KB, static_cast<int64_t>(kb0) + (u_end - u)));
      u += kb1 - kb0;
      const int64_t a0 = (tile % p.tiles_a) * kTileM;
      const int64_t b0 = (tile / p.tiles_a) * BN;
      const int buf = seg & 1;
      const bool direct = kb0 == 0 && kb1 == KB;
      mbar_wait(tmem_full_bar + buf, (seg >> 1) & 1);
      tc_fence_after();
      const int64_t arow = a0 + q * 32 + lane;      // M-side row owned by this thread
      const uint32_t taddr = tmem_base + buf * kAccCols + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[NB][32];
#pragma unroll
        for (int w = 0; w < NB; ++w) {
          if constexpr (BN % 32 == 0) tmem_ld32(taddr + w * BN + c0, r[w]);
          else tmem_ld16(taddr + w * BN + c0, r[w]);
        }
        if (c0 + 32 >= BN) {                        // last chunk is in registers: hand the buffer back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty_bar + buf);
        }
        constexpr int kCols = (BN % 32 == 0) ? 32 : 16;
#pragma unroll
        for (int j = 0; j < kCols; ++j) {
          const int64_t brow = b0 + c0 + j;         // N-side row
          if (arow >= p.rows_a || brow >= p.rows_b) continue;
          const int64_t i = kSwap ? brow : arow;    // output row (m)
          const int64_t jn = kSwap ? arow : brow;   // output column (n)
          if (direct) {
            if constexpr (KIND != 0) float_epilogue_store<T>(p.fl, __uint_as_float(r[0][j]), i, jn);
            else if constexpr (NB == 2) glu_epilogue_store<T>(p.glu, static_cast<int32_t>(r[0][j]), static_cast<int32_t>(r[1][j]), i, jn);
            else dense_epilogue_store<T>(p.dense, static_cast<int32_t>(r[0][j]), i, jn);
          } else {
            if constexpr (KIND == 0) {
              atomicAdd(p.ws + i * ldw + jn, static_cast<int32_t>(r[0][j]));
              if constexpr (NB == 2) atomicAdd(p.ws + plane + i * ldw + jn, static_cast<int32_t>(r[1][j]));
            } else {
              atomicAdd(reinterpret_cast<float*>(p.ws) + i * ldw + jn, __uint_as_float(r[0][j]));
            }
          }
        }
      }
      if (direct) continue;
      // ---- shared tile: ticket; the last of the contributing CTAs finishes it ----
      __threadfence();
      epi_bar_sync();
      if (et == 0) {
        const int contributors = cta_of_unit((tile + 1) * KB - 1, U, P) - cta_of_unit(tile * KB, U, P) + 1;
        s_last = atomicAdd(p.counters + tile, 1) == contributors - 1;
      }
      epi_bar_sync();
      if (s_last) {
        __threadfence();
        for (int e = et; e < kTileM * BN; e += 128) {
          // consecutive threads -> consecutive output columns (n)
          const int64_t arow2 = kSwap ? a0 + e % kTileM : a0 + e / BN;
          const int64_t brow2 = kSwap ? b0 + e / kTileM : b0 + e % BN;
          if (arow2 >= p.rows_a || brow2 >= p.rows_b) continue;
          const int64_t i = kSwap ? brow2 : arow2, jn = kSwap ? arow2 : brow2;
          int32_t* w0 = p.ws + i * ldw + jn;
          const int32_t v = __ldcg(w0);
          *w0 = 0;
          if constexpr (KIND != 0) {
            float_epilogue_store<T>(p.fl, __int_as_float(v), i, jn);
          } else if constexpr (NB == 2) {
            int32_t* w1 = w0 + plane;
            const int32_t v2 = __ldcg(w1);
            *w1 = 0;
            glu_epilogue_store<T>(p.glu, v, v2, i, jn);
          } else {
            dense_epilogue_store<T>(p.dense, v, i, jn);
          }
        }
        if (et == 0) p.counters[tile] = 0;
      }
      epi_bar_sync();                               // s_last is reused by the next shared tile
    }
  }

  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---- host side ----
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CT2_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !ptr) throw std::runtime_error("cuTensorMapEncodeTiled is unavailable");
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

// [rows, k] row-major matrix of `elem` bytes; box = box_rows x 128 bytes, 128B swizzle, zero OOB fill.
CUtensorMap make_map(const void* base, int64_t rows, int64_t k, int elem, int kind, int box_rows) {
  CUtensorMap m;
  const CUtensorMapDataType dt = kind == 0 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                               : kind == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(k) * elem};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kSwizzleBytes / elem), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = get_encode_fn()(&m, dt, 2, const_cast<void*>(base), dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string(r));
  return m;
}

template <typename T, int KIND, int BN, int NB, bool kSwap>
void launch_tc(const void* x, const void* w, const void* w2, int64_t m, int64_t n, int64_t k, TcParams p,
               cudaStream_t st) {
  using S = TcSmem<BN, NB, kSwap>;
  constexpr int elem = KindTraits<KIND>::kElem;
  auto kernel = gemm_tc_kernel<T, KIND, BN, NB, kSwap>;
  static bool configured = false;
  if (!configured) {
    CT2_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(S::kBytes)));
    configured = true;
  }
  const CUtensorMap tmx = make_map(x, m, k, elem, KIND, kSwap ? BN : kTileM);
  const CUtensorMap tmw = make_map(w, n, k, elem, KIND, kSwap ? kTileM : BN);
  const CUtensorMap tmw2 = make_map(w2 ? w2 : w, n, k, elem, KIND, kSwap ? kTileM : BN);
  p.rows_a = kSwap ? n : m;
  p.rows_b = kSwap ? m : n;
  p.k = k;
  p.tiles_a = div_up(p.rows_a, kTileM);
  p.tiles_b = div_up(p.rows_b, BN);
  p.kb_total = div_up(k, kSwizzleBytes / elem);
  SplitKWorkspace& wsp = SplitKWorkspace::get(st);
  const int64_t tiles = static_cast<int64_t>(p.tiles_a) * p.tiles_b;
  const int64_t units = tiles * p.kb_total;
  int64_t ctas = std::min<int64_t>(wsp.sm_count, units);
  // tiles shared between CTAs go through the scratch: fall back to whole tiles per CTA when it cannot hold them
  const bool scratch_ok = static_cast<size_t>(m) * n * NB <= wsp.accum_elems && static_cast<size_t>(tiles) <= wsp.num_counters;
  p.whole_tiles = scratch_ok ? 0 : 1;
  if (!scratch_ok) ctas = std::min<int64_t>(wsp.sm_count, tiles);   // tile-aligned CTA ranges
  p.ws = wsp.accum;
  p.counters = wsp.counters;
  kernel<<<static_cast<unsigned>(ctas), kTcThreads, S::kBytes, st>>>(tmx, tmw, tmw2, p);
  check_launch();
}

template <typename T, int KIND, int NB>
void launch_tc_shape(const void* x, const void* w, const void* w2, int64_t m, int64_t n, int64_t k,
                     const TcParams& p, cudaStream_t st) {
  if (m <= 16) launch_tc<T, KIND, 16, NB, true>(x, w, w2, m, n, k, p, st);
  else if (m <= 32) launch_tc<T, KIND, 32, NB, true>(x, w, w2, m, n, k, p, st);
  else if (m <= 64) launch_tc<T, KIND, 64, NB, true>(x, w, w2, m, n, k, p, st);
  else if constexpr (NB == 2) launch_tc<T, KIND, 128, NB, false>(x, w, w2, m, n, k, p, st);
  else launch_tc<T, KIND, 256, NB, false>(x, w, w2, m, n, k, p, st);
}

}  // namespace

void gemm_s8_tc(const int8_t* A, const int8_t* B, int64_t M, int64_t N, int64_t K, const DenseEpilogue& epi,
                int dtype, cudaStream_t st) {
  if (M == 0 || N == 0) return;
  CT2_REQUIRE(K % 16 == 0, "gemm_s8: k must be a multiple of 16");
  TcParams p{};
  p.dense = epi;
  CT2_DISPATCH_DTYPE(dtype, (launch_tc_shape<T, 0, 1>(A, B, nullptr, M, N, K, p, st)));
}

void gemm_s8_glu_tc(const int8_t* A, const int8_t* Bgate, const int8_t* Bup, int64_t M, int64_t N, int64_t K,
                    const GluEpilogue& glu, int dtype, cudaStream_t st) {
  if (M == 0 || N == 0) return;
  CT2_REQUIRE(K % 16 == 0, "gemm_s8: k must be a multiple of 16");
  TcParams p{};
  p.glu = glu;
  CT2_DISPATCH_DTYPE(dtype, (launch_tc_shape<T, 0, 2>(A, Bgate, Bup, M, N, K, p, st)));
}

// a [m,k] T, b [n,k] T -> c [m,n] T (fp32 accumulate), T = f16 or bf16
void gemm_f16_tc(const void* A, const void* B, const void* bias, const void* residual, int act, int64_t M,
                 int64_t N, int64_t K, void* C, int dtype, cudaStream_t st) {
  if (M == 0 || N == 0) return;
  CT2_REQUIRE(K % 8 == 0, "gemm_f16: k must be a multiple of 8");
  CT2_REQUIRE(dtype == CT2B200_F16 || dtype == CT2B200_BF16, "gemm_f16: dtype must be float16 or bfloat16");
  TcParams p{};
  p.fl = FloatEpilogue{bias, residual, C, act, N};
  if (dtype == CT2B200_F16) launch_tc_shape<__half, 1, 1>(A, B, nullptr, M, N, K, p, st);
  else launch_tc_shape<__nv_bfloat16, 2, 1>(A, B, nullptr, M, N, K, p, st);
}

}  // namespace ct2b200
