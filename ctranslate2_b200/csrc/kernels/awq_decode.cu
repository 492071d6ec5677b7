// awq_decode.cu — AWQ-INT4 weight-streaming GEMM of the decode step (m <= 64) on tcgen05, A operand in TENSOR MEMORY.
//
// Replaces ops::GemmAwq / ops::GemvAwq (+ ops::Sum over split-K planes, + bias / activation / Mul) of the reference
// (src/ops/awq/gemm_gpu.cu, gemv_gpu.cu; dispatch src/layers/common.cc:402-438) by one kernel per Dense.
//
// Same plan as gemm_decode.cu (one tile per CTA, "swap AB": 128 output channels = the UMMA M side, activations = N side;
// DSMEM split-K cluster chosen so that tiles x CS fills the SMs in one wave; weights prefetched before griddepcontrol.wait),
// plus what 4-bit weights need:
//   * the only bytes that come from HBM are the packed nibbles.  A ring slot holds a SUPER-BLOCK of four K blocks (256 input
//     channels): per weight one TMA box of 128 rows x 128 bytes (SWIZZLE_128B, so that the 32 rows a warp reads land in
//     different banks), the {scale, zero} pairs of its groups (512 B each, cp.async.bulk from the group-major array) and the
//     four activation atoms, so the transform warps never touch global memory.  The box width matters: TMA works row by row,
//     and with one box per 64-channel block (128 rows x 32 bytes — the first version of this kernel) the row requests, not the
//     bytes, set the pace: ncu showed every transform warp parked on the "slot full" barrier, 10 GB/s per SM and DRAM at 7-16 %.
//     (Before that, round 1 fetched the pairs with per-thread global loads inside the loop: 64 % long_scoreboard.)
//   * the dequantized fp16 operand never goes back to shared memory.  int4 -> fp16 at HBM speed is 5.8 TB/s of nibbles =
//     23 TB/s of fp16, i.e. 82 B/clk/SM written + 82 B/clk/SM read by the tensor core: more than the 128 B/clk of the shared
//     memory port.  Instead the thread that owns output channel r converts the 64 channels of its row in registers
//     ((q - z) * s: exact subtraction, one fp16 rounding — the arithmetic of the reference's dequantize_s4_to_fp16x2 +
//     sub.f16x2 + fma.rn.f16x2) and writes them with ONE tcgen05.st.32x32b.x32 into TMEM lane r, columns [32 x stage): the
//     K-major A layout tcgen05.mma reads directly (cute::UMMA::tmem_frg, M = 128: lane = row, two fp16 per 32-bit column).
//   * tcgen05.mma.kind::f16 with A from TMEM, B (activations) from 128B-swizzled smem, fp32 accumulators in TMEM;
//     epilogue = gemm_decode_common.cuh (float arm).
#include <map>
#include <mutex>
#include <tuple>

#include "awq_common.cuh"
#include "gemm_decode_common.cuh"
#include "kernels.h"

namespace ct2b200 {
namespace {

using namespace tc;
using namespace dec;

constexpr int kThreads = 704;          // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue, warps 6-21 transform
constexpr int kDeqWarps = 16;          // 4 transform groups of 4 warps; group g owns the K blocks it % 4 == g, so four
constexpr int kGroups = 4;             // blocks are converted CONCURRENTLY (LDS -> lop3/hfma -> tcgen05.st is a latency chain)
constexpr int kGroupWarps = kDeqWarps / kGroups;
constexpr int kBKh = 64;               // fp16 channels per K block (one 128-byte swizzle atom of the activation operand)
constexpr int kSub = 4;                         // K blocks per ring slot = transform groups: group g converts sub-block g
constexpr int kSlotK = kSub * kBKh;             // 256 input channels per slot
constexpr int kPacked = kTileM * kSlotK / 2;    // 16 KB of nibbles per weight per slot (128 rows x 128 B, SWIZZLE_128B)
constexpr int kPairBytes = 512;                 // the 128 {scale, zero} pairs of one group of a weight
constexpr int kPairs = kSub * kPairBytes;       // up to 4 groups per slot (group 64); 2 KB keeps what follows 1024-byte aligned
constexpr int kAColsPerBlock = kBKh / 2;        // 32 TMEM columns hold the 64 fp16 channels of a block
// TMEM: TWO accumulator sets (one per MMA issuer: a single thread issuing the 8-32 small MMAs of a slot, with their barrier
// waits and commits, was the pace-setter of the kernel — 1 700 of 2 400 cycles per slot in the stamp trace) in columns
// [0, 2 * NB * BN), dequantized A stages above.  Block `it` uses A stage it % kAStages.
constexpr int kIssuers = 2;
template <int BN, int NB> struct AStages {
  static constexpr int acc_cols = kIssuers * NB * BN;
  static constexpr int raw = (512 - acc_cols) / (NB * kAColsPerBlock);
  static constexpr int value = raw > 12 ? 12 : raw;             // 14 -> 12 | 7 / 6 / 4 (NB = 2, BN = 16 / 32 / 64)
};
constexpr int kMaxAStages = 12;
constexpr int kMaxP = 8;

struct AwqDecParams {
  DecParams d;
  int64_t k;
  int group;
  int p_stages;                  // depth of the packed / pairs / activation ring
  int sb_total;                  // super-blocks (256 channels) along K
  const __half2* sz[2];          // {scale, zero} [k/group, n] (group-major): 512 contiguous bytes per tile and group
};

template <int BN, int NB>
struct AwqDecSmem {
  static constexpr int kAct = BN * kSwizzleBytes;                           // one activation atom (64 channels)
  static constexpr int kP = NB * (kPacked + kPairs) + kSub * kAct;          // one ring slot: nibbles | pairs | 4 activation atoms
  static constexpr int kCtrl = 1024;
  static size_t red_bytes(int cs) { return cs > 1 ? static_cast<size_t>(cs) * NB * (BN / 16) * ((16 + cs - 1) / cs) * kTileM * 4 : 0; }
  static size_t bytes(int p_stages, int cs) {
    return static_cast<size_t>(p_stages) * kP + kCtrl + red_bytes(cs) + 1024;
  }
};

// D[tmem] (+)= A[tmem] * B[smem descriptor]
__device__ __forceinline__ void umma_ts_f16(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate));
}
// 32 lanes x 32 columns: thread t of the warp writes r[0..31] to its lane, columns taddr.col + [0, 32)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// global -> shared bulk copy (bytes % 16 == 0, both addresses 16-byte aligned), completion on an mbarrier
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// -DCT2B200_AWQ_TRACE (python -m ctranslate2_b200.build --variant awqtrace): CTA 0 records SM cycle stamps of the pipeline
// events of its first 16 super-blocks and prints them when it is done.  Not compiled into the product library.
#ifdef CT2B200_AWQ_TRACE
#define AWQ_TRACE_DECL __shared__ long long s_trace[16][12];
#define AWQ_TRACE(sb, slot) do { if (blockIdx.x == 0 && (sb) < 16) s_trace[sb][slot] = clock64(); } while (0)
#else
#define AWQ_TRACE_DECL
#define AWQ_TRACE(sb, slot) do { } while (0)
#endif

template <int BN, int NB, int CS>
__global__ void __launch_bounds__(kThreads, 1)
    awq_decode_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                      const __grid_constant__ CUtensorMap tm_w2, const AwqDecParams ap) {
  using S = AwqDecSmem<BN, NB>;
  using T = __half;
  const DecParams& p = ap.d;
  constexpr uint32_t kTmemCols = 512;
  constexpr int cp16 = (16 + CS - 1) / CS;
  constexpr int cpr = (BN / 16) * cp16;
  constexpr int kAStages = AStages<BN, NB>::value;
  constexpr int kAccCols = AStages<BN, NB>::acc_cols;      // accumulator set i at columns [i * NB * BN, +NB * BN)
  static_assert(kAStages >= kSub, "every transform group needs an A stage");

  extern __shared__ uint8_t smem_raw[];
  AWQ_TRACE_DECL
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* p_ring = smem;                                         // [p_stages][nibbles NB x 4 KB | pairs NB x 512 B | x BN x 128 B]
  const int PD = ap.p_stages;
  uint8_t* ctrl = p_ring + static_cast<size_t>(PD) * S::kP;
  uint64_t* p_full = reinterpret_cast<uint64_t*>(ctrl);           // [kMaxP] TMA landed
  uint64_t* p_free = p_full + kMaxP;                              // [kMaxP] transform warps + the MMA commit
  uint64_t* a_ready = p_free + kMaxP;                             // [kAStages] A stage written to TMEM (the group's warps)
  uint64_t* a_free = a_ready + kMaxAStages;                       // [kAStages] MMAs that read it retired
  uint64_t* acc_bar = a_free + kMaxAStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
  uint32_t* red = reinterpret_cast<uint32_t*>(ctrl + S::kCtrl);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x / CS;
  const int crank = CS > 1 ? static_cast<int>(blockIdx.x % CS) : 0;
  const int sb_lo = crank * ap.sb_total / CS, sb_hi = (crank + 1) * ap.sb_total / CS;
  const int nsb = sb_hi - sb_lo;                                   // super-blocks of this CTA
  const int a0 = tile * p.tile_rows;
  const int ngs = max(1, kSlotK / ap.group);                       // distinct groups inside a super-block (group >= 64)

  if (threadIdx.x == 0) {
    for (int s = 0; s < PD; ++s) {
      mbar_init(p_full + s, 1);
      mbar_init(p_free + s, kDeqWarps + kIssuers);
    }
    for (int s = 0; s < kAStages; ++s) {
      mbar_init(a_ready + s, kGroupWarps);
      mbar_init(a_free + s, 1);
    }
    mbar_init(acc_bar, kIssuers);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch();
  if (CS > 1) cluster_arrive();

  // ===== MMA issuer `half` (warp 1: 0, warp 2: 1): sub-blocks 2 half, 2 half + 1 of every slot into accumulator set `half` =====
  auto mma_issuer = [&](int half) {
    constexpr uint32_t idesc = make_idesc<1>(BN);
    const uint32_t acc = tmem_base + half * (NB * BN);
#pragma unroll 1
    for (int sb = 0; sb < nsb; ++sb) {
      const int sp = sb % PD;
      mbar_wait(p_full + sp, (sb / PD) & 1);                    // activations of this super-block landed
      if (half == 0) AWQ_TRACE(sb, 1);
      const uint32_t act0 = smem_u32(p_ring + static_cast<size_t>(sp) * S::kP + NB * (kPacked + kPairs));
#pragma unroll 1
      for (int jj = 0; jj < kSub / kIssuers; ++jj) {
        const int j = half * (kSub / kIssuers) + jj;
        const int it = sb * kSub + j, sa = it % kAStages;
        mbar_wait(a_ready + sa, (it / kAStages) & 1);           // weights dequantized into TMEM
        if (half == 0 && jj == 0) AWQ_TRACE(sb, 2);
        if (half == 0 && jj == 1) AWQ_TRACE(sb, 3);
        tc_fence_after();
        const uint32_t ta = tmem_base + kAccCols + sa * (NB * kAColsPerBlock);
        const uint64_t db = make_smem_desc(act0 + j * S::kAct);
#pragma unroll
        for (int w = 0; w < NB; ++w)
#pragma unroll
          for (int k = 0; k < kBKh / 16; ++k)                   // K = 16 per instruction = 8 TMEM columns of A
            umma_ts_f16(acc + w * BN, ta + w * kAColsPerBlock + k * 8, db + 2 * k, idesc, (sb > 0 || jj > 0 || k > 0) ? 1u : 0u);
        umma_commit(a_free + sa);
      }
      umma_commit(p_free + sp);
      if (half == 0) AWQ_TRACE(sb, 4);
    }
    umma_commit(acc_bar);
  };

  if (warp == 0) {
    // ===== TMA producer: packed nibbles, {scale, zero} pairs (both weights are constants: issued before the grid
    // dependency resolves) and the activation block =====
    if (elect_one()) {
      const int rows_here = static_cast<int>(min(static_cast<int64_t>(p.tile_rows), p.n - a0));
      const uint32_t pair_bytes = static_cast<uint32_t>(rows_here) * 4u;
      const uint32_t tx = static_cast<uint32_t>(NB * p.tile_rows * (kSlotK / 2)) + NB * ngs * pair_bytes + kSub * S::kAct;
      auto weights = [&](int s, int sb) {
        uint8_t* st = p_ring + static_cast<size_t>(s) * S::kP;
        const int64_t g0 = (static_cast<int64_t>(sb) * kSlotK) / ap.group;
        tma_load_2d(st, &tm_w, p_full + s, sb * (kSlotK / 2), a0, kEvictFirst);
        if (NB == 2) tma_load_2d(st + kPacked, &tm_w2, p_full + s, sb * (kSlotK / 2), a0, kEvictFirst);
        for (int g = 0; g < ngs; ++g) {
          bulk_load(st + NB * kPacked + g * kPairBytes, ap.sz[0] + (g0 + g) * p.n + a0, pair_bytes, p_full + s);
          if (NB == 2) bulk_load(st + NB * kPacked + kPairs + g * kPairBytes, ap.sz[1] + (g0 + g) * p.n + a0, pair_bytes, p_full + s);
        }
      };
      auto acts = [&](int s, int sb) {
        uint8_t* at = p_ring + static_cast<size_t>(s) * S::kP + NB * (kPacked + kPairs);
#pragma unroll
        for (int j = 0; j < kSub; ++j) tma_load_2d(at + j * S::kAct, &tm_x, p_full + s, (sb * kSub + j) * kBKh, 0, kEvictLast);
      };
      const int pre = min(PD, nsb);
#pragma unroll 1
      for (int i = 0; i < pre; ++i) {
        AWQ_TRACE(i, 0);
        mbar_expect_tx(p_full + i, tx);
        weights(i, sb_lo + i);
      }
      griddep_wait();
#pragma unroll 1
      for (int i = 0; i < pre; ++i) acts(i, sb_lo + i);
#pragma unroll 1
      for (int it = pre; it < nsb; ++it) {
        const int s = it % PD;
        mbar_wait(p_free + s, ((it / PD) & 1) ^ 1);
        AWQ_TRACE(it, 0);
        mbar_expect_tx(p_full + s, tx);
        weights(s, sb_lo + it);
        acts(s, sb_lo + it);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) mma_issuer(0);
  } else if (warp >= 6) {
    // ===== transform warps: nibbles -> fp16 (q - z) * s, registers -> TMEM =====
    // thread -> (transform group, tile row).  A warp may only touch TMEM lanes [32 * (warp % 4), +32): the row quadrant of a
    // warp is warp % 4 (any bijection inside the group of four consecutive warps works).
    const int grp = (warp - 6) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    // group grp converts sub-block grp of every super-block; its 32 bytes of row r are the 16-byte chunks 2 grp, 2 grp + 1
    // of the row's 128 bytes, stored at chunk ^ (r % 8) by the 128-byte swizzle
    const int pair_slot = (grp * kBKh) / ap.group < ngs ? (grp * kBKh) / ap.group : ngs - 1;
    const uint32_t ch0 = static_cast<uint32_t>((2 * grp) ^ (r & 7)) * 16u, ch1 = static_cast<uint32_t>((2 * grp + 1) ^ (r & 7)) * 16u;
#pragma unroll 1
    for (int sb = 0; sb < nsb; ++sb) {
      const int it = sb * kSub + grp;
      const int sp = sb % PD, sa = it % kAStages;
      mbar_wait(p_full + sp, (sb / PD) & 1);
      if (threadIdx.x == 6 * 32) AWQ_TRACE(sb, 5);
      if (threadIdx.x == 18 * 32) AWQ_TRACE(sb, 9);
      const uint8_t* pk = p_ring + static_cast<size_t>(sp) * S::kP;
      const uint32_t ta = tmem_base + lane_base + kAccCols + sa * (NB * kAColsPerBlock);
#pragma unroll
      for (int w = 0; w < NB; ++w) {
        uint32_t v[32];
        const __half2 sz = *reinterpret_cast<const __half2*>(pk + NB * kPacked + w * kPairs + pair_slot * kPairBytes + r * 4);
        const __half sc = __low2half(sz), zp = __high2half(sz);
        const __half2 zb = __half2half2(__hadd(__float2half(1024.f), zp));
        const __half2 zt = __half2half2(__hneg(__hadd(__float2half(64.f), zp)));
        const __half2 s2 = __half2half2(sc);
        const uint8_t* row = pk + w * kPacked + r * kSwizzleBytes;
        const uint4 w0 = *reinterpret_cast<const uint4*>(row + ch0);
        const uint4 w1 = *reinterpret_cast<const uint4*>(row + ch1);
        const uint32_t words[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int c = 0; c < 8; ++c) {                  // word c = channels 8c .. 8c+7 = TMEM columns 4c .. 4c+3
          const uint4 d = awq_dequant_word(words[c], zb, zt, s2);
          v[4 * c + 0] = d.x;
          v[4 * c + 1] = d.y;
          v[4 * c + 2] = d.z;
          v[4 * c + 3] = d.w;
        }
        // the A stage is free once the MMAs of block it - kAStages have retired (waited for AFTER the first conversion)
        if (w == 0) {
          if (threadIdx.x == 6 * 32) AWQ_TRACE(sb, 6);
          if (it >= kAStages) {
            mbar_wait(a_free + sa, ((it / kAStages) & 1) ^ 1);
            tc_fence_after();
          }
          if (threadIdx.x == 6 * 32) AWQ_TRACE(sb, 7);
        }
        tmem_st32(ta + w * kAColsPerBlock, v);
      }
      tmem_st_wait();
      if (threadIdx.x == 6 * 32) AWQ_TRACE(sb, 8);
      if (threadIdx.x == 18 * 32) AWQ_TRACE(sb, 10);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(a_ready + sa);
        mbar_arrive(p_free + sp);
      }
    }
  } else {
    // ===== epilogue warps (2..5): thread = output channel =====
    const int q = warp & 3;
    const int rloc = q * 32 + lane;
    const int64_t arow = static_cast<int64_t>(a0) + rloc;
    const bool row_ok = rloc < p.tile_rows && arow < p.n;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    if (warp == 2) {                                   // the second MMA issuer lives in an epilogue warp (idle until the end)
      if (elect_one()) mma_issuer(1);
      __syncwarp();
    }
    griddep_wait();
    float bias_t = 0.f;
    if (row_ok && p.bias) bias_t = to_f32(static_cast<const T*>(p.bias)[arow]);
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    // the two accumulator sets are partial sums over alternate halves of every slot
    auto load_acc = [&](int c0, uint32_t (&r)[NB][16]) {
#pragma unroll
      for (int w = 0; w < NB; ++w) {
        uint32_t r1[16];
        tmem_ld16x16(taddr + w * BN + c0, r[w]);
        tmem_ld16x16(taddr + NB * BN + w * BN + c0, r1);
#pragma unroll
        for (int j = 0; j < 16; ++j) r[w][j] = __float_as_uint(__uint_as_float(r[w][j]) + __uint_as_float(r1[j]));
      }
    };
    if constexpr (CS == 1) {
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t r[NB][16];
        load_acc(c0, r);
        if (row_ok && c0 < p.m) dec_finish<T, 1, NB, 16>(p, r, arow, c0, 1, 16, 1.f, 1.f, bias_t);
      }
    } else {
      cluster_wait();
      uint32_t peer[CS];
#pragma unroll
      for (int o = 0; o < CS; ++o) {
        const uint32_t local = smem_u32(red + static_cast<size_t>(crank) * NB * cpr * kTileM + rloc);
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer[o]) : "r"(local), "r"(o));
      }
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t r[NB][16];
        load_acc(c0, r);
        const uint32_t chunk_off = static_cast<uint32_t>((c0 / 16) * cp16 * kTileM * 4);
#pragma unroll
        for (int j = 0; j < 16; ++j)
#pragma unroll
          for (int w = 0; w < NB; ++w) {
            const uint32_t off = static_cast<uint32_t>((w * cpr + j / CS) * kTileM * 4);
            asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(peer[j % CS] + chunk_off + off), "r"(r[w][j]) : "memory");
          }
      }
    }
  }

  if constexpr (CS > 1) {
    __syncwarp();
    if (warp < 2 || warp >= 6) cluster_wait();         // phase 1 (the epilogue warps consumed it above)
    cluster_arrive();
    cluster_wait();
    if (warp >= 2 && warp < 6) {
      const int q = warp & 3;
      const int rloc = q * 32 + lane;
      const int64_t arow = static_cast<int64_t>(a0) + rloc;
      const bool row_ok = rloc < p.tile_rows && arow < p.n;
      float bias_t = 0.f;
      if (row_ok && p.bias) bias_t = to_f32(static_cast<const T*>(p.bias)[arow]);
      const int nvalid = (16 - crank + CS - 1) / CS;
#pragma unroll 1
      for (int ch = 0; ch < BN / 16; ++ch) {
        uint32_t r[NB][cp16];
#pragma unroll
        for (int w = 0; w < NB; ++w)
#pragma unroll
          for (int jj = 0; jj < cp16; ++jj) {
            float acc = 0.f;
#pragma unroll
            for (int src = 0; src < CS; ++src)         // fixed rank order: deterministic
              acc += __uint_as_float(red[(static_cast<size_t>(src * NB + w) * cpr + ch * cp16 + jj) * kTileM + rloc]);
            r[w][jj] = __float_as_uint(acc);
          }
        const int col0 = ch * 16 + crank;
        if (row_ok && col0 < p.m) dec_finish<T, 1, NB, cp16>(p, r, arow, col0, CS, nvalid, 1.f, 1.f, bias_t);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
#ifdef CT2B200_AWQ_TRACE
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const long long t0 = s_trace[0][0];
    printf("awq trace BN=%d NB=%d CS=%d PD=%d nsb=%d: per super-block [tma issue | mma: acts landed, a_ready(0), a_ready(3), commit | "
           "group0: slot full, converted, a_free, st done | group3: slot full, st done] cycles since the first issue\n",
           BN, NB, CS, PD, nsb);
    for (int sb = 0; sb < min(nsb, 16); ++sb)
      printf("  sb %2d: %6lld | %6lld %6lld %6lld %6lld | %6lld %6lld %6lld %6lld | %6lld %6lld\n", sb, s_trace[sb][0] - t0,
             s_trace[sb][1] - t0, s_trace[sb][2] - t0, s_trace[sb][3] - t0, s_trace[sb][4] - t0, s_trace[sb][5] - t0,
             s_trace[sb][6] - t0, s_trace[sb][7] - t0, s_trace[sb][8] - t0, s_trace[sb][9] - t0, s_trace[sb][10] - t0);
  }
#endif
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---- host side ----
struct AwqPlan {
  int cs = 0, tile_rows = 128, tiles = 0, p_stages = 4;
};

CUtensorMap make_packed_map(const void* wp, int64_t n, int64_t k, int box_rows) {
  // wp as bytes [n, k/2]; box = box_rows x 128 bytes (a super-block of 256 channels), 128-byte swizzle
  CUtensorMap m;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(k / 2), static_cast<cuuint64_t>(n)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(k / 2)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kSlotK / 2), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = get_tensor_map_encoder()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(wp), dims, strides, box,
                                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled (awq) failed with code " + std::to_string(r));
  return m;
}

template <int BN, int NB>
int p_stages_for(int cs, int nkb) {
  using S = AwqDecSmem<BN, NB>;
#ifdef CT2B200_AWQ_TRACE
  const size_t cap = 212 * 1024;          // room for the static trace buffer
#else
  const size_t cap = 220 * 1024;
#endif
  const size_t fixed = S::kCtrl + S::red_bytes(cs) + 1024;
  if (fixed + 2 * S::kP > cap) return 0;
  int st = static_cast<int>((cap - fixed) / S::kP);
  st = std::min(st, kMaxP);
  return std::max(2, std::min(st, std::max(nkb, 2)));
}

template <int BN, int NB, int CS>
void configure_once() {
#ifdef CT2B200_AWQ_TRACE
  allow_dynamic_smem(awq_decode_kernel<BN, NB, CS>, 222 * 1024);
#else
  allow_dynamic_smem(awq_decode_kernel<BN, NB, CS>, 226 * 1024);
#endif
}

template <int BN, int NB, int CS>
int clusters_for(int p_stages, int sm_count) {
  configure_once<BN, NB, CS>();
  static std::mutex mu;
  static std::map<std::pair<int, int>, int> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find({dev, p_stages});
  if (it != cache.end()) return it->second;
  const int n = max_clusters(awq_decode_kernel<BN, NB, CS>, CS, kThreads, AwqDecSmem<BN, NB>::bytes(p_stages, CS), sm_count);
  cache[{dev, p_stages}] = n;
  return n;
}

template <int BN, int NB>
AwqPlan plan_awq(int64_t n, int kb_total /* super-blocks */, int sm_count) {
  static std::mutex mu;
  static std::map<std::tuple<int, int64_t, int>, AwqPlan> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  const int force_cs = env_int("CT2B200_GEMM_CS", 0);
  const int force_rows = env_int("CT2B200_GEMM_ROWS", 0);
  const bool forced = force_cs != 0 || force_rows != 0;
  if (!forced) {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find({dev, n, kb_total});
    if (it != cache.end()) return it->second;
  }
  AwqPlan best;
  double best_cost = 1e30;
  for (int cs = 1; cs <= 4; ++cs) {
    if (force_cs && cs != force_cs) continue;
    if (cs > 1 && kb_total < 2 * cs) continue;
    const int nkb = (kb_total + cs - 1) / cs;
    const int ps = p_stages_for<BN, NB>(cs, nkb);
    if (ps == 0) continue;
    int maxc = 0;
    switch (cs) {
      case 1: maxc = clusters_for<BN, NB, 1>(ps, sm_count); break;
      case 2: maxc = clusters_for<BN, NB, 2>(ps, sm_count); break;
      case 3: maxc = clusters_for<BN, NB, 3>(ps, sm_count); break;
      default: maxc = clusters_for<BN, NB, 4>(ps, sm_count); break;
    }
    // Tile height.  The transform converts all 128 rows of a block whatever the tile height, so full-height tiles waste nothing;
    // but since the kernel waits on the HBM stream rather than on the conversion (stamp trace, profiles/README.md), a shorter
    // tile that puts the stream on more SMs wins when 128-row tiles leave many idle: gate/up of Llama-3-8B is 112 tiles of 128
    // rows on 148 SMs, 138 tiles of 104.  cost = streamed rows x K blocks per CTA (+ the cluster exchange).
    const int full = static_cast<int>((n + 127) / 128);
    int cand[2] = {128, 128};
    if (!force_rows && full * 10 < maxc * 9) {
      const int r = static_cast<int>((((n + maxc - 1) / maxc) + 7) / 8 * 8);
      if (r >= 64 && r < 128) cand[1] = r;
    }
    for (int ci = 0; ci < 2; ++ci) {
      const int rows = force_rows ? force_rows : cand[ci];
      const int tiles = static_cast<int>((n + rows - 1) / rows);
      if (tiles > maxc) continue;
      const double cost = static_cast<double>(nkb) * (rows / 128.0) + (cs > 1 ? 1.0 : 0.0);
      if (cost < best_cost) {
        best_cost = cost;
        best.cs = cs;
        best.tile_rows = rows;
        best.tiles = tiles;
        best.p_stages = ps;
      }
    }
  }
  if (!forced) {
    std::lock_guard<std::mutex> lock(mu);
    cache[{dev, n, kb_total}] = best;
  }
  return best;
}

template <int BN, int NB, int CS>
void launch(const CUtensorMap& tmx, const CUtensorMap& tmw, const CUtensorMap& tmw2, const AwqDecParams& p, const AwqPlan& plan,
            cudaStream_t st) {
  configure_once<BN, NB, CS>();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(plan.tiles * CS));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = AwqDecSmem<BN, NB>::bytes(plan.p_stages, CS);
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CS > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CS;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  CT2_CUDA_CHECK(cudaLaunchKernelEx(&cfg, awq_decode_kernel<BN, NB, CS>, tmx, tmw, tmw2, p));
  check_launch();
}

template <int BN, int NB>
bool run(const void* x, const AwqNative& w, const AwqNative* w2, int64_t m, AwqDecParams p, cudaStream_t st) {
  const int sb_total = static_cast<int>(w.k / kSlotK);
  const AwqPlan plan = plan_awq<BN, NB>(w.n, sb_total, sm_count_of_current_device());
  if (plan.cs == 0) return false;
  p.d.n = w.n;
  p.d.m = m;
  p.d.kb_total = sb_total * kSub;
  p.sb_total = sb_total;
  p.d.tile_rows = plan.tile_rows;
  p.d.stages = plan.p_stages;
  p.k = w.k;
  p.group = w.group;
  p.p_stages = plan.p_stages;
  p.sz[0] = static_cast<const __half2*>(w.sz);
  p.sz[1] = static_cast<const __half2*>(w2 ? w2->sz : w.sz);
  const CUtensorMap tmx = make_operand_map(x, m, w.k, 2, 1, BN);
  const CUtensorMap tmw = make_packed_map(w.wp, w.n, w.k, plan.tile_rows);
  const CUtensorMap tmw2 = make_packed_map(w2 ? w2->wp : w.wp, w.n, w.k, plan.tile_rows);
  switch (plan.cs) {
    case 1: launch<BN, NB, 1>(tmx, tmw, tmw2, p, plan, st); break;
    case 2: launch<BN, NB, 2>(tmx, tmw, tmw2, p, plan, st); break;
    case 3: launch<BN, NB, 3>(tmx, tmw, tmw2, p, plan, st); break;
    default: launch<BN, NB, 4>(tmx, tmw, tmw2, p, plan, st); break;
  }
  return true;
}

template <int NB>
bool run_m(const void* x, const AwqNative& w, const AwqNative* w2, int64_t m, const AwqDecParams& p, cudaStream_t st) {
  if (m <= 16) return run<16, NB>(x, w, w2, m, p, st);
  if (m <= 32) return run<32, NB>(x, w, w2, m, p, st);
  return run<64, NB>(x, w, w2, m, p, st);
}

// CT2B200_AWQ_DECODE=0 falls back to the general stream-K kernel of awq.cu (CT2B200_AWQ_DECODE_GLU does the same for the
// fused gate/up launch only).  Measured in the decode graph of a Llama-3-8B AWQ model (tools/decode_once.py, B200): with the
// four concurrent transform groups this kernel takes 3.93 ms / step at bsz 1 and 4.84 ms at bsz 32, the general kernel
// 4.04 / 6.0 ms.  Both are bound by the int4 -> fp16 transform (~1 TB/s of packed weights), not by HBM.
bool enabled() { return env_int("CT2B200_AWQ_DECODE", CT2B200_DEFAULT_AWQ_DECODE) != 0; }
bool glu_enabled() { return env_int("CT2B200_AWQ_DECODE_GLU", env_int("CT2B200_AWQ_DECODE", CT2B200_DEFAULT_AWQ_DECODE)) != 0; }

}  // namespace

// false = shape not covered (more tiles than one wave holds, group not a multiple of 64): caller uses gemm_awq_tc_kernel
bool dense_awq_decode(const void* x, const AwqNative& w, const void* bias, const void* residual, int act, int64_t m, void* y,
                      cudaStream_t st) {
  if (!enabled() || m < 1 || m > 64 || w.group % kBKh != 0 || w.k % kSlotK != 0 || w.sz == nullptr || w.n % 8 != 0) return false;
  AwqDecParams p{};
  p.d.bias = bias;
  p.d.residual = residual;
  p.d.y = y;
  p.d.act = act;
  p.d.ldy = w.n;
  return run_m<1>(x, w, nullptr, m, p, st);
}

bool dense_awq_glu_decode(const void* x, const AwqNative& wg, const AwqNative& wu, int act, int64_t m, void* h, cudaStream_t st) {
  if (!glu_enabled() || m < 1 || m > 64 || wg.group % kBKh != 0 || wg.k % kSlotK != 0 || wg.group != wu.group ||
      wg.sz == nullptr || wu.sz == nullptr || wg.n % 8 != 0)
    return false;
  AwqDecParams p{};
  p.d.y = h;
  p.d.act = act;
  p.d.ldy = wg.n;
  return run_m<2>(x, wg, &wu, m, p, st);
}

}  // namespace ct2b200
