// attention_mma.cu — causal prefill attention on tensor cores for fp16/bf16 activations.
// Flash-attention style: a CTA owns 64 query rows of one (batch, head); 4 warps x 16 rows; K/V tiles of 64
// keys are staged through shared memory with cp.async (double buffered) straight from the un-replicated GQA
// cache [slot, Hkv, max_len, D]; S = Q K^T and O += P V run on mma.sync.m16n8k16 with fp32 accumulation and
// an fp32 online softmax; P never leaves registers.  Replaces MatMul + SoftMax + MatMul over a materialised
// [B,H,T,S] score tensor (reference src/layers/attention.cc:178-287, 536-602).
// (Decode attention is HBM-bound and lives in attention.cu; tcgen05 needs M >= 64 rows per head-tile, which
// a 16-row warp tile does not give, so the warp-level MMA is the right instrument for this shape.)
#include <type_traits>

#include <cstdlib>

#include "../common.cuh"
#include "kernels.h"

namespace ct2b200 {

namespace {

constexpr int kQTile = 64, kKTile = 64, kThreads = 128;

__device__ __forceinline__ void cp16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const void* p) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}
__device__ __forceinline__ void ldsm4_t(uint32_t (&r)[4], const void* p) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}
template <typename T>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (sizeof(T) == 2 && std::is_same<T, __half>::value) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <typename T, int D>
__global__ void __launch_bounds__(kThreads)
    attention_prefill_mma_kernel(const T* __restrict__ qkv, const T* __restrict__ k_cache, const T* __restrict__ v_cache,
                                 int64_t time, int64_t offset, int H, int Hkv, int64_t max_len, float scale_log2,
                                 T* __restrict__ out) {
  constexpr int LD = D + 8;                   // padded smem pitch (elements): conflict-free ldmatrix
  constexpr int CH = D / 8;                   // 16-byte chunks per row
  extern __shared__ __align__(16) uint8_t smem_raw[];
  T* sQ = reinterpret_cast<T*>(smem_raw);     // [64][LD]
  T* sK = sQ + kQTile * LD;                   // [2][64][LD]
  T* sV = sK + 2 * kKTile * LD;               // [2][64][LD]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t q0 = static_cast<int64_t>(blockIdx.x) * kQTile;
  const int h = blockIdx.y;
  const int64_t b = blockIdx.z;
  const int kvh = h / (H / Hkv);
  const int64_t row_w = static_cast<int64_t>(H + 2 * Hkv) * D;
  const T* qbase = qkv + (b * time + q0) * row_w + static_cast<int64_t>(h) * D;
  const T* kc = k_cache + (b * Hkv + kvh) * max_len * D;
  const T* vc = v_cache + (b * Hkv + kvh) * max_len * D;

  // keys visible to this query tile: 0 .. offset + min(q0+63, time-1)
  const int64_t last_q = min(q0 + kQTile, time) - 1;
  const int nkeys = static_cast<int>(offset + last_q + 1);
  const int ntiles = (nkeys + kKTile - 1) / kKTile;

  auto load_kv = [&](int stage, int kt) {
    const int64_t k0 = static_cast<int64_t>(kt) * kKTile;
    for (int c = tid; c < kKTile * CH; c += kThreads) {
      const int r = c / CH, ch = c % CH;
      const bool ok = k0 + r < nkeys;
      const int64_t off = (ok ? k0 + r : 0) * D + ch * 8;
      cp16(sK + (stage * kKTile + r) * LD + ch * 8, kc + off, ok);
      cp16(sV + (stage * kKTile + r) * LD + ch * 8, vc + off, ok);
    }
  };
  // Q tile + first K/V tile
  for (int c = tid; c < kQTile * CH; c += kThreads) {
    const int r = c / CH, ch = c % CH;
    const bool ok = q0 + r < time;
    cp16(sQ + r * LD + ch * 8, qbase + (ok ? r : 0) * row_w + ch * 8, ok);
  }
  load_kv(0, 0);
  asm volatile("cp.async.commit_group;\n" ::);

  float o[D / 8][4];
#pragma unroll
  for (int j = 0; j < D / 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  uint32_t qf[D / 16][4];
  const int64_t qpos0 = offset + q0 + warp * 16 + g;       // absolute position of row g (row g+8: +8)

  for (int kt = 0; kt < ntiles; ++kt) {
    const int stage = kt & 1;
    if (kt + 1 < ntiles) load_kv(stage ^ 1, kt + 1);
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 1;\n" ::);
    __syncthreads();
    if (kt == 0) {
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk)      // A fragments of this warp's 16 query rows
        ldsm4(qf[kk], sQ + (warp * 16 + (lane & 15)) * LD + kk * 16 + (lane >> 4) * 8);
    }
    const T* ks = sK + stage * kKTile * LD;
    const T* vs = sV + stage * kKTile * LD;

    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[kKTile / 8][4];
#pragma unroll
    for (int j = 0; j < kKTile / 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
#pragma unroll
      for (int j = 0; j < kKTile / 8; j += 2) {
        uint32_t bf[4];   // {b0,b1} of key tile j, {b0,b1} of key tile j+1
        ldsm4(bf, ks + (j * 8 + (lane & 7) + (lane >> 4) * 8) * LD + kk * 16 + ((lane >> 3) & 1) * 8);
        mma16816<T>(s[j], qf[kk], bf[0], bf[1]);
        mma16816<T>(s[j + 1], qf[kk], bf[2], bf[3]);
      }
    }
    // ---- scale, causal mask, online softmax (rows g and g+8) ----
    const int64_t kbase = static_cast<int64_t>(kt) * kKTile;
    float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int j = 0; j < kKTile / 8; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int64_t key = kbase + j * 8 + 2 * t + (r & 1);
        const int64_t qp = qpos0 + (r >= 2 ? 8 : 0);
        const float v = (key <= qp) ? s[j][r] * scale_log2 : -INFINITY;
        s[j][r] = v;
        mx[r >> 1] = fmaxf(mx[r >> 1], v);
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      corr[r] = (mx[r] == -INFINITY) ? 1.f : exp2f(m_run[r] - mx[r]);
      m_run[r] = mx[r];
    }
#pragma unroll
    for (int j = 0; j < kKTile / 8; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float p = (s[j][r] == -INFINITY) ? 0.f : exp2f(s[j][r] - mx[r >> 1]);
        s[j][r] = p;
        rs[r >> 1] += p;
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int j = 0; j < D / 8; ++j) {
      o[j][0] *= corr[0]; o[j][1] *= corr[0];
      o[j][2] *= corr[1]; o[j][3] *= corr[1];
    }
    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < kKTile / 16; ++kk) {
      uint32_t pa[4];
      pa[0] = pack2<T>(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack2<T>(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack2<T>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack2<T>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int j = 0; j < D / 8; j += 2) {
        uint32_t bf[4];   // V^T fragments: {b0,b1} for dims j*8.., {b0,b1} for dims (j+1)*8..
        ldsm4_t(bf, vs + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + j * 8 + (lane >> 4) * 8);
        mma16816<T>(o[j], pa, bf[0], bf[1]);
        mma16816<T>(o[j + 1], pa, bf[2], bf[3]);
      }
    }
    __syncthreads();    // everyone is done with this stage before it is refilled
  }

  // ---- finalize: row sums across the quad, normalise, store ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int64_t qi = q0 + warp * 16 + g + r * 8;
    if (qi >= time) continue;
    const float inv = l_run[r] > 0.f ? 1.f / l_run[r] : 0.f;
    T* orow = out + ((b * time + qi) * H + h) * D;
#pragma unroll
    for (int j = 0; j < D / 8; ++j)
      *reinterpret_cast<uint32_t*>(orow + j * 8 + 2 * t) = pack2<T>(o[j][2 * r] * inv, o[j][2 * r + 1] * inv);
  }
}

template <typename T, int D>
void launch_mma(const void* qkv, const void* kc, const void* vc, int64_t batch, int64_t time, int64_t offset, int H,
                int Hkv, int64_t max_len, float scale, void* out, cudaStream_t st) {
  constexpr size_t smem = static_cast<size_t>(kQTile + 4 * kKTile) * (D + 8) * sizeof(T);
  auto kernel = attention_prefill_mma_kernel<T, D>;
  static bool configured = false;
  if (!configured) {
    CT2_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = true;
  }
  dim3 grid(div_up(time, kQTile), H, static_cast<unsigned>(batch));
  kernel<<<grid, kThreads, smem, st>>>(static_cast<const T*>(qkv), static_cast<const T*>(kc), static_cast<const T*>(vc),
                                       time, offset, H, Hkv, max_len, scale * 1.4426950408889634f, static_cast<T*>(out));
  check_launch();
}

}  // namespace

// fp16 / bf16, head_dim 64 or 128, no per-row lengths.  Returns false when the shape is not covered (caller falls
// back to the generic kernel in attention.cu).
bool launch_attention_prefill_mma(const void* qkv, const void* kc, const void* vc, int64_t batch, int64_t time,
                                  int64_t offset, int H, int Hkv, int D, int64_t max_len, float scale, void* out,
                                  int dtype, cudaStream_t st) {
  if (dtype == CT2B200_F32 || (D != 128 && D != 64)) return false;
  if (batch * time == 0) return true;
  if (dtype == CT2B200_F16) {
    if (D == 128) launch_mma<__half, 128>(qkv, kc, vc, batch, time, offset, H, Hkv, max_len, scale, out, st);
    else launch_mma<__half, 64>(qkv, kc, vc, batch, time, offset, H, Hkv, max_len, scale, out, st);
  } else {
    if (D == 128) launch_mma<__nv_bfloat16, 128>(qkv, kc, vc, batch, time, offset, H, Hkv, max_len, scale, out, st);
    else launch_mma<__nv_bfloat16, 64>(qkv, kc, vc, batch, time, offset, H, Hkv, max_len, scale, out, st);
  }
  return true;
}

void launch_attention_prefill(const void* qkv, const void* kc, const void* vc, const int32_t* lengths, int64_t batch,
                              int64_t time, int64_t offset, int H, int Hkv, int D, int64_t max_len, float scale,
                              void* out, int dtype, cudaStream_t st) {
  static const bool force_simple = [] {
    const char* e = std::getenv("CT2B200_ATTN_PREFILL");
    return e && std::string(e) == "simple";
  }();
  if (!force_simple && lengths == nullptr &&
      launch_attention_prefill_mma(qkv, kc, vc, batch, time, offset, H, Hkv, D, max_len, scale, out, dtype, st))
    return;
  launch_attention_prefill_simple(qkv, kc, vc, lengths, batch, time, offset, H, Hkv, D, max_len, scale, out, dtype, st);
}

}  // namespace ct2b200
