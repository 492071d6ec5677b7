// tp_rows.cu — the row kernels of the tensor-parallel decoder: the two collectives of a Llama layer fused into the
// kernels that consume them, over NVLink peer memory (CUDA IPC mappings of every rank's exchange buffer).
//
// Reference (src/layers/common.cc:348-401, transformer.cc:45-48, attention.cc:608-612, ops/nccl_ops_gpu.cu:52-85):
//   row-parallel Dense = GatherAll(activations) -> Quantize -> Gemm(K slice) -> Dequantize -> ReduceAll(sum), each
//   collective a host-blocking NCCL call.  Here:
//   * all-reduce(sum) of the out-proj / down-proj partials is the PROLOGUE of the next kernel on the residual stream
//     (RMSNorm + Quantize of the following sub-block): every rank leaves its partial [rows, d] in its own exchange
//     buffer (the GEMM simply writes there), raises a flag on every peer, and each rank pulls the `world` partials of
//     its rows over NVLink, sums them in rank order (bit-identical on every rank), adds the residual and goes on to
//     normalise / quantize the row it already holds in registers.  One kernel, no separate collective launch.
//   * the activation all-gather in front of Quantize only serves to find the per-row amax of the whole row: it is
//     replaced by an 8-byte {epoch, amax} exchange per row (flag and payload in one store, so no fence or second
//     round trip), fused into the quantization kernel.  The scales are bit-identical to the single-GPU ones.
// Flags and epochs: `tick` counts forward passes on this rank (device side, so the CUDA-graph replay stays valid);
// sync point `idx` of a pass waits for epoch tick * 1024 + idx + 1.  Two partial buffers / flag sets alternate
// (attention sub-block, FFN sub-block): a buffer is rewritten only after every peer has passed the other one.
#include "../common.cuh"
#include "kernels.h"

namespace ct2b200 {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// 16 bytes of a peer's partial: system-scope load, never served from a stale L1 line
template <typename T>
__device__ __forceinline__ Vec16<T> ld16_sys(const T* p) {
  Vec16<T> r;
  uint4 u;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p) : "memory");
  *reinterpret_cast<uint4*>(&r) = u;
  return r;
}

__global__ void tp_tick_kernel(uint32_t* tick) {
  griddep_launch();
  griddep_wait();
  if (threadIdx.x == 0) *tick = *tick + 1;
}

// MODE 0: x += sum(partials); q, scale = Quantize(RMSNorm(x, gamma))      (INT8 arm)
// MODE 1: x += sum(partials); y = RMSNorm(x, gamma) as T                  (float / AWQ arm)
// MODE 2: x += sum(partials)                                             (end of the layer stack)
template <typename T, int MODE, int NV>
__global__ void __launch_bounds__(kThreads)
    tp_reduce_rows_kernel(TpLink tp, int buf, int sync_idx, T* __restrict__ x, const T* __restrict__ gamma, int64_t cols,
                          float eps, int8_t* __restrict__ q, float* __restrict__ scale, T* __restrict__ y_out) {
  constexpr int N = Vec16<T>::N;
  __shared__ float red[32];
  griddep_launch();
  griddep_wait();                                   // our own partial (the preceding GEMM) is complete and visible
  const uint32_t epoch = *tp.tick * 1024u + static_cast<uint32_t>(sync_idx) + 1u;
  if (blockIdx.x == 0 && threadIdx.x < tp.world && static_cast<int>(threadIdx.x) != tp.rank) {
    __threadfence_system();
    st_release_sys(tp.flags_peer[threadIdx.x] + buf * 8 + tp.rank, epoch);
  }
  if (threadIdx.x < tp.world && static_cast<int>(threadIdx.x) != tp.rank) {
    const uint32_t* f = tp.flags_local + buf * 8 + threadIdx.x;
    while (ld_acquire_sys(f) != epoch) __nanosleep(20);
  }
  __syncthreads();

  const int64_t row = blockIdx.x;
  const int64_t nv = cols / N;
  T* xr = x + row * cols;
  float v[NV][N];
  bool have[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int64_t vi = threadIdx.x + static_cast<int64_t>(k) * kThreads;
    have[k] = vi < nv;
#pragma unroll
    for (int i = 0; i < N; ++i) v[k][i] = 0.f;
    if (!have[k]) continue;
    float acc[N];
#pragma unroll
    for (int i = 0; i < N; ++i) acc[i] = 0.f;
    for (int r = 0; r < tp.world; ++r) {            // rank order: every rank computes the same bits
      const T* pr = static_cast<const T*>(tp.parts[buf][r]) + row * cols + vi * N;
      const Vec16<T> d = r == tp.rank ? ld16(pr) : ld16_sys(pr);
#pragma unroll
      for (int i = 0; i < N; ++i) acc[i] += to_f32(d.v[i]);
    }
    const Vec16<T> xv = ld16(xr + vi * N);
    Vec16<T> o;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      v[k][i] = round_to<T>(round_to<T>(acc[i]) + to_f32(xv.v[i]));   // ReduceAll in T, then Add(residual) in T
      o.v[i] = from_f32<T>(v[k][i]);
    }
    st16(xr + vi * N, o);
  }
  if constexpr (MODE != 2) {

  Vec16<T> gv[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k)
    if (have[k]) gv[k] = ld16(gamma + (threadIdx.x + static_cast<int64_t>(k) * kThreads) * N);
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int i = 0; i < N; ++i) ss += v[k][i] * v[k][i];
  ss = block_reduce<false>(ss, red);
  const float inv = rsqrtf(ss / static_cast<float>(cols) + eps);
#pragma unroll
  for (int k = 0; k < NV; ++k)
    if (have[k]) {
#pragma unroll
      for (int i = 0; i < N; ++i) v[k][i] = round_to<T>(v[k][i] * inv * to_f32(gv[k].v[i]));
    }
  if constexpr (MODE == 1) {
#pragma unroll
    for (int k = 0; k < NV; ++k)
      if (have[k]) {
        Vec16<T> o;
#pragma unroll
        for (int i = 0; i < N; ++i) o.v[i] = from_f32<T>(v[k][i]);
        st16(y_out + row * cols + (threadIdx.x + static_cast<int64_t>(k) * kThreads) * N, o);
      }
  } else {
  float amax = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int i = 0; i < N; ++i) amax = fmaxf(amax, fabsf(v[k][i]));
  amax = block_reduce<true>(amax, red);
  const float s = amax != 0.f ? 127.f / amax : 1.f;
  int8_t* qr = q + row * cols;
#pragma unroll
  for (int k = 0; k < NV; ++k)
    if (have[k]) {
      int8_t out[N];
#pragma unroll
      for (int i = 0; i < N; ++i) out[i] = static_cast<int8_t>(nearbyintf(v[k][i] * s));
      int8_t* dst = qr + (threadIdx.x + static_cast<int64_t>(k) * kThreads) * N;
      if constexpr (N == 8) *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<uint2*>(out);
      else *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<uint32_t*>(out);
    }
  if (threadIdx.x == 0) scale[row] = s;
  }
  }
}

// Quantize of a row whose columns are split over the ranks (attention output by heads, FFN hidden by columns):
// q = rint(x * 127 / amax_of_the_whole_row), the amax exchanged as {epoch, amax} words.  slot 0 / 1 alternate.
template <typename T, int NV>
__global__ void __launch_bounds__(kThreads)
    tp_quantize_rows_kernel(TpLink tp, int slot, int sync_idx, const T* __restrict__ x, int64_t cols,
                            int8_t* __restrict__ q, float* __restrict__ scale) {
  constexpr int N = Vec16<T>::N;
  __shared__ float red[32];
  __shared__ float s_peer[8];
  griddep_launch();
  griddep_wait();
  const uint32_t epoch = *tp.tick * 1024u + static_cast<uint32_t>(sync_idx) + 1u;
  const int64_t row = blockIdx.x;
  const int64_t nv = cols / N;
  const T* xr = x + row * cols;
  float v[NV][N];
  bool have[NV];
  float amax = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int64_t vi = threadIdx.x + static_cast<int64_t>(k) * kThreads;
    have[k] = vi < nv;
#pragma unroll
    for (int i = 0; i < N; ++i) v[k][i] = 0.f;
    if (have[k]) {
      const Vec16<T> d = ld16(xr + vi * N);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        v[k][i] = to_f32(d.v[i]);
        amax = fmaxf(amax, fabsf(v[k][i]));
      }
    }
  }
  amax = block_reduce<true>(amax, red);
  if (threadIdx.x < tp.world) {
    const int peer = threadIdx.x;
    const unsigned long long word = (static_cast<unsigned long long>(epoch) << 32) | __float_as_uint(amax);
    if (peer != tp.rank) st_relaxed_sys64(tp.amax_peer[peer] + (static_cast<int64_t>(slot) * 8 + tp.rank) * tp.amax_rows + row, word);
    float other = amax;
    if (peer != tp.rank) {
      const unsigned long long* src = tp.amax_local + (static_cast<int64_t>(slot) * 8 + peer) * tp.amax_rows + row;
      unsigned long long w = ld_relaxed_sys64(src);
      while (static_cast<uint32_t>(w >> 32) != epoch) {
        __nanosleep(20);
        w = ld_relaxed_sys64(src);
      }
      other = __uint_as_float(static_cast<uint32_t>(w));
    }
    s_peer[peer] = other;
  }
  __syncthreads();
  float gmax = 0.f;
  for (int r = 0; r < tp.world; ++r) gmax = fmaxf(gmax, s_peer[r]);
  const float s = gmax != 0.f ? 127.f / gmax : 1.f;
  int8_t* qr = q + row * cols;
#pragma unroll
  for (int k = 0; k < NV; ++k)
    if (have[k]) {
      int8_t out[N];
#pragma unroll
      for (int i = 0; i < N; ++i) out[i] = static_cast<int8_t>(nearbyintf(v[k][i] * s));
      int8_t* dst = qr + (threadIdx.x + static_cast<int64_t>(k) * kThreads) * N;
      if constexpr (N == 8) *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<uint2*>(out);
      else *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<uint32_t*>(out);
    }
  if (threadIdx.x == 0) scale[row] = s;
}

template <typename T, int MODE>
void launch_reduce(const TpLink& tp, int buf, int sync_idx, T* x, const T* gamma, int64_t rows, int64_t cols, float eps,
                   int8_t* q, float* scale, T* y_out, cudaStream_t st) {
  constexpr int N = Vec16<T>::N;
  CT2_REQUIRE(cols % N == 0, "tensor parallel: d_model must be a multiple of 16 bytes");
  const int per = static_cast<int>((cols / N + kThreads - 1) / kThreads);
  if (per <= 1) launch_pdl(tp_reduce_rows_kernel<T, MODE, 1>, dim3(rows), dim3(kThreads), 0, st, tp, buf, sync_idx, x, gamma, cols, eps, q, scale, y_out);
  else if (per <= 2) launch_pdl(tp_reduce_rows_kernel<T, MODE, 2>, dim3(rows), dim3(kThreads), 0, st, tp, buf, sync_idx, x, gamma, cols, eps, q, scale, y_out);
  else if (per <= 4) launch_pdl(tp_reduce_rows_kernel<T, MODE, 4>, dim3(rows), dim3(kThreads), 0, st, tp, buf, sync_idx, x, gamma, cols, eps, q, scale, y_out);
  else if (per <= 8) launch_pdl(tp_reduce_rows_kernel<T, MODE, 8>, dim3(rows), dim3(kThreads), 0, st, tp, buf, sync_idx, x, gamma, cols, eps, q, scale, y_out);
  else throw InvalidArgument("tensor parallel: d_model too large for the fused reduce kernel");
  check_launch();
}

template <typename T>
void launch_tp_quantize(const TpLink& tp, int slot, int sync_idx, const T* x, int64_t rows, int64_t cols, int8_t* q,
                        float* scale, cudaStream_t st) {
  constexpr int N = Vec16<T>::N;
  CT2_REQUIRE(cols % N == 0, "tensor parallel: shard width must be a multiple of 16 bytes");
  CT2_REQUIRE(rows <= tp.amax_rows, "tensor parallel: more rows than the amax exchange holds");
  const int per = static_cast<int>((cols / N + kThreads - 1) / kThreads);
  if (per <= 1) launch_pdl(tp_quantize_rows_kernel<T, 1>, dim3(rows), dim3(kThreads), 0, st, tp, slot, sync_idx, x, cols, q, scale);
  else if (per <= 2) launch_pdl(tp_quantize_rows_kernel<T, 2>, dim3(rows), dim3(kThreads), 0, st, tp, slot, sync_idx, x, cols, q, scale);
  else if (per <= 4) launch_pdl(tp_quantize_rows_kernel<T, 4>, dim3(rows), dim3(kThreads), 0, st, tp, slot, sync_idx, x, cols, q, scale);
  else if (per <= 8) launch_pdl(tp_quantize_rows_kernel<T, 8>, dim3(rows), dim3(kThreads), 0, st, tp, slot, sync_idx, x, cols, q, scale);
  else throw InvalidArgument("tensor parallel: shard too wide for the fused quantize kernel");
  check_launch();
}

}  // namespace

void launch_tp_tick(uint32_t* tick, cudaStream_t st) {
  launch_pdl(tp_tick_kernel, dim3(1), dim3(32), 0, st, tick);
  check_launch();
}

void launch_tp_reduce_norm_quantize(const TpLink& tp, int buf, int sync_idx, void* x, const void* gamma, int64_t rows,
                                    int64_t cols, float eps, int8_t* q, float* scale, int dtype, cudaStream_t st) {
  if (rows == 0) return;
  CT2_DISPATCH_DTYPE(dtype, (launch_reduce<T, 0>(tp, buf, sync_idx, static_cast<T*>(x), static_cast<const T*>(gamma), rows,
                                                 cols, eps, q, scale, nullptr, st)));
}
void launch_tp_reduce_norm(const TpLink& tp, int buf, int sync_idx, void* x, const void* gamma, int64_t rows, int64_t cols,
                           float eps, void* y, int dtype, cudaStream_t st) {
  if (rows == 0) return;
  CT2_DISPATCH_DTYPE(dtype, (launch_reduce<T, 1>(tp, buf, sync_idx, static_cast<T*>(x), static_cast<const T*>(gamma), rows,
                                                 cols, eps, nullptr, nullptr, static_cast<T*>(y), st)));
}
void launch_tp_reduce(const TpLink& tp, int buf, int sync_idx, void* x, int64_t rows, int64_t cols, int dtype,
                      cudaStream_t st) {
  if (rows == 0) return;
  CT2_DISPATCH_DTYPE(dtype, (launch_reduce<T, 2>(tp, buf, sync_idx, static_cast<T*>(x), nullptr, rows, cols, 0.f, nullptr,
                                                 nullptr, nullptr, st)));
}
void launch_tp_quantize_rows(const TpLink& tp, int slot, int sync_idx, const void* x, int64_t rows, int64_t cols, int8_t* q,
                             float* scale, int dtype, cudaStream_t st) {
  if (rows == 0) return;
  CT2_DISPATCH_DTYPE(dtype, (launch_tp_quantize<T>(tp, slot, sync_idx, static_cast<const T*>(x), rows, cols, q, scale, st)));
}

}  // namespace ct2b200
