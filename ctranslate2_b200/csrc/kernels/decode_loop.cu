// decode_loop.cu — device side of GreedySearch::search (reference src/decoding.cc:844-971) so that the
// per-token loop needs no host round trip inside a CUDA graph: DisableTokens(end ids while step <
// min_length, decoding.cc:852-856) + BestSampler (TopK k=1, src/sampling.cc:25-32, lowest-index ties) +
// prefix forcing (update_sample_with_prefix, decoding.cc:21-67) + position bookkeeping.
#include <algorithm>

#include "../common.cuh"

namespace ct2b200 {

namespace {

// gen[0] = start_len (cache length when decoding started), gen[1] = min_length, gen[2] = num_end_ids,
// gen[3] = number of forced steps available in `forced`
template <typename T>
__global__ void __launch_bounds__(1024)
    sample_greedy_kernel(const T* __restrict__ logits, int64_t vocab, const int32_t* __restrict__ gen,
                         const int32_t* __restrict__ end_ids, const int32_t* __restrict__ forced, int64_t batch,
                         int32_t* __restrict__ next_ids, int32_t* __restrict__ out_ids, int32_t* __restrict__ lens) {
  __shared__ float sv[32];
  __shared__ int32_t si[32];
  const int64_t b = blockIdx.x;
  const T* row = logits + b * vocab;
  const int step = lens[b] - gen[0];
  const bool disable_end = step < gen[1];
  const int num_end = gen[2];
  float best = -INFINITY;
  int32_t besti = INT32_MAX;
  for (int64_t j = threadIdx.x; j < vocab; j += blockDim.x) {
    float v = to_f32(row[j]);
    if (disable_end) {
      for (int e = 0; e < num_end; ++e)
        if (end_ids[e] == j) v = -INFINITY;
    }
    if (v > best || (v == best && j < besti)) {
      best = v;
      besti = static_cast<int32_t>(j);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float v = __shfl_xor_sync(0xffffffffu, best, o);
    const int32_t i = __shfl_xor_sync(0xffffffffu, besti, o);
    if (v > best || (v == best && i < besti)) { best = v; besti = i; }
  }
  if (lane == 0) { sv[warp] = best; si[warp] = besti; }
  __syncthreads();
  if (warp == 0) {
    best = lane < nw ? sv[lane] : -INFINITY;
    besti = lane < nw ? si[lane] : INT32_MAX;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float v = __shfl_xor_sync(0xffffffffu, best, o);
      const int32_t i = __shfl_xor_sync(0xffffffffu, besti, o);
      if (v > best || (v == best && i < besti)) { best = v; besti = i; }
    }
    if (lane == 0) {
      out_ids[static_cast<int64_t>(step) * batch + b] = besti;
      int32_t nxt = besti;
      if (step + 1 < gen[3]) {            // a prompt token is still pending for this row: force it
        const int32_t f = forced[static_cast<int64_t>(step + 1) * batch + b];
        if (f >= 0) nxt = f;
      }
      next_ids[b] = nxt;
      lens[b] += 1;
    }
  }
}

template <typename T> __global__ void to_f32_kernel(const T* x, int64_t n, float* y) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    y[i] = to_f32(x[i]);
}
template <typename T> __global__ void from_f32_kernel(const float* x, int64_t n, T* y) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    y[i] = from_f32<T>(x[i]);
}
__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    p[i] = v;
}

}  // namespace

void launch_sample_greedy(const void* logits, int64_t batch, int64_t vocab, const int32_t* gen, const int32_t* end_ids,
                          const int32_t* forced, int32_t* next_ids, int32_t* out_ids, int32_t* lens, int dtype,
                          cudaStream_t st) {
  if (batch == 0) return;
  CT2_DISPATCH_DTYPE(dtype, (sample_greedy_kernel<T><<<batch, 1024, 0, st>>>(static_cast<const T*>(logits), vocab, gen,
                                                                           end_ids, forced, batch, next_ids,
                                                                           out_ids, lens)));
  check_launch();
}

void launch_convert_to_f32(const void* x, int64_t n, float* y, int dtype, cudaStream_t st) {
  if (n == 0) return;
  const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, 148 * 8));
  CT2_DISPATCH_DTYPE(dtype, (to_f32_kernel<T><<<blocks, 256, 0, st>>>(static_cast<const T*>(x), n, y)));
  check_launch();
}
void launch_convert_from_f32(const float* x, int64_t n, void* y, int dtype, cudaStream_t st) {
  if (n == 0) return;
  const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, 148 * 8));
  CT2_DISPATCH_DTYPE(dtype, (from_f32_kernel<T><<<blocks, 256, 0, st>>>(x, n, static_cast<T*>(y))));
  check_launch();
}
void launch_fill_i32(int32_t* p, int64_t n, int32_t v, cudaStream_t st) {
  if (n == 0) return;
  const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, 148 * 8));
  fill_i32_kernel<<<blocks, 256, 0, st>>>(p, n, v);
  check_launch();
}

}  // namespace ct2b200
