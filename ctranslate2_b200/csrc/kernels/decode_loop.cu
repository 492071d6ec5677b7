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
// Two-stage arg-max: grid (chunks, batch); every CTA scans one slice of the vocabulary with 16-byte loads and
// parks its (value, index) best; the last CTA of a row (ticket) reduces the slices and does the bookkeeping.
// Order is (value desc, index asc) everywhere => lowest-index ties, independent of the reduction tree.
constexpr int kSampleThreads = 256;

__device__ __forceinline__ void argmax_merge(float& bv, int32_t& bi, float v, int32_t i) {
  if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
}

// kScores: also the log-probability of the chosen token under LogSoftMax of the processed logits (return_scores,
// decoding.cc:875-880, 919-920): every chunk keeps sum exp(v - chunk max); the chunk max IS its argmax value, and the
// chosen token is the global argmax, so log p = -log(sum over chunks of s_c * exp(m_c - M)).
template <typename T, bool kScores>
__global__ void __launch_bounds__(kSampleThreads)
    sample_greedy_kernel(const T* __restrict__ logits, int64_t vocab, const int32_t* __restrict__ gen,
                         const int32_t* __restrict__ end_ids, const int32_t* __restrict__ forced, int64_t batch,
                         int32_t* __restrict__ next_ids, int32_t* __restrict__ out_ids, int32_t* __restrict__ lens,
                         float* __restrict__ part_v, int32_t* __restrict__ part_i, int32_t* __restrict__ tickets,
                         float* __restrict__ part_s, float* __restrict__ step_scores) {
  constexpr int N = Vec16<T>::N;
  __shared__ float sv[32];
  __shared__ int32_t si[32];
  __shared__ bool s_last;
  griddep_launch();
  griddep_wait();
  const int chunk = blockIdx.x, nchunks = gridDim.x;
  const int64_t b = blockIdx.y;
  const T* row = logits + b * vocab;
  const int step = lens[b] - gen[0];
  const bool disable_end = step < gen[1];
  const int num_end = gen[2];
  // slice [j0, j1) in units of N elements (the tail < N elements belongs to the last slice)
  const int64_t nvec = vocab / N;
  const int64_t v0 = chunk * nvec / nchunks, v1 = (chunk + 1) * nvec / nchunks;
  float best = -INFINITY;
  int32_t besti = INT32_MAX;
  float sum_m = -INFINITY, sum_s = 0.f;          // kScores: running max and sum exp(v - max) of this thread
  const bool vec_ok = (reinterpret_cast<uintptr_t>(row) & 15) == 0;
  auto consider = [&](float v, int64_t j) {
    if (disable_end)
      for (int e = 0; e < num_end; ++e)
        if (end_ids[e] == j) v = -INFINITY;
    argmax_merge(best, besti, v, static_cast<int32_t>(j));
    if constexpr (kScores) {
      if (v > sum_m) {
        sum_s = sum_s * __expf(sum_m - v) + 1.f;  // exp(-inf) = 0 on the first finite value
        sum_m = v;
      } else if (v != -INFINITY) {
        sum_s += __expf(v - sum_m);
      }
    }
  };
  if (vec_ok) {
    for (int64_t vi = v0 + threadIdx.x; vi < v1; vi += kSampleThreads) {
      const Vec16<T> d = ld16(row + vi * N);
#pragma unroll
      for (int i = 0; i < N; ++i) consider(to_f32(d.v[i]), vi * N + i);
    }
  } else {
    for (int64_t j = v0 * N + threadIdx.x; j < v1 * N; j += kSampleThreads) consider(to_f32(row[j]), j);
  }
  if (chunk == nchunks - 1)
    for (int64_t j = nvec * N + threadIdx.x; j < vocab; j += kSampleThreads) consider(to_f32(row[j]), j);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kSampleThreads >> 5;
  __shared__ float ss[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    argmax_merge(best, besti, __shfl_xor_sync(0xffffffffu, best, o), __shfl_xor_sync(0xffffffffu, besti, o));
  if constexpr (kScores) {                       // rescale every thread's sum to the warp max (= best after the merge)
    sum_s = sum_m == -INFINITY ? 0.f : sum_s * __expf(sum_m - best);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum_s += __shfl_xor_sync(0xffffffffu, sum_s, o);
  }
  if (lane == 0) { sv[warp] = best; si[warp] = besti; ss[warp] = sum_s; }
  __syncthreads();
  if (warp == 0) {
    const float wbest = lane < nw ? sv[lane] : -INFINITY;
    float wsum = lane < nw ? ss[lane] : 0.f;
    best = wbest;
    besti = lane < nw ? si[lane] : INT32_MAX;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      argmax_merge(best, besti, __shfl_xor_sync(0xffffffffu, best, o), __shfl_xor_sync(0xffffffffu, besti, o));
    if constexpr (kScores) {
      wsum = wbest == -INFINITY ? 0.f : wsum * __expf(wbest - best);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    }
    if (lane == 0) {
      part_v[b * nchunks + chunk] = best;
      part_i[b * nchunks + chunk] = besti;
      if constexpr (kScores) part_s[b * nchunks + chunk] = wsum;
      __threadfence();
      s_last = atomicAdd(tickets + b, 1) == nchunks - 1;
    }
  }
  __syncthreads();
  if (!s_last || warp != 0) return;
  __threadfence();
  best = -INFINITY;
  besti = INT32_MAX;
  for (int c = lane; c < nchunks; c += 32) argmax_merge(best, besti, __ldcg(part_v + b * nchunks + c), __ldcg(part_i + b * nchunks + c));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    argmax_merge(best, besti, __shfl_xor_sync(0xffffffffu, best, o), __shfl_xor_sync(0xffffffffu, besti, o));
  if constexpr (kScores) {
    float tot = 0.f;
    for (int c = lane; c < nchunks; c += 32) {
      const float mc = __ldcg(part_v + b * nchunks + c);
      tot += mc == -INFINITY ? 0.f : __ldcg(part_s + b * nchunks + c) * __expf(mc - best);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0) step_scores[static_cast<int64_t>(step) * batch + b] = -logf(tot);
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
    tickets[b] = 0;
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
__global__ void mul_inplace_f16_kernel(__half* a, const __half* b, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    a[i] = __hmul(a[i], b[i]);
}
__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    p[i] = v;
}

}  // namespace

int sample_greedy_chunks(int64_t vocab) { return vocab >= 32768 ? 32 : (vocab >= 4096 ? 8 : 1); }

// scratch: part_v float [batch*chunks], part_i int32 [batch*chunks], tickets int32 [batch] (zero between launches)
void launch_sample_greedy(const void* logits, int64_t batch, int64_t vocab, const int32_t* gen, const int32_t* end_ids,
                          const int32_t* forced, int32_t* next_ids, int32_t* out_ids, int32_t* lens, float* part_v,
                          int32_t* part_i, int32_t* tickets, float* part_s, float* step_scores, int dtype,
                          cudaStream_t st) {
  if (batch == 0) return;
  dim3 grid(sample_greedy_chunks(vocab), static_cast<unsigned>(batch));
  if (step_scores) {
    CT2_DISPATCH_DTYPE(dtype, (launch_pdl(sample_greedy_kernel<T, true>, grid, dim3(kSampleThreads), 0, st,
                                          static_cast<const T*>(logits), vocab, gen, end_ids, forced, batch, next_ids,
                                          out_ids, lens, part_v, part_i, tickets, part_s, step_scores)));
  } else {
    CT2_DISPATCH_DTYPE(dtype, (launch_pdl(sample_greedy_kernel<T, false>, grid, dim3(kSampleThreads), 0, st,
                                          static_cast<const T*>(logits), vocab, gen, end_ids, forced, batch, next_ids,
                                          out_ids, lens, part_v, part_i, tickets, part_s, step_scores)));
  }
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
void launch_mul_inplace_f16(void* a, const void* b, int64_t n, cudaStream_t st) {
  if (n == 0) return;
  const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, 148 * 8));
  mul_inplace_f16_kernel<<<blocks, 256, 0, st>>>(static_cast<__half*>(a), static_cast<const __half*>(b), n);
  check_launch();
}
template <typename T> __global__ void mul_inplace_kernel(T* a, const T* b, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    a[i] = from_f32<T>(to_f32(a[i]) * to_f32(b[i]));
}
// ops::Mul (gate * up) in T
void launch_mul_inplace(void* a, const void* b, int64_t n, int dtype, cudaStream_t st) {
  if (n == 0) return;
  const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, 148 * 8));
  CT2_DISPATCH_DTYPE(dtype, (mul_inplace_kernel<T><<<blocks, 256, 0, st>>>(static_cast<T*>(a), static_cast<const T*>(b), n)));
  check_launch();
}
void launch_fill_i32(int32_t* p, int64_t n, int32_t v, cudaStream_t st) {
  if (n == 0) return;
  const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, 148 * 8));
  fill_i32_kernel<<<blocks, 256, 0, st>>>(p, n, v);
  check_launch();
}

}  // namespace ct2b200
