// tma_probe.cu — how fast can one CTA per SM pull a weight stream out of HBM, as a function of the shape of the TMA request?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tools/probes/tma_probe.cu -lcuda && ./tma_probe
// Every CTA streams its own contiguous-in-rows region through a ring of 16 KB shared-memory stages; a consumer warp only waits
// for a stage and hands it back (no math), so the number is the ingest rate of the copy engine + memory system:
//   mode 0  2-D tensor box 128 B x 128 rows of a row-major [rows, 4096 B] matrix (what gemm_decode / awq_decode issue today)
//   mode 1  the same box on a matrix whose row pitch IS 128 B (tile-blocked layout: the 16 KB are contiguous in HBM)
//   mode 2  one cp.async.bulk of 16 KB contiguous bytes per stage (tile-blocked layout, no tensor map)
//   mode 3  2-D tensor box 256 B x 64 rows (no swizzle) of the row-major matrix
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int kStageBytes = 16384;
constexpr int kStages = 12;            // 192 KB in flight per CTA

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}

__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ CUtensorMap tm, const uint8_t* base, int mode, int blocks_per_cta,
                                               int k_blocks) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full[kStages], empty[kStages];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    if (elect_one()) {
      for (int it = 0; it < blocks_per_cta; ++it) {
        const int s = it % kStages;
        if (it >= kStages) mbar_wait(empty + s, ((it / kStages) & 1) ^ 1);
        mbar_expect(full + s, kStageBytes);
        const uint32_t dst = smem_u32(smem + s * kStageBytes), bar = smem_u32(full + s);
        const int64_t blk = static_cast<int64_t>(blockIdx.x) * blocks_per_cta + it;       // global block index
        if (mode == 2) {
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst), "l"(base + blk * kStageBytes), "r"(kStageBytes), "r"(bar) : "memory");
        } else {
          int c0, c1;
          if (mode == 0) { c0 = static_cast<int>(blk % k_blocks) * 128; c1 = static_cast<int>(blk / k_blocks) * 128; }      // [rows, 4096 B]
          else if (mode == 1) { c0 = 0; c1 = static_cast<int>(blk) * 128; }                                                    // pitch 128 B
          else { c0 = static_cast<int>(blk % (k_blocks / 2)) * 256; c1 = static_cast<int>(blk / (k_blocks / 2)) * 64; }      // 256 B x 64 rows
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(dst), "l"(&tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
        }
      }
    }
  } else {
    if (elect_one()) {
      for (int it = 0; it < blocks_per_cta; ++it) {
        const int s = it % kStages;
        mbar_wait(full + s, (it / kStages) & 1);
        mbar_arrive(empty + s);
      }
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeFn encode = reinterpret_cast<EncodeFn>(fn);
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int k_bytes = 4096, k_blocks = k_bytes / 128;
  const int blocks_per_cta = 64 * k_blocks / 8;                 // 256 stages = 4 MB per CTA
  const int64_t total_blocks = static_cast<int64_t>(sms) * blocks_per_cta;
  const int64_t bytes = total_blocks * kStageBytes;
  uint8_t* buf;
  CK(cudaMalloc(&buf, bytes));
  CK(cudaMemset(buf, 1, bytes));
  const size_t smem = kStages * kStageBytes + 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  printf("%d SMs, %.2f GB per pass, %d x 16 KB stages in flight per CTA\n", sms, bytes / 1e9, kStages);
  for (int mode = 0; mode < 4; ++mode) {
    CUtensorMap tm;
    cuuint64_t dims[2], strides[1];
    cuuint32_t box[2], estr[2] = {1, 1};
    CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
    const int64_t rows = total_blocks / k_blocks * 128;
    if (mode == 0 || mode == 2) { dims[0] = k_bytes; dims[1] = rows; strides[0] = k_bytes; box[0] = 128; box[1] = 128; }
    if (mode == 1) { dims[0] = 128; dims[1] = total_blocks * 128; strides[0] = 128; box[0] = 128; box[1] = 128; }
    if (mode == 3) { dims[0] = k_bytes; dims[1] = rows; strides[0] = k_bytes; box[0] = 256; box[1] = 64; sw = CU_TENSOR_MAP_SWIZZLE_NONE; }
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("mode %d: encode failed %d\n", mode, static_cast<int>(r)); continue; }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      probe<<<sms, 64, smem>>>(tm, buf, mode, blocks_per_cta, k_blocks);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep == 2) printf("mode %d: %.3f ms, %.0f GB/s (%.1f GB/s per SM)\n", mode, ms, bytes / ms / 1e6, bytes / ms / 1e6 / sms);
    }
  }
  return 0;
}
