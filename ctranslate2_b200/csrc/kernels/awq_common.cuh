// awq_common.cuh — the int4 -> fp16 conversion shared by the AWQ kernels (awq.cu, awq_decode.cu).
#pragma once

#include <cuda_fp16.h>

#include <cstdint>

namespace ct2b200 {

// the 8 channels of one native word -> 8 fp16 values in channel order: (q - z) * s
__device__ __forceinline__ uint4 awq_dequant_word(uint32_t w, __half2 z_bot, __half2 z_top, __half2 s2) {
  // bottom nibbles come out as 1024 + q, top nibbles as 1024 + 16 q (the reference's I4s_TO_F16s_MAGIC_NUM trick)
  constexpr uint32_t kLut = (0xf0 & 0xcc) | 0xaa, kBot = 0x000f000f, kTop = 0x00f000f0, kMagic = 0x64006400;
  const uint32_t t = w >> 8;
  uint32_t h0, h1, h2, h3;
  asm volatile("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(h0) : "r"(w), "n"(kBot), "n"(kMagic), "n"(kLut));
  asm volatile("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(h1) : "r"(w), "n"(kTop), "n"(kMagic), "n"(kLut));
  asm volatile("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(h2) : "r"(t), "n"(kBot), "n"(kMagic), "n"(kLut));
  asm volatile("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(h3) : "r"(t), "n"(kTop), "n"(kMagic), "n"(kLut));
  const __half2 sixteenth = __float2half2_rn(0.0625f);
  // z_bot = 1024 + z (exact), z_top = -(64 + z) (exact): both subtractions are exact in fp16, then one rounding
  __half2 v0 = __hmul2(__hsub2(*reinterpret_cast<__half2*>(&h0), z_bot), s2);
  __half2 v1 = __hmul2(__hfma2(*reinterpret_cast<__half2*>(&h1), sixteenth, z_top), s2);
  __half2 v2 = __hmul2(__hsub2(*reinterpret_cast<__half2*>(&h2), z_bot), s2);
  __half2 v3 = __hmul2(__hfma2(*reinterpret_cast<__half2*>(&h3), sixteenth, z_top), s2);
  uint4 r;
  r.x = *reinterpret_cast<uint32_t*>(&v0);
  r.y = *reinterpret_cast<uint32_t*>(&v1);
  r.z = *reinterpret_cast<uint32_t*>(&v2);
  r.w = *reinterpret_cast<uint32_t*>(&v3);
  return r;
}

// the same 8 channels WITHOUT the scale: exact integers q - z as fp16 (for kernels that apply the scale to a partial sum)
__device__ __forceinline__ uint4 awq_unscaled_word(uint32_t w, __half2 z_bot, __half2 z_top) {
  constexpr uint32_t kLut = (0xf0 & 0xcc) | 0xaa, kBot = 0x000f000f, kTop = 0x00f000f0, kMagic = 0x64006400;
  const uint32_t t = w >> 8;
  uint32_t h0, h1, h2, h3;
  asm volatile("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(h0) : "r"(w), "n"(kBot), "n"(kMagic), "n"(kLut));
  asm volatile("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(h1) : "r"(w), "n"(kTop), "n"(kMagic), "n"(kLut));
  asm volatile("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(h2) : "r"(t), "n"(kBot), "n"(kMagic), "n"(kLut));
  asm volatile("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(h3) : "r"(t), "n"(kTop), "n"(kMagic), "n"(kLut));
  const __half2 sixteenth = __float2half2_rn(0.0625f);
  __half2 v0 = __hsub2(*reinterpret_cast<__half2*>(&h0), z_bot);
  __half2 v1 = __hfma2(*reinterpret_cast<__half2*>(&h1), sixteenth, z_top);
  __half2 v2 = __hsub2(*reinterpret_cast<__half2*>(&h2), z_bot);
  __half2 v3 = __hfma2(*reinterpret_cast<__half2*>(&h3), sixteenth, z_top);
  uint4 r;
  r.x = *reinterpret_cast<uint32_t*>(&v0);
  r.y = *reinterpret_cast<uint32_t*>(&v1);
  r.z = *reinterpret_cast<uint32_t*>(&v2);
  r.w = *reinterpret_cast<uint32_t*>(&v3);
  return r;
}

}  // namespace ct2b200
