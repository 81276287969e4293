// Descriptor helpers shared by the tcgen05 attention kernels (attention_tc.cu: T <= 256, one CTA per head;
// attention_tcs.cu: streamed key / query blocks for T >= 384).
#pragma once
#include "unet_kernels.cuh"

namespace kdip {

// idesc with optional MN-major operands (cute::UMMA::InstrDescriptor: bit 15 = a_major, bit 16 = b_major; 1 = MN-major)
__host__ __device__ inline uint32_t umma_idesc_bf16_major(int M, int N, int a_mn, int b_mn) {
  return umma_idesc_bf16(M, N) | ((uint32_t)(a_mn ? 1 : 0) << 15) | ((uint32_t)(b_mn ? 1 : 0) << 16);
}
// shared-memory descriptor with explicit leading / stride byte offsets (128B swizzle)
__device__ __forceinline__ uint64_t umma_desc_sw128_ls(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

}  // namespace kdip
