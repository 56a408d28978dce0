// oduck_policy_tc.cuh -- actor-MLP layers on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// One launch per Dense layer: Y[M x N] = act(X[M x K] . W[K x N] + b).  A CTA owns a 128 x NT output tile:
//   * operands are staged 32 k-columns at a time into shared memory in the canonical K-major, no-swizzle UMMA layout
//     (8 x 16-byte core matrices; element (r, k) at (r/8)*SBO + (k/4)*LBO + (r%8)*16 + (k%4)*4 bytes, LBO = 128, SBO = 1024),
//   * one elected thread issues 4 x tcgen05.mma.kind::tf32 (M = 128, N = NT, K = 8) per stage, fp32 accumulator in TMEM,
//   * tcgen05.commit -> mbarrier tells the CTA when the stage buffer may be refilled (two stages ping-pong),
//   * epilogue: tcgen05.ld (32 lanes x 32 columns per warp), bias + swish -> global, or (last layer) the whole
//     NormalTanh head per thread, since one thread then holds one env's 28 outputs.
// Observation normalisation is fused into the first layer's operand load.  fp32 storage everywhere; the MMA rounds the
// operands to tf32 (10-bit mantissa), accumulation is fp32.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TC_M 128
#define TC_KC 32                      // k-columns per stage (one 128-byte row segment)
#define TC_THREADS 128

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);            // start address
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;   // leading (K-direction core-matrix) byte offset
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;   // stride (M/N-direction 8-row group) byte offset
  d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
  return d;                                            // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}
// kind::tf32, D = f32, A/B = tf32, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

