// issue_loop_sass.cu -- compile-only study (nvcc -cubin, cuobjdump -sass): what the MMA issue loop of k_augru_pair
// should look like.  All 32 lanes of the MMA warp stay in uniform control flow, one lane elected ONCE with elect.sync
// issues; stage index, parity and operand offsets are compile-time constants.  Resulting SASS: UTCHMMA back to back,
// 2-3 instructions per MMA (an add + R2UR for a changed descriptor word), no ELECT / BRA.U.ANY wrapper -- against ~17
// instructions per MMA in the product kernel's `if (lane == 0)` form.  Not run on a GPU (budget); next round's first step.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -cubin -o /tmp/issue.cubin tools/experiments/issue_loop_sass.cu
//   cuobjdump -sass /tmp/issue.cubin | grep -c "ELECT"      -> 1
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .b32 r;\n\t.reg .pred p;\n\telect.sync r|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void mma2(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void commit2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ uint64_t desc_of(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// variant: whole warp in uniform control flow, elected lane issues
__global__ void __cluster_dims__(2,1,1) k(uint32_t tbase, int steps) {
  extern __shared__ uint8_t smem[];
  __shared__ uint64_t bar_full[6], bar_empty[6];
  const uint32_t a_lo = ((smem_u32(smem) >> 4) & 0x3fff) | (8u << 16), b_lo = ((smem_u32(smem + 65536) >> 4) & 0x3fff) | (8u << 16);
  if (threadIdx.x >= 32) return;
  const uint32_t leader = elect_one();
  for (int t = 0; t < steps; ++t) {
#pragma unroll
    for (int u = 0; u < 24; ++u) {
      const int stage = u % 6; const uint32_t par = (u / 6) & 1;
      mbar_wait(&bar_full[stage], par);
      if (leader) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint32_t bo = (stage * 16384 + j * 256) >> 4, ao = (((u % 8) * 2 + j) * 256) >> 4;
          const uint64_t dbh = desc_of(b_lo + bo, 0x4020), dbl = desc_of(b_lo + bo + 512, 0x4020);
          const uint64_t dah = desc_of(a_lo + ao, 0x4100), dal = desc_of(a_lo + ao + 2048, 0x4100);
          mma2(tbase + (u / 8) * 128, dah, dbh, 0x8400490, (u % 8 | j) ? 1u : 0u);
          mma2(tbase + (u / 8) * 128, dal, dbh, 0x8400490, 1u);
          mma2(tbase + (u / 8) * 128, dah, dbl, 0x8400490, 1u);
        }
        commit2(&bar_empty[stage]);
      }
      __syncwarp();
    }
  }
}
