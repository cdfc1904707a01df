// r4_augru_pair.cuh -- constants, PTX wrappers and weight image of the 2-CTA AUGRU kernel (r4_augru_pair2.cuh).
// Design notes of the first generation, kept because the layout facts still hold:
// the AUGRU recurrence (deepctr VecAttGRUCell, nets/utils.py:123-124) as ONE tensor-core
// instruction stream per 2-CTA cluster: tcgen05.mma.cta_group::2, M = 128 (64 feature rows per CTA), N = 256,
// bf16 hi/lo operands, fp32 accumulators in both CTAs' TMEM.  The arithmetic (r4_augru_tc.cuh):
//     r = sigmoid(Xr_t + h Wr)      u = sigmoid(Xu_t + h Wu)      c = tanh(Xc_t + (r*h) Wc)
//     u' = (1 - score_t) u ;  h <- u' h + (1 - u') c              (3 bf16 products per fp32 product)
// What the pair buys (measured with tools/pair_probe.cu): a 128x256x16 MMA takes 64 cycles instead of 128,
// each CTA streams only its half of the weight columns, each CTA's epilogue owns 64 rows instead of 128 (so
// a 4096-row batch x 2 sequences fills 128 SMs instead of 64), and the accumulators of one gate take 128 TMEM
// columns per CTA -- r, u and c no longer alias, and r*h gets its own A buffer, so the gate epilogues overlap
// the next gate's MMAs:
//     leader MMA thread : [r] ......... [u] ......... | wait r*h | [c] ......... | wait h' | ...
//     epilogue warps    :      wait r -> r*h  (|| u)   wait u -> 1+e^-u (|| c)     wait c -> h'
// TMEM layout of a pair MMA (cute tmem_frg_2sm "2x2" atom, confirmed by the probe): in each CTA, lanes 0-63 hold
// accumulator columns 0-127 of the CTA's 64 rows and lanes 64-127 hold columns 128-255.
//
// Per CTA: shared memory = A(h) hi/lo 64 KB + A(r*h) hi/lo 64 KB + 6-stage ring of 16 KB weight stages (one
// 32-deep K block of the CTA's 128 weight columns, hi then lo split, pre-tiled on the host so a stage is one bulk
// copy).  Both CTAs' stages must have landed before the leader issues: the peer's control warp relays its
// local "full" barrier to the leader with a remote mbarrier arrive (~460 cycles one way, hidden by the ring).
// Warp roles: 0-7 epilogue (TMEM lane quarter q = warp & 3, column half = warp >> 2; thread = one row x 64
// hidden columns), 8 = MMA issuer (leader) / ring relay (peer), 9 = TMA producer, 10-11 idle.
#pragma once
#include "r4_augru_tc.cuh"

namespace r4tc {

constexpr int P_RC = 64;                              // rows per CTA
constexpr int P_NB = 128;                             // weight columns (B rows) per CTA
constexpr int P_HALF_BYTES = P_NB * KB * 2;           // 8192: one split of one 32-deep K block of the CTA's 128 columns
constexpr int P_STAGE_BYTES = 2 * P_HALF_BYTES;       // 16384: ring stage = [hi | lo] of one K block (6 MMAs, 384 cycles)
constexpr int P_STAGES_PER_STEP = 3 * NKB;            // 24
// Ring depth.  Measured (tools/augru_probe.cu, round 2, cycles per step of k_augru_pair2 / k_augru_pp): 6 stages 17.1 k /
// 31.3 k, 4 stages 16.0 k / 31.4 k, 3 stages 15.7 k / 29.6 k, 2 stages 20.5 k.  A deep ring refills in bursts that compete
// with the gate epilogues for the shared-memory port (the MMA warp's own clocks: with 6 stages it waits 7-8 k cycles per
// step for operand quarters, i.e. for the epilogue, with 3 stages 2.7 k); 3 stages = 48 KB keeps the weight stream just
// ahead of the MMAs.  24 % (2 * P_NST) == 0 keeps stage and parity compile-time.
#ifndef R4P_NST
#define R4P_NST 3
#endif
constexpr int P_NST = R4P_NST;
#ifndef R4P_COMMIT_GROUP
#define R4P_COMMIT_GROUP 1
#endif
constexpr int P_CG = R4P_COMMIT_GROUP;                // ring stages released per tcgen05.commit (must divide P_NST and 24)
constexpr int P_A_BYTES = P_RC * HID * 2;             // 32768 per split
constexpr int P_SMEM_BYTES = 4 * P_A_BYTES + P_NST * P_STAGE_BYTES + 128;
constexpr int P_RANK_IMAGE_BYTES = P_STAGES_PER_STEP * P_STAGE_BYTES; // 393216 per CTA rank
constexpr int P_TC_R = 0, P_TC_U = 128, P_TC_C = 256;                  // TMEM column bases of the gates

__device__ __forceinline__ void mma2_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// completion of all earlier MMAs of this thread -> one arrival on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void commit2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// acquire.cluster try_wait is expensive (a producer loop built on it issued one stage per ~450 cycles); it is used
// only for the two per-step barriers the leader's MMA thread waits on, which collect remote relaxed arrivals.
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
               "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t laddr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(laddr), "r"(rank)); return r;
}
__device__ __forceinline__ void arrive_cl(uint32_t caddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(caddr) : "memory");
}
// A remote arrive with .release.cluster blocks its thread ~320 cycles (tools/pair_probe.cu); .relaxed costs ~5.  The
// relaxed form is enough where the data being published lives in the ARRIVING CTA's own shared memory and has
// already been made visible to the async proxy (TMA completion, or fence.proxy.async by every writer + __syncwarp):
// the consumer is this SM's tensor core, started by the leader only after it has observed the arrival.
#ifndef R4P_DIRECT_ARRIVE
#define R4P_DIRECT_ARRIVE 1   // direct (no relay) remote arrive: 0 = relaxed.cluster, 1 = release.cta (the default semantics of
#endif                        // mbarrier.arrive: CUTLASS's ClusterBarrier::arrive(cta_id)), 2 = release.cluster.  Measured
                              // (cycles per step, <RELAY 0, TMAP 0>): 15.5 k | 15.5 k | 30.1 k
__device__ __forceinline__ void arrive_remote(uint32_t caddr) {
#if R4P_DIRECT_ARRIVE == 1
  asm volatile("mbarrier.arrive.release.cta.shared::cluster.b64 _, [%0];" :: "r"(caddr) : "memory");
#elif R4P_DIRECT_ARRIVE == 2
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(caddr) : "memory");
#else
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" :: "r"(caddr) : "memory");
#endif
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// The kernel itself lives in r4_augru_pair2.cuh (k_augru_pair2); this header keeps the shared constants, the PTX
// wrappers of the 2-CTA instructions and the host-side weight image.

// Ring stage s8 (0..7) of every gate holds K block 0,4,1,5,2,6,3,7: the epilogue of k_augru_pair2 completes the blocks
// {q, 4 + q} of an A operand with its q-th column chunk, and the MMAs follow it in that order.
__host__ __device__ constexpr int pair_kb(int s8) { return (s8 & 1) * 4 + (s8 >> 1); }

// host: fp32 recurrent weights -> [rank 2][mat r,u,c][8 K blocks][hi, lo] stages of 128 columns x 32 K (8 KB).
// Wg: [256][512] rows = h index, columns [r | u];  Wc: [256][256].
inline void build_pair_image(const float* Wg, const float* Wc, uint8_t* img) {
  for (int rank = 0; rank < 2; ++rank)
    for (int mat = 0; mat < 3; ++mat)
      for (int kb = 0; kb < NKB; ++kb)
        for (int sp = 0; sp < 2; ++sp) {
          const int kbsrc = pair_kb(kb);               // the order in which the kernels walk the K blocks of a gate
          uint8_t* st = img + (size_t)rank * P_RANK_IMAGE_BYTES + (size_t)((mat * NKB + kb) * 2 + sp) * P_HALF_BYTES;
          for (int nl = 0; nl < P_NB; ++nl)
            for (int kk = 0; kk < KB; ++kk) {
              const int k = kbsrc * KB + kk, n = rank * P_NB + nl;
              float w = mat == 0 ? Wg[(size_t)k * 2 * HID + n] : (mat == 1 ? Wg[(size_t)k * 2 * HID + HID + n] : Wc[(size_t)k * HID + n]);
              uint16_t hi = host_bf16_bits(w);
              uint16_t v = sp == 0 ? hi : host_bf16_bits(w - host_bf16_val(hi));
              memcpy(st + (nl / 8) * B_SBO + (kk / 8) * LBO + (nl % 8) * 16 + (kk % 8) * 2, &v, 2);
            }
        }
}

}  // namespace r4tc
