// r4_augru_pair2.cuh -- second generation of the 2-CTA AUGRU recurrence (deepctr VecAttGRUCell, nets/utils.py:123-124).
// Same arithmetic, TMEM / shared-memory layout and weight image as k_augru_pair (r4_augru_pair.cuh); what changed is
// everything around the tensor pipe, each item answering a measurement of round 1:
//   * issue loop: the whole MMA warp walks the step in uniform control flow, ONE lane elected once (elect.sync) issues,
//     ring stage / mbarrier parity / operand offsets are compile-time constants (a step = 24 ring uses = 4 revolutions
//     of the 6 stages, so they repeat every step).  Round 1: ~100 SASS instructions of descriptor arithmetic + an
//     ELECT/BRA.U.ANY wrapper per MMA paced the pair MMAs at ~105 cycles; this form issues at 52-65.
//   * rolling input prefetch: the epilogue keeps the inputs of the NEXT gate phase in registers (x[4][16]); a chunk is
//     reloaded with the next phase's columns the moment it has been consumed, so every L2 load has a whole phase
//     (~2 k cycles) to land.  Round 1: a 1-deep pipeline over 16-column chunks left each phase (R 5.3 k, U 4.8 k,
//     C 5.5 k cycles) bound by L2 latency, and the phases run back to back on the same 8 warps.
//   * hand-over: every epilogue warp of BOTH CTAs publishes a finished operand quarter with st.shared ->
//     fence.proxy.async -> __syncwarp -> one mbarrier.arrive.release.cta on the LEADER's barrier (a shared::cluster address
//     for the peer: the instruction CUTLASS's ClusterBarrier::arrive(cta_id) issues for the peer-epilogue -> leader-MMA
//     signals of its 2-SM kernels).  The data a peer warp publishes stays in the PEER's shared memory and is read by the
//     peer SM's half of the cta_group::2 MMA; what crosses the cluster is only the arrival.  Round 1 used a .relaxed remote
//     arrive (advisor finding: not a release); .release.cluster on the arriving warps costs 2x the kernel (30 k cycles per
//     step: ~320 cycles per arrive, 8 per step and warp, on the critical path of the chasing MMAs).  RELAY = 1 keeps the
//     strictly cluster-scoped chain as an option: the peer's warps arrive on a barrier in their OWN CTA and one relay lane
//     per operand (warps 10 / 11 of the peer) forwards each completed quarter with mbarrier.arrive.release.cluster --
//     19.3 k cycles per step against 15.5 k.
//   * weight ring by tensor-map TMA (TMAP = 1, optional; the per-CTA bulk-copy ring measured faster): both CTAs issue cp.async.bulk.tensor.2d.cta_group::2 for
//     their own half of the stage with the LEADER's "full" barrier as the completion target, the leader's producer
//     posts one expect_tx for both halves.  No relay thread, no remote arrive: the transaction count is the signal.
//     The image is viewed as a 2-D tensor of 1 KB rows, a stage is a 256 x 16 box of u32 (16 KB, dense).
#pragma once
#include <cuda.h>
#include "r4_augru_pair.cuh"

// Defaults of the product build (r4_capi.cu: augru_pair_impl 1), chosen by measurement (tools/augru_probe.cu, round 2, 64
// tiles, cycles per step): <RELAY 0, TMAP 0> 15.5 k | <0,1> 17.2 k | <1,1> 19.3 k.
#ifndef R4P2_RELAY
#define R4P2_RELAY 0
#endif
#ifndef R4P2_TMAP
#define R4P2_TMAP 0
#endif
#ifndef R4P2_PRESCALE
#define R4P2_PRESCALE 0   // 1: the cached input halves arrive pre-multiplied (r, u columns by -log2(e), c columns by 2 log2(e)),
#endif                    //    so a gate pre-activation in the exp2 domain is ONE fma(acc, scale, x') instead of add + mul
#ifndef R4P2_SWPIPE
#define R4P2_SWPIPE 0     // 1: the bf16 split + shared-memory store of chunk k-1 is issued next to the MUFU section of chunk k
#endif

namespace r4tc {

struct AugruPairParams {
  AugruTcParams b;
  CUtensorMap tmap[2];        // per sequence: the pair weight image as [rows of 1 KB][256 x u32] (TMAP variants only)
};

// Descriptor of a SWIZZLE_NONE K-major operand split into its two words: `lo` carries the start address (>> 4, 14 bits)
// and LBO, `hi` carries SBO and the version bit.  Advancing the operand by `bytes` is `lo + (bytes >> 4)` (shared
// memory is < 256 KB, the address field cannot carry out), so a descriptor costs ONE add in the issue loop.
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo) { return ((saddr >> 4) & 0x3fffu) | ((lbo >> 4) << 16); }
__device__ __forceinline__ constexpr uint32_t desc_hi(uint32_t sbo) { return ((sbo >> 4) & 0x3fffu) | (1u << 14); }
__device__ __forceinline__ uint64_t desc_of(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .b32 r;\n\t.reg .pred p;\n\telect.sync r|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred;
}

// this CTA's 16 KB box of the weight image -> own shared memory, completion bytes -> the barrier at `mbar_cluster`
// (a shared::cluster address: the leader's "full" barrier for both CTAs of the pair)
__device__ __forceinline__ void tma_box_cg2(uint32_t dst_smem, const CUtensorMap* tm, int c0, int c1, uint32_t mbar_cluster) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               :: "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(mbar_cluster) : "memory");
}

// the same, MULTICAST: one L2 read delivers the box to the same shared-memory offset of every CTA in `mask`; with
// cta_group::2 each destination's completion bytes go to the barrier of ITS pair's even CTA (the operand names the
// even-CTA position of the issuer's pair).  This is what lets several CTA pairs of a cluster share one weight stream.
__device__ __forceinline__ void tma_box_cg2_mc(uint32_t dst_smem, const CUtensorMap* tm, int c0, int c1, uint32_t mbar_cluster, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;"
               :: "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(mbar_cluster), "h"(mask) : "memory");
}
// tcgen05.commit -> one arrival on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void commit2_mask(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(smem_u32(bar)), "h"(mask) : "memory");
}

constexpr int P2_TM_ROW_BYTES = 1024;                             // tensor-map row = 256 x u32
constexpr int P2_TM_BOX_ROWS = P_STAGE_BYTES / P2_TM_ROW_BYTES;   // 16 rows per ring stage

// One gate of one step: 8 ring stages x 2 K16 slices x (A_hi*B_hi + A_lo*B_hi + A_hi*B_lo).  Every gate walks the K
// blocks of its A operand in the order pair_kb() = 0,4,1,5,2,6,3,7: the epilogue finishes blocks {q, 4 + q} with its
// q-th column chunk and publishes them on quarter barrier q, so the gate's MMAs chase the epilogue that produces their
// operand, 12 MMAs behind it (`qbar` = the 4 quarter barriers of the A operand, nullptr when it is already complete).
// build_pair_image lays the weights out in the same order.  No tcgen05 fence per stage: the weights come from TMA.
template <int GATE>
__device__ __forceinline__ void issue_gate2(uint32_t leader, uint32_t tbase, uint32_t aHi_lo, uint32_t aLo_lo, uint32_t b_lo,
                                            uint64_t* bar_full, uint64_t* bar_empty, uint64_t* qbar, uint32_t qpar,
                                            uint16_t empty_mask = 3, long long* wait_acc = nullptr) {
  constexpr uint32_t idesc = make_idesc(TM, HID);
  constexpr uint32_t dcol = GATE == 0 ? P_TC_R : (GATE == 1 ? P_TC_U : P_TC_C);
  constexpr uint32_t a_hi = desc_hi(A_SBO), b_hi = desc_hi(B_SBO);
#pragma unroll
  for (int s8 = 0; s8 < NKB; ++s8) {
    const int u = GATE * NKB + s8;
    const int stage = u % P_NST;
    const uint32_t par = (uint32_t)((u / P_NST) & 1);
    const int kb = pair_kb(s8);
    // wait_acc (probe builds only): [0] cycles waiting for operand quarters, [1] cycles waiting for ring stages
    long long w0 = wait_acc ? clock64() : 0;
    if (qbar != nullptr && (s8 & 1) == 0) { mbar_wait_cl(&qbar[s8 >> 1], qpar); tc_fence_after(); }
    long long w1 = wait_acc ? clock64() : 0;
    mbar_wait(&bar_full[stage], par);
    if (wait_acc) { wait_acc[0] += w1 - w0; wait_acc[1] += clock64() - w1; }
    if (leader) {
#pragma unroll
      for (int j = 0; j < KB / 16; ++j) {
        const uint32_t bo = (uint32_t)(stage * P_STAGE_BYTES + j * 2 * LBO) >> 4;
        const uint32_t ao = (uint32_t)((kb * (KB / 16) + j) * 2 * LBO) >> 4;
        const uint64_t dbh = desc_of(b_lo + bo, b_hi), dbl = desc_of(b_lo + bo + (P_HALF_BYTES >> 4), b_hi);
        const uint64_t dah = desc_of(aHi_lo + ao, a_hi), dal = desc_of(aLo_lo + ao, a_hi);
        mma2_bf16(tbase + dcol, dah, dbh, idesc, (s8 | j) ? 1u : 0u);
        mma2_bf16(tbase + dcol, dal, dbh, idesc, 1u);
        mma2_bf16(tbase + dcol, dah, dbl, idesc, 1u);
      }
      commit2_mask(&bar_empty[stage], empty_mask);
    }
    __syncwarp();
  }
}

#ifndef R4_ABL
#define R4_ABL 0          // ABLATION PROBES ONLY (results become wrong): 1 no bf16 split + st.shared, 2 no tcgen05.ld/st in the
#endif                    // epilogue, 4 no global input loads, 8 no MUFU -- which part of a gate phase costs what (tools/augru_probe.cu)
#ifndef R4_SOFTRCP
#define R4_SOFTRCP 0      // reciprocals of the gate epilogues on the FMA pipe: 0 none (all MUFU.RCP), 1 all, 2 every second element
#endif
// 1 / x for 1 <= x < 2^122 without the MUFU pipe: exponent-flip seed (5 % error), one cubic step e + e^2 and one Newton
// step -- 5 FMAs + 1 integer subtract, relative error <= 1.2e-7 (1 ulp class, like rcp.approx).  The gate epilogues need
// 5 MUFU operations per state element (3 ex2 + 2 rcp) and are bound by that pipe (16 results / clk / SM); moving the
// reciprocals to the FMA pipe (128 lanes / clk / SM) trades 1 MUFU slot for 6 issue slots.
__device__ __forceinline__ float soft_rcp(float x) {
  float y = __int_as_float(0x7EF311C7 - __float_as_int(x));
  float e = fmaf(-x, y, 1.0f);
  y = fmaf(y, fmaf(e, e, e), y);
  e = fmaf(-x, y, 1.0f);
  return fmaf(y, e, y);
}
// reciprocal of element j of a chunk: the pipe is a compile-time choice per element
__device__ __forceinline__ float rcp_sel(float x, int j) {
  return (R4_SOFTRCP == 1 || (R4_SOFTRCP == 2 && (j & 1))) ? soft_rcp(x) : rcp_approx(x);
}
constexpr float P2_NL2E = -1.4426950408889634f, P2_2L2E = 2.8853900817779268f;
// gate pre-activation in the exp2 domain: scale * (acc + x); with R4P2_PRESCALE x already carries the scale
__device__ __forceinline__ float preact2(float acc, float x, float scale) {
#if R4P2_PRESCALE
  return fmaf(acc, scale, x);
#else
  return scale * (acc + x);
#endif
}
// 8 fp32 -> bf16 hi / lo core-matrix rows of an A operand
__device__ __forceinline__ void split_store8(const float* v, uint8_t* hi_base, uint8_t* lo_base, uint32_t off) {
#if R4_ABL & 1
  if (v[0] == 1234.5678f) *reinterpret_cast<float*>(hi_base + off) = v[1];     // keeps the values live, never taken
#else
  uint4 hi, lo;
  split8(v, hi, lo);
  *reinterpret_cast<uint4*>(hi_base + off) = hi;
  *reinterpret_cast<uint4*>(lo_base + off) = lo;
#endif
}
#if R4_ABL & 2
#define P2_TMEM_LD16(addr, dst) do { _Pragma("unroll") for (int q_ = 0; q_ < 16; ++q_) (dst)[q_] = 0.25f; } while (0)
#define P2_TMEM_ST16(addr, src) do { if ((src)[0] == 1234.5678f) asm volatile("" :: "f"((src)[1])); } while (0)
#define P2_TMEM_WAIT_LD() do { } while (0)
#define P2_TMEM_WAIT_ST() do { } while (0)
#else
#define P2_TMEM_LD16(addr, dst) tmem_ld16((addr), (dst))
#define P2_TMEM_ST16(addr, src) tmem_st16((addr), (src))
#define P2_TMEM_WAIT_LD() tmem_wait_ld()
#define P2_TMEM_WAIT_ST() tmem_wait_st()
#endif
#if R4_ABL & 8
__device__ __forceinline__ float p2_ex2(float x) { return x * 0.001f + 1.0f; }
__device__ __forceinline__ float p2_rcp(float x, int) { return 2.0f - x * 0.5f; }
#else
__device__ __forceinline__ float p2_ex2(float x) { return ex2_approx(x); }
__device__ __forceinline__ float p2_rcp(float x, int j) { return rcp_sel(x, j); }
#endif

// CS = CTAs per cluster (2, 4 or 8; set at launch with cudaLaunchAttributeClusterDimension): CS / 2 CTA pairs working on
// consecutive row tiles of the SAME sequence share ONE weight stream -- pair 0's producers multicast every ring stage to
// all pairs.  Why: chip-wide the recurrence is bound by L2 throughput, not by the tensor pipe (ablation probes, round 2:
// 128 CTAs x (393 KB weights + 196 KB inputs) per step = 75 MB against ~6300 B/clk of L2 -> 12 k cycles; without the
// input loads 13.6 k, without any epilogue work 11.4 k of the 17.3 k-cycle step).  Sharing the stream between 2 (4)
// pairs cuts the weight traffic to 1/2 (1/4).  Requires TMAP.
template <int RELAY, int TMAP, int CS = 2>
__global__ void __launch_bounds__(NTHREADS, 1) k_augru_pair2(const __grid_constant__ AugruPairParams pp) {
  static_assert(CS == 2 || TMAP, "weight-stream sharing needs the tensor-map ring");
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_full[P_NST], bar_empty[P_NST], bar_h[4], bar_rh[4], bar_r, bar_u, bar_c;
  __shared__ uint32_t tmem_base_s;
  const AugruTcParams& p = pp.b;
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint8_t* sHhi = smem;                       // A operand: h
  uint8_t* sHlo = smem + P_A_BYTES;
  uint8_t* sRhi = smem + 2 * P_A_BYTES;       // A operand: r*h
  uint8_t* sRlo = smem + 3 * P_A_BYTES;
  uint8_t* sB = smem + 4 * P_A_BYTES;
  const AugruTcSeq& S = p.s[blockIdx.y];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = cluster_ctarank();    // rank in the cluster: pair = crank >> 1
  const uint32_t rank = crank & 1;             // rank inside the CTA pair: 0 = leader (issues the MMAs)
  const uint32_t lead = crank & ~1u;           // cluster rank of this pair's leader
  const int n_tiles = (p.R + TM - 1) / TM;
  const int tile = min((int)(blockIdx.x >> 1), n_tiles - 1);   // padding pairs of the last cluster redo the last tile, unseen
  const bool pad_pair = (int)(blockIdx.x >> 1) >= n_tiles;
  const int m0 = tile * TM;                    // the pair's 128-row tile
  constexpr uint16_t ALL_MASK = (uint16_t)((1u << CS) - 1);
  const uint16_t pair_mask = (uint16_t)(3u << lead);

  if (tid == 0) {
    // "full": TMAP -> only the leader's barrier is used: one expect_tx arrival covering both CTAs' bytes;
    //         else the leader's collects its own TMA (expect_tx arrival) and the peer's relay arrival.
    for (int i = 0; i < P_NST; ++i) { mbar_init(&bar_full[i], (rank == 0 && !TMAP) ? 2 : 1); mbar_init(&bar_empty[i], CS / 2); }
    // hand-over barriers: RELAY -> 8 local warps (+ 1 relay arrival on the leader); else 8 warps x 2 CTAs on the leader
    const int nh = RELAY ? (rank == 0 ? 9 : 8) : 16;
    for (int i = 0; i < 4; ++i) { mbar_init(&bar_h[i], nh); mbar_init(&bar_rh[i], nh); }
    mbar_init(&bar_r, 1); mbar_init(&bar_u, 1); mbar_init(&bar_c, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                         // the peer's barriers exist before anyone arrives remotely
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  const uint32_t bar_h_leader = mapa_rank(smem_u32(&bar_h[0]), lead), bar_rh_leader = mapa_rank(smem_u32(&bar_rh[0]), lead);

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");   // frees 128 x 128 registers = what 232 for the 256 epilogue threads takes
    if (warp == 9) {
      // ===== TMA producer: this CTA's half of the 24-stage weight stream of a step, 64 times =====
      if (lane == 0) {
        int stage = 0; uint32_t phase = 0;
        if (TMAP) {
          const CUtensorMap* tm = &pp.tmap[blockIdx.y];
          const uint32_t full0_leader = mapa_rank(smem_u32(&bar_full[0]), lead);
          const uint32_t sB_u = smem_u32(sB);
          // CS > 2: the two CTAs of pair 0 each multicast THEIR column half to the same-parity CTA of every pair
          uint16_t mc_mask = 0;
          for (int c = (int)rank; c < CS; c += 2) mc_mask |= (uint16_t)(1u << c);
          for (int t = 0; t < STEPS; ++t) {
            int row = (int)rank * P_STAGES_PER_STEP * P2_TM_BOX_ROWS;
            for (int i = 0; i < P_STAGES_PER_STEP; ++i, row += P2_TM_BOX_ROWS) {
              mbar_wait(&bar_empty[stage], phase ^ 1);          // every pair of the cluster has released the stage
              if (rank == 0) mbar_expect_tx(&bar_full[stage], 2 * P_STAGE_BYTES);
              if (CS == 2) tma_box_cg2(sB_u + stage * P_STAGE_BYTES, tm, 0, row, full0_leader + stage * 8);
              else if (crank < 2) tma_box_cg2_mc(sB_u + stage * P_STAGE_BYTES, tm, 0, row, full0_leader + stage * 8, mc_mask);
              if (++stage == P_NST) { stage = 0; phase ^= 1; }
            }
          }
        } else {
          const uint8_t* img = S.Wimg + (size_t)rank * P_RANK_IMAGE_BYTES;
          for (int t = 0; t < STEPS; ++t) {
            const uint8_t* src = img;
            for (int i = 0; i < P_STAGES_PER_STEP; ++i, src += P_STAGE_BYTES) {
              mbar_wait(&bar_empty[stage], phase ^ 1);
              mbar_expect_tx(&bar_full[stage], P_STAGE_BYTES);
              bulk_g2s(sB + stage * P_STAGE_BYTES, src, P_STAGE_BYTES, &bar_full[stage]);
              if (++stage == P_NST) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    } else if (warp == 8 && rank == 1) {
      // ===== ring relay (only without the tensor map): tell the leader that this CTA's copy of stage s has landed =====
      if (!TMAP && lane == 0) {
        const uint32_t remote0 = mapa_rank(smem_u32(&bar_full[0]), lead);
        int stage = 0; uint32_t phase = 0;
        for (int i = 0; i < STEPS * P_STAGES_PER_STEP; ++i) {
          mbar_wait(&bar_full[stage], phase);
          arrive_remote(remote0 + stage * 8);
          if (++stage == P_NST) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 8) {
      // ===== MMA issuer (leader CTA): uniform control flow, one elected lane issues =====
      const uint32_t leader = elect_one();
      uint32_t hHi_d = desc_lo(smem_u32(sHhi), LBO), hLo_d = desc_lo(smem_u32(sHlo), LBO), rHi_d = desc_lo(smem_u32(sRhi), LBO),
               rLo_d = desc_lo(smem_u32(sRlo), LBO), b_d = desc_lo(smem_u32(sB), LBO);
      for (int t = 0; t < STEPS; ++t) {
        const uint32_t par = t & 1;
        long long* dbg = (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && leader) ? p.dbg + t * 16 : nullptr;
        if (dbg) dbg[0] = clock64();
        // keep the 144 descriptors of a step OUT of the loop-invariant set: hoisted, they live in local memory (this warp
        // has 40 registers) and every MMA pays a local load; rebuilt from these five words each costs one add
        asm volatile("" : "+r"(hHi_d), "+r"(hLo_d), "+r"(rHi_d), "+r"(rLo_d), "+r"(b_d));
        long long wacc[2] = {0, 0};
        long long* wa = dbg ? wacc : nullptr;
        if (dbg) dbg[1] = clock64();
        issue_gate2<0>(leader, tbase, hHi_d, hLo_d, b_d, bar_full, bar_empty, bar_h, par, ALL_MASK, wa);   // chases phase C of step t-1
        if (leader) commit2_mask(&bar_r, pair_mask);
        if (dbg) dbg[2] = clock64();
        issue_gate2<1>(leader, tbase, hHi_d, hLo_d, b_d, bar_full, bar_empty, nullptr, 0, ALL_MASK, wa);
        if (leader) commit2_mask(&bar_u, pair_mask);
        if (dbg) dbg[3] = clock64();
        if (dbg) dbg[4] = clock64();
        issue_gate2<2>(leader, tbase, rHi_d, rLo_d, b_d, bar_full, bar_empty, bar_rh, par, ALL_MASK, wa);  // chases phase R of this step
        if (leader) commit2_mask(&bar_c, pair_mask);
        if (dbg) { dbg[5] = clock64(); dbg[6] = wacc[0]; dbg[7] = wacc[1]; }
      }
    } else if (RELAY && rank == 1 && lane == 0) {
      // ===== hand-over relays (peer CTA): forward each completed local phase with a cluster-scope release =====
      // A barrier cannot complete its next phase before this one has been forwarded: that needs the leader's next gate on
      // the same operand, which waits for this very arrival.  warp 11: the quarters of h; warp 10: the quarters of r*h.
      if (warp == 11) {
        for (int k = 0; k <= STEPS; ++k)
          for (int i = 0; i < 4; ++i) { mbar_wait(&bar_h[i], k & 1); arrive_cl(bar_h_leader + i * 8); }
      } else {
        for (int t = 0; t < STEPS; ++t)
          for (int i = 0; i < 4; ++i) { mbar_wait(&bar_rh[i], t & 1); arrive_cl(bar_rh_leader + i * 8); }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ===== epilogue warps: thread = (row of this CTA, 4 chunks of 16 hidden columns) =====
    // Chunk ch of a thread covers hidden columns half * 128 + ch * 32 + sub * 16 .. + 15: the two warps (sub = 0, 1) of a
    // lane quarter together finish the 32-wide K block half * 4 + ch with their ch-th chunk, so over the CTA pair chunk ch
    // completes K blocks {ch, 4 + ch} of the A operand -- quarter barrier ch, the unit the MMA warp chases.
    const int q = warp & 3, sub = warp >> 2;
    const int rl = (q & 1) * 32 + lane;                    // row inside this CTA
    const int prow = (int)rank * P_RC + rl;                // row inside the pair's 128-row tile
    const int hcb = (q >> 1) * 128 + sub * 16;             // hidden column of chunk ch: hcb + ch * 32
    const uint32_t tcb = (uint32_t)sub * 16;               // TMEM column inside a gate of chunk ch: tcb + ch * 32
    int r = m0 + prow;
    const bool valid = r < p.R && !pad_pair;
    if (r >= p.R) r = p.R - 1;
    const int ci = S.shared ? 0 : (p.row0 + r) / p.div;
    const float* xt = S.XT + ((size_t)(ci / TM) * STEPS) * XT_COLS * TM;
    const int ln4 = (ci % TM) * 4;
    const float* st = S.scoresT + ((size_t)(m0 / TM) * STEPS) * TM + prow;
    const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
    const uint32_t a_row_off = (uint32_t)(rl / 8) * A_SBO + (uint32_t)(rl % 8) * 16 + (uint32_t)(hcb / 8) * LBO;
    const bool local_arrive = RELAY || rank == 0;          // arrive on this CTA's own barrier (else: relaxed, on the leader's)
    // publish quarter `i` of an A operand: this thread's stores -> async proxy, then one arrival per warp
    auto publish = [&](uint64_t* bars, uint32_t bars_leader, int i) {
      tc_fence_before();
      proxy_fence();
      __syncwarp();
      if (lane == 0) { if (local_arrive) mbar_arrive(&bars[i]); else arrive_remote(bars_leader + i * 8); }
    };
    float h[64], x[4][16];
#pragma unroll
    for (int i = 0; i < 64; ++i) h[i] = 0.f;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {                       // h0 = 0 into the A operand
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const uint32_t off = a_row_off + (uint32_t)(ch * 4 + g) * LBO;
        *reinterpret_cast<uint4*>(sHhi + off) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(sHlo + off) = make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) publish(bar_h, bar_h_leader, i);
#if R4_ABL & 4
#define R4P2_LOADX(dst, base, colbase) do { _Pragma("unroll") for (int q_ = 0; q_ < 16; ++q_) (dst)[q_] = 0.125f * (float)(ln4 & 3); } while (0)
#else
#define R4P2_LOADX(dst, base, colbase) load_x16(dst, (base), (colbase), ln4)
#endif
    // rolling input buffer: x[ch] always holds chunk ch of the NEXT phase to run (here: the r gate of step 0)
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) { R4P2_LOADX(x[c4], xt, hcb + c4 * 32); }
    float one_minus_s = 1.0f - __ldg(st);

    for (int t = 0; t < STEPS; ++t) {
      const uint32_t par = t & 1;
      long long* dbg = (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) ? p.dbg + t * 16 : nullptr;
      const float* xs = xt + (size_t)t * XT_COLS * TM;
      const bool more = t + 1 < STEPS;
      const float* xs_next = xs + (more ? (size_t)XT_COLS * TM : 0);
      const float oms_next = 1.0f - __ldg(st + (size_t)(more ? t + 1 : t) * TM);
      // Pull this CTA's half of the NEXT step's input lines (768 columns x 2 lines) from HBM into L2.
      if (more) {
        const float* xn = S.XT + (((size_t)(ci / TM) * STEPS + (t + 1)) * XT_COLS) * TM;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int id = i * 256 + tid;                    // 1536 lines: (column quad, 8 lines of this CTA's 64 lanes x 16 B)
          asm volatile("prefetch.global.L2 [%0];" :: "l"(xn + (size_t)(id >> 3) * 4 * TM + rank * (P_RC * 4) + (id & 7) * 32));
        }
      }
      // ---- phase R (overlaps the u MMAs; the c MMAs chase it quarter by quarter): r*h -> its own A operand;
      //      x[] <- the u gate's inputs ----
      {
        float a[2][16];
        if (dbg) dbg[8] = clock64();
        mbar_wait(&bar_r, par);
        if (dbg) dbg[9] = clock64();
        tc_fence_after();
        P2_TMEM_LD16(tlane + P_TC_R + tcb, a[0]);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int cur = ch & 1, nxt = cur ^ 1;
          P2_TMEM_WAIT_LD();
          if (ch < 3) P2_TMEM_LD16(tlane + P_TC_R + tcb + (ch + 1) * 32, a[nxt]);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            a[cur][j] = p2_rcp(1.0f + p2_ex2(fminf(preact2(a[cur][j], x[ch][j], P2_NL2E), 60.0f)), j) * h[ch * 16 + j];
          R4P2_LOADX(x[ch], xs, HID + hcb + ch * 32);
#pragma unroll
          for (int g = 0; g < 2; ++g)
            split_store8(a[cur] + g * 8, sRhi, sRlo, a_row_off + (uint32_t)(ch * 4 + g) * LBO);
          publish(bar_rh, bar_rh_leader, ch);
        }
      }
      if (dbg) dbg[10] = clock64();
      // ---- phase U (overlaps the c MMAs): E = 1 + exp(-(acc_u + Xu)) back into TMEM; x[] <- the c gate's inputs ----
      {
        float a[2][16];
        mbar_wait(&bar_u, par);
        if (dbg) dbg[11] = clock64();
        tc_fence_after();
        P2_TMEM_LD16(tlane + P_TC_U + tcb, a[0]);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int cur = ch & 1, nxt = cur ^ 1;
          P2_TMEM_WAIT_LD();
          if (ch < 3) P2_TMEM_LD16(tlane + P_TC_U + tcb + (ch + 1) * 32, a[nxt]);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            a[cur][j] = 1.0f + p2_ex2(fminf(preact2(a[cur][j], x[ch][j], P2_NL2E), 60.0f));
          R4P2_LOADX(x[ch], xs, 2 * HID + hcb + ch * 32);
          P2_TMEM_ST16(tlane + P_TC_U + tcb + ch * 32, a[cur]);
        }
        P2_TMEM_WAIT_ST();
      }
      if (dbg) dbg[12] = clock64();
      // ---- phase C (the next step's r MMAs chase it quarter by quarter): c = tanh(acc_c + Xc) = 1 - 2/(1 + F), u = 1/E
      //      with ONE reciprocal of E*F; x[] <- next step's r inputs ----
      {
        float a[2][16], u[2][16];
        mbar_wait(&bar_c, par);
        if (dbg) dbg[13] = clock64();
        tc_fence_after();
        P2_TMEM_LD16(tlane + P_TC_C + tcb, a[0]);
        P2_TMEM_LD16(tlane + P_TC_U + tcb, u[0]);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int cur = ch & 1, nxt = cur ^ 1;
          P2_TMEM_WAIT_LD();
          if (ch < 3) {
            P2_TMEM_LD16(tlane + P_TC_C + tcb + (ch + 1) * 32, a[nxt]);
            P2_TMEM_LD16(tlane + P_TC_U + tcb + (ch + 1) * 32, u[nxt]);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float F = 1.0f + p2_ex2(fminf(preact2(a[cur][j], x[ch][j], P2_2L2E), 60.0f));
            const float E = u[cur][j];
            const float rc = p2_rcp(E * F, j);                      // E, F <= 1 + 2^60: the product is finite
            const float c = fmaf(-2.0f, rc * E, 1.0f);               // tanh
            const float up = one_minus_s * (rc * F);                 // (1 - s) sigmoid
            const float hn = fmaf(up, h[ch * 16 + j] - c, c);        // u' h + (1 - u') c
            h[ch * 16 + j] = hn;
            a[cur][j] = hn;
          }
          R4P2_LOADX(x[ch], xs_next, hcb + ch * 32);      // (last step: a harmless re-read of this step's lines)
#pragma unroll
          for (int g = 0; g < 2; ++g)
            split_store8(a[cur] + g * 8, sHhi, sHlo, a_row_off + (uint32_t)(ch * 4 + g) * LBO);
          publish(bar_h, bar_h_leader, ch);
        }
      }
      if (dbg) dbg[14] = clock64();
      one_minus_s = oms_next;
    }
#undef R4P2_LOADX
    if (valid) {
      float* o = S.out + (size_t)(m0 + prow) * p.out_ld + hcb;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(o + ch * 32 + i) = make_float4(h[ch * 16 + i], h[ch * 16 + i + 1], h[ch * 16 + i + 2], h[ch * 16 + i + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                         // neither CTA frees TMEM / exits while the pair's MMAs or arrivals are in flight
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(512));
}

// host: the pair weight image of one sequence (2 ranks x 24 stages x 16 KB, device pointer) as a tensor map of
// [768 rows][256 x u32]; box = 256 x 16 = one ring stage.  cuTensorMapEncodeTiled is fetched through the runtime
// (cudaGetDriverEntryPoint), so the library does not link libcuda.  Returns 0 on success.
inline int make_pair_tensor_map(const void* dev_img, CUtensorMap* out) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeFn>(f);
  }();
  if (!fn) return 1;
  const cuuint64_t gdim[2] = {256, (cuuint64_t)(2 * P_RANK_IMAGE_BYTES / P2_TM_ROW_BYTES)};
  const cuuint64_t gstride[1] = {P2_TM_ROW_BYTES};
  const cuuint32_t box[2] = {256, (cuuint32_t)P2_TM_BOX_ROWS};
  const cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(dev_img), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? 0 : 2;
}

}  // namespace r4tc
