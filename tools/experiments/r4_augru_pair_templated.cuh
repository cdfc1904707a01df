// r4_augru_pair.cuh -- the AUGRU recurrence (deepctr VecAttGRUCell, nets/utils.py:123-124) as ONE tensor-core
// instruction stream per 2-CTA cluster: tcgen05.mma.cta_group::2, M = 128 (64 feature rows per CTA), N = 256,
// bf16 hi/lo operands, fp32 accumulators in both CTAs' TMEM.  Same arithmetic as k_augru_tc (r4_augru_tc.cuh):
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
#include "../../rl4rs_b200/csrc/r4_augru_pair.cuh"   // helpers (mma2_bf16, commit2, mbar_wait_cl, ...) + constants

#ifndef R4PT_DEEPX
#define R4PT_DEEPX 0     // 1: load all 64 inputs of a gate phase BEFORE waiting for the gate (next-round experiment)
#endif

namespace r4tc {

// Descriptor of a SWIZZLE_NONE K-major operand split into its two words: `lo` carries the start address (>> 4, 14 bits)
// and LBO, `hi` carries SBO and the version bit.  Advancing the operand by `bytes` is `lo + (bytes >> 4)` (shared
// memory is < 256 KB, the address field cannot carry out), so a descriptor costs ONE add in the issue loop.
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo) { return ((saddr >> 4) & 0x3fffu) | ((lbo >> 4) << 16); }
__device__ __forceinline__ constexpr uint32_t desc_hi(uint32_t sbo) { return ((sbo >> 4) & 0x3fffu) | (1u << 14); }
__device__ __forceinline__ uint64_t desc_of(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// One gate of one step, issued by the leader's MMA thread: 8 ring stages x 2 K16 slices x (A_hi*B_hi + A_lo*B_hi +
// A_hi*B_lo).  Everything that can be is a compile-time constant: a step uses the ring 24 times = 4 revolutions of
// its 6 stages, so the stage index AND the mbarrier parity of use `u = GATE*8 + s8` are the same in every step, and all
// operand offsets are immediates.  (With a run-time stage counter the single issuing thread spent ~100 instructions
// per stage on descriptor arithmetic and register -> uniform-register moves and paced the MMAs at ~105 cycles instead
// of the 64-66 tools/ts_probe.cu measures for the same operands.)
// The r and u gates walk the K blocks of h in the order 0,2,4,6,1,3,5,7 -- the order in which the epilogue finishes
// them -- and the r gate waits for `half_bar` (the odd blocks) before its second half; build_pair_image lays the
// weights out in the same order.  No tcgen05 fence per stage: the weights come from TMA.
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .b32 r;\n\t.reg .pred p;\n\telect.sync r|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred;
}

template <int GATE>
__device__ __forceinline__ void issue_gate(uint32_t leader, uint32_t tbase, uint32_t aHi_lo, uint32_t aLo_lo, uint32_t b_lo, uint64_t* bar_full,
                                           uint64_t* bar_empty, uint64_t* half_bar, uint32_t half_par) {
  constexpr uint32_t idesc = make_idesc(TM, HID);
  constexpr uint32_t dcol = GATE == 0 ? P_TC_R : (GATE == 1 ? P_TC_U : P_TC_C);
  constexpr uint32_t a_hi = desc_hi(A_SBO), b_hi = desc_hi(B_SBO);
#pragma unroll
  for (int s8 = 0; s8 < NKB; ++s8) {
    const int u = GATE * NKB + s8;
    const int stage = u % P_NST;
    const uint32_t par = (uint32_t)((u / P_NST) & 1);
    const int kb = GATE < 2 ? ((s8 & 3) * 2 + (s8 >> 2)) : s8;
    if (GATE == 0 && s8 == NKB / 2) { mbar_wait_cl(half_bar, half_par); tc_fence_after(); }
    mbar_wait(&bar_full[stage], par);
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
    commit2(&bar_empty[stage]);
    }
    __syncwarp();
  }
}

// Wimg of a sequence here = [rank 2][24 stages][hi 8 KB | lo 8 KB] (build_pair_image), everything else as AugruTcParams.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1) k_augru_pair_t(AugruTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_full[P_NST], bar_empty[P_NST], bar_h0, bar_h1, bar_rh, bar_r, bar_u, bar_c;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint8_t* sHhi = smem;                       // A operand: h
  uint8_t* sHlo = smem + P_A_BYTES;
  uint8_t* sRhi = smem + 2 * P_A_BYTES;       // A operand: r*h
  uint8_t* sRlo = smem + 3 * P_A_BYTES;
  uint8_t* sB = smem + 4 * P_A_BYTES;
  const AugruTcSeq& S = p.s[blockIdx.y];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int m0 = (blockIdx.x >> 1) * TM;      // the pair's 128-row tile

  if (tid == 0) {
    // the leader's "full" collects its own TMA (expect_tx arrival + bytes) and the peer's relay arrival
    for (int i = 0; i < P_NST; ++i) { mbar_init(&bar_full[i], rank == 0 ? 2 : 1); mbar_init(&bar_empty[i], 1); }
    mbar_init(&bar_h0, 16); mbar_init(&bar_h1, 16); mbar_init(&bar_rh, 16);          // 8 epilogue warps x 2 CTAs, one arrival each
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

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");   // frees 128 x 128 registers = what 232 for the 256 epilogue threads takes
    if (warp == 9) {
      // ===== TMA producer: this CTA's half of the 24-stage weight stream of a step, 64 times =====
      if (lane == 0) {
        const uint8_t* img = S.Wimg + (size_t)rank * P_RANK_IMAGE_BYTES;
        int stage = 0; uint32_t phase = 0;
        for (int t = 0; t < STEPS; ++t) {
          const uint8_t* src = img;
          for (int i = 0; i < P_STAGES_PER_STEP; ++i, src += P_STAGE_BYTES) {
            if (stage % P_CG == 0) mbar_wait(&bar_empty[stage / P_CG], phase ^ 1);
            mbar_expect_tx(&bar_full[stage], P_STAGE_BYTES);
            bulk_g2s(sB + stage * P_STAGE_BYTES, src, P_STAGE_BYTES, &bar_full[stage]);
            if (++stage == P_NST) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 8 && rank == 1) {
      // ===== ring relay: tell the leader that this CTA's copy of stage s has landed =====
      if (lane == 0) {
        const uint32_t remote0 = mapa_rank(smem_u32(&bar_full[0]), 0);
        int stage = 0; uint32_t phase = 0;
        for (int i = 0; i < STEPS * P_STAGES_PER_STEP; ++i) {
          mbar_wait(&bar_full[stage], phase);
          arrive_cl_relaxed(remote0 + stage * 8);
          if (++stage == P_NST) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 8) {
      // ===== MMA issuer (leader CTA): the whole warp walks the loop in uniform control flow, ONE lane elected once issues
      // (tools/experiments/issue_loop_sass.cu: back-to-back UTCHMMA, no ELECT / BRA.U.ANY wrapper per instruction) =====
      const uint32_t leader = elect_one();
      {
        const uint32_t hHi = smem_u32(sHhi), hLo = smem_u32(sHlo), rHi = smem_u32(sRhi), rLo = smem_u32(sRlo), bBase = smem_u32(sB);
        uint32_t hHi_d = desc_lo(hHi, LBO), hLo_d = desc_lo(hLo, LBO), rHi_d = desc_lo(rHi, LBO), rLo_d = desc_lo(rLo, LBO),
                 b_d = desc_lo(bBase, LBO);
        for (int t = 0; t < STEPS; ++t) {
          const uint32_t par = t & 1;
          long long* dbg = (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && leader) ? p.dbg + t * 16 : nullptr;
          if (dbg) dbg[0] = clock64();
          // keep the 144 descriptors of a step OUT of the loop-invariant set: hoisted, they live in local memory (this warp
          // has 80 registers) and every MMA pays a local load; rebuilt from these five words each costs one add
          asm volatile("" : "+r"(hHi_d), "+r"(hLo_d), "+r"(rHi_d), "+r"(rLo_d), "+r"(b_d));
          mbar_wait_cl(&bar_h0, par);     // both CTAs' even K blocks of h (hi/lo) are in shared memory
          tc_fence_after();
          if (dbg) dbg[1] = clock64();
          issue_gate<0>(leader, tbase, hHi_d, hLo_d, b_d, bar_full, bar_empty, &bar_h1, par);   // ... the odd ones by its second half
          if (leader) commit2(&bar_r);
          if (dbg) dbg[2] = clock64();
          issue_gate<1>(leader, tbase, hHi_d, hLo_d, b_d, bar_full, bar_empty, nullptr, 0);
          if (leader) commit2(&bar_u);
          if (dbg) dbg[3] = clock64();
          mbar_wait_cl(&bar_rh, par);     // both CTAs' r*h written
          tc_fence_after();
          if (dbg) dbg[4] = clock64();
          issue_gate<2>(leader, tbase, rHi_d, rLo_d, b_d, bar_full, bar_empty, nullptr, 0);
          if (leader) commit2(&bar_c);
          if (dbg) { dbg[5] = clock64(); dbg[6] = 0; dbg[7] = 0; }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ===== epilogue warps: thread = (row of this CTA, 64 hidden columns) =====
    const int q = warp & 3, sub = warp >> 2;
    const int rl = (q & 1) * 32 + lane;                    // row inside this CTA
    const int prow = (int)rank * P_RC + rl;                // row inside the pair's 128-row tile
    const int hc0 = (q >> 1) * 128 + sub * 64;             // first hidden column of this thread
    const uint32_t tcol = (uint32_t)sub * 64;              // TMEM column offset inside a gate
    int r = m0 + prow;
    const bool valid = r < p.R;
    if (!valid) r = p.R - 1;
    const int ci = S.shared ? 0 : (p.row0 + r) / p.div;
    const float* xt = S.XT + ((size_t)(ci / TM) * STEPS) * XT_COLS * TM + (ci % TM);
    const float* st = S.scoresT + ((size_t)(m0 / TM) * STEPS) * TM + prow;
    const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
    const uint32_t a_row_off = (uint32_t)(rl / 8) * A_SBO + (uint32_t)(rl % 8) * 16;
    const uint32_t bar_h0_leader = mapa_rank(smem_u32(&bar_h0), 0), bar_h1_leader = mapa_rank(smem_u32(&bar_h1), 0), bar_rh_leader = mapa_rank(smem_u32(&bar_rh), 0);
    float h[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) h[i] = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) {                          // h0 = 0 into the A operand
      uint32_t off = a_row_off + (uint32_t)((hc0 + g * 8) / 8) * LBO;
      *reinterpret_cast<uint4*>(sHhi + off) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sHlo + off) = make_uint4(0, 0, 0, 0);
    }
    proxy_fence();
    __syncwarp();
    if (lane == 0) {
      if (rank == 0) { mbar_arrive(&bar_h0); mbar_arrive(&bar_h1); }
      else { arrive_cl_relaxed(bar_h0_leader); arrive_cl_relaxed(bar_h1_leader); }
    }

    for (int t = 0; t < STEPS; ++t) {
      const uint32_t par = t & 1;
      long long* dbg = (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) ? p.dbg + t * 16 : nullptr;
      const float* xs = xt + (size_t)t * XT_COLS * TM;
      const float one_minus_s = 1.0f - __ldg(st + (size_t)t * TM);
      // Pull this CTA's half of the NEXT step's input lines (768 columns x 2 lines) from HBM into L2.
      if (t + 1 < STEPS) {
        const float* xn = S.XT + (((size_t)(ci / TM) * STEPS + (t + 1)) * XT_COLS) * TM;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int id = i * 256 + tid;                    // 1536 lines: (column, 32-lane group of this CTA)
          asm volatile("prefetch.global.L2 [%0];" :: "l"(xn + (size_t)(id >> 1) * TM + (rank * 2 + (id & 1)) * 32));
        }
      }
#define R4P_LOADX(dst, colbase) _Pragma("unroll") for (int j = 0; j < 16; ++j) dst[j] = __ldg(xs + (size_t)((colbase) + j) * TM)
      // ---- phase R (overlaps the u MMAs): r*h -> its own A operand ----
      {
#if R4PT_DEEPX
        float x[4][16], a[2][16];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) { R4P_LOADX(x[c4], hc0 + c4 * 16); }
#else
        float x[2][16], a[2][16];
        R4P_LOADX(x[0], hc0);
#endif
        if (dbg) dbg[8] = clock64();
        mbar_wait(&bar_r, par);
        if (dbg) dbg[9] = clock64();
        tc_fence_after();
        tmem_ld16(tlane + P_TC_R + tcol, a[0]);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int cur = ch & 1, nxt = cur ^ 1;
          tmem_wait_ld();
#if R4PT_DEEPX
          if (ch < 3) tmem_ld16(tlane + P_TC_R + tcol + (ch + 1) * 16, a[nxt]);
#pragma unroll
          for (int j = 0; j < 16; ++j) a[cur][j] = fast_sigmoid(a[cur][j] + x[ch][j]) * h[ch * 16 + j];
#else
          if (ch < 3) { R4P_LOADX(x[nxt], hc0 + (ch + 1) * 16); tmem_ld16(tlane + P_TC_R + tcol + (ch + 1) * 16, a[nxt]); }
#pragma unroll
          for (int j = 0; j < 16; ++j) a[cur][j] = fast_sigmoid(a[cur][j] + x[cur][j]) * h[ch * 16 + j];
#endif
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 hi, lo;
            split8(a[cur] + g * 8, hi, lo);
            uint32_t off = a_row_off + (uint32_t)((hc0 + ch * 16 + g * 8) / 8) * LBO;
            *reinterpret_cast<uint4*>(sRhi + off) = hi;
            *reinterpret_cast<uint4*>(sRlo + off) = lo;
          }
        }
      }
      tc_fence_before();
      proxy_fence();
      __syncwarp();
      if (lane == 0) { if (rank == 0) mbar_arrive(&bar_rh); else arrive_cl_relaxed(bar_rh_leader); }
      if (dbg) dbg[10] = clock64();
      // ---- phase U (overlaps the c MMAs): E = 1 + exp(-(acc_u + Xu)) back into TMEM ----
      {
#if R4PT_DEEPX
        float x[4][16], a[2][16];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) { R4P_LOADX(x[c4], HID + hc0 + c4 * 16); }
#else
        float x[2][16], a[2][16];
        R4P_LOADX(x[0], HID + hc0);
#endif
        mbar_wait(&bar_u, par);
        if (dbg) dbg[11] = clock64();
        tc_fence_after();
        tmem_ld16(tlane + P_TC_U + tcol, a[0]);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int cur = ch & 1, nxt = cur ^ 1;
          tmem_wait_ld();
#if R4PT_DEEPX
          if (ch < 3) tmem_ld16(tlane + P_TC_U + tcol + (ch + 1) * 16, a[nxt]);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            a[cur][j] = 1.0f + ex2_approx(fminf(-1.4426950408889634f * (a[cur][j] + x[ch][j]), 60.0f));
#else
          if (ch < 3) { R4P_LOADX(x[nxt], HID + hc0 + (ch + 1) * 16); tmem_ld16(tlane + P_TC_U + tcol + (ch + 1) * 16, a[nxt]); }
#pragma unroll
          for (int j = 0; j < 16; ++j)
            a[cur][j] = 1.0f + ex2_approx(fminf(-1.4426950408889634f * (a[cur][j] + x[cur][j]), 60.0f));
#endif
          tmem_st16(tlane + P_TC_U + tcol + ch * 16, a[cur]);
        }
        tmem_wait_st();
      }
      if (dbg) dbg[12] = clock64();
      // ---- phase C: c = tanh(acc_c + Xc) = 1 - 2/(1 + F), u = 1/E with ONE reciprocal of E*F ----
      {
#if R4PT_DEEPX
        float x[4][16], a[2][16], u[2][16];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) { R4P_LOADX(x[c4], 2 * HID + hc0 + c4 * 16); }
#else
        float x[2][16], a[2][16], u[2][16];
        R4P_LOADX(x[0], 2 * HID + hc0);
#endif
        mbar_wait(&bar_c, par);
        if (dbg) dbg[13] = clock64();
        tc_fence_after();
        tmem_ld16(tlane + P_TC_C + tcol, a[0]);
        tmem_ld16(tlane + P_TC_U + tcol, u[0]);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int cur = ch & 1, nxt = cur ^ 1;
          tmem_wait_ld();
          if (ch < 3) {
#if !R4PT_DEEPX
            R4P_LOADX(x[nxt], 2 * HID + hc0 + (ch + 1) * 16);
#endif
            tmem_ld16(tlane + P_TC_C + tcol + (ch + 1) * 16, a[nxt]);
            tmem_ld16(tlane + P_TC_U + tcol + (ch + 1) * 16, u[nxt]);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float F = 1.0f + ex2_approx(fminf(2.8853900817779268f * (a[cur][j] + x[R4PT_DEEPX ? ch : cur][j]), 60.0f));
            const float E = u[cur][j];
            const float rc = rcp_approx(E * F);                      // E, F <= 1 + 2^60: the product is finite
            const float c = fmaf(-2.0f, rc * E, 1.0f);               // tanh
            const float up = one_minus_s * (rc * F);                 // (1 - s) sigmoid
            const float hn = fmaf(up, h[ch * 16 + j] - c, c);        // u' h + (1 - u') c
            h[ch * 16 + j] = hn;
            a[cur][j] = hn;
          }
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 hi, lo;
            split8(a[cur] + g * 8, hi, lo);
            uint32_t off = a_row_off + (uint32_t)((hc0 + ch * 16 + g * 8) / 8) * LBO;
            *reinterpret_cast<uint4*>(sHhi + off) = hi;
            *reinterpret_cast<uint4*>(sHlo + off) = lo;
          }
          if (ch == 1) {              // this thread's even K block of h' is complete: release the first half of the next r gate
            proxy_fence();
            __syncwarp();
            if (lane == 0) { if (rank == 0) mbar_arrive(&bar_h0); else arrive_cl_relaxed(bar_h0_leader); }
          }
        }
      }
#undef R4P_LOADX
      tc_fence_before();
      proxy_fence();
      __syncwarp();
      if (dbg) dbg[14] = clock64();
      if (lane == 0) { if (rank == 0) mbar_arrive(&bar_h1); else arrive_cl_relaxed(bar_h1_leader); }
    }
    if (valid) {
      float* o = S.out + (size_t)(m0 + prow) * p.out_ld + hc0;
#pragma unroll
      for (int i = 0; i < 64; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(h[i], h[i + 1], h[i + 2], h[i + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                         // neither CTA frees TMEM / exits while the pair's MMAs or arrivals are in flight
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(512));
}

}  // namespace r4tc
