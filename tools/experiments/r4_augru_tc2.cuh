// r4_augru_tc2.cuh -- AUGRU recurrence, version 2: a CLUSTER PAIR of CTAs per 128-row tile.
//
// Why: with one CTA per tile (r4_augru_tc.cuh) the r/u/c accumulators (3 x 256 fp32 columns) do not fit
// the 512 TMEM columns, so c aliases r and the tensor pipe idles while the epilogue converts r*h and h
// (measured 56 % tensor-active), and a 4096-row pass fills only 64 of 148 SMs.  Here the two CTAs of a
// cluster split the 256 output columns: CTA q owns columns [128q, 128q+128) of r, u and c.
//   * TMEM per CTA: r | c | u(even step) | u(odd step) = 4 x 128 columns, no aliasing  ->  the c-GEMM chases the
//     r*h K blocks as the epilogue produces them, and the next step's u/r GEMMs chase the new h.
//   * every CTA needs the full K = 256 operand: epilogue warps write their bf16 hi/lo groups into the LOCAL A
//     buffer (K-block-major layout: a warp's 32 rows x one 32-wide K block are 2 KB contiguous per split) and
//     one lane pushes the finished block to the peer with cp.async.bulk shared::cta -> shared::cluster
//     (asynchronous; per-thread st.shared::cluster measured ~11 B/clk and stalled the epilogue);
//     per-K-block mbarriers (4 warp arrivals + the copy's tx bytes) release the MMA issuer block by block.
//   * "all r (c) GEMMs of the pair are finished" is one mbarrier per CTA with 2 arrivals, fed by
//     tcgen05.commit ... multicast::cluster from both issuers (the A buffers are about to be overwritten).
//   * twice the CTAs per pass, half the weight bytes per CTA (each streams only its column half).
// Numerics identical to version 1 (bf16 hi/lo split, 3 products, fp32 state in registers).
#pragma once
#include "../../rl4rs_b200/csrc/r4_augru_tc.cuh"

namespace r4tc2 {
using namespace r4tc;

constexpr int NH = 128;                         // output columns per CTA
constexpr int NST2 = 8;                         // weight ring stages
constexpr int STAGE2_BYTES = NH * KB * 2;       // 8192
constexpr int STAGES_PER_STEP2 = 3 * NKB * 2;   // 48 stages of 8 KB per CTA per step
constexpr int W_IMAGE2_BYTES = 2 * STAGES_PER_STEP2 * STAGE2_BYTES;   // both ranks: 786432
constexpr int SMEM2_BYTES = 2 * A_BYTES + NST2 * STAGE2_BYTES + 1024;
constexpr int NTHREADS2 = 384;
// TMEM columns
constexpr int T_R = 0, T_C = 128, T_U0 = 256;
// A operand, K-block-major: [kb 8][row group 16][4 core matrices x 128 B]
constexpr int A2_KB_BYTES = TM * KB * 2;         // 8192 per split per K block
constexpr int A2_SBO = (KB / 8) * 128;           // 512 between 8-row groups
constexpr int WARP_KB_BYTES = 32 * KB * 2;       // 2048: one warp's rows of one K block
// K blocks become ready in this order (each 64-column thread group finishes its first block, then its second)
__host__ __device__ constexpr int kb_order(int i) { return (i & 3) * 2 + (i >> 2); }   // 0,2,4,6,1,3,5,7

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t laddr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(laddr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t caddr, const uint4& v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(caddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_remote(uint32_t caddr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" :: "r"(caddr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_s2peer(uint32_t dst_caddr, uint32_t src_laddr, uint32_t bytes, uint32_t bar_caddr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst_caddr), "r"(src_laddr), "r"(bytes), "r"(bar_caddr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local_cl(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
               "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n"
               :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void proxy_fence_all() { asm volatile("fence.proxy.async.shared::cluster;" ::: "memory"); }

// host: per rank q, stream order u, r, c; K blocks in kb_order; (hi, lo); stage = [128 n x 32 k] core matrices
inline void build_weight_image2(const float* Wg, const float* Wc, uint8_t* img) {
  for (int q = 0; q < 2; ++q)
    for (int mat = 0; mat < 3; ++mat)
      for (int i = 0; i < NKB; ++i)
        for (int sp = 0; sp < 2; ++sp) {
          const int kb = kb_order(i);
          uint8_t* st = img + (size_t)(((q * 3 + mat) * NKB + i) * 2 + sp) * STAGE2_BYTES;
          for (int n = 0; n < NH; ++n)
            for (int kk = 0; kk < KB; ++kk) {
              int k = kb * KB + kk, col = q * NH + n;
              float w = mat == 0 ? Wg[(size_t)k * 2 * HID + HID + col] : (mat == 1 ? Wg[(size_t)k * 2 * HID + col] : Wc[(size_t)k * HID + col]);
              uint16_t hi = host_bf16_bits(w);
              uint16_t v = sp == 0 ? hi : host_bf16_bits(w - host_bf16_val(hi));
              memcpy(st + (n / 8) * B_SBO + (kk / 8) * LBO + (n % 8) * 16 + (kk % 8) * 2, &v, 2);
            }
        }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS2, 1) k_augru_tc2(AugruTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sAhi = smem;
  uint8_t* sAlo = smem + A_BYTES;
  uint8_t* sB = smem + 2 * A_BYTES;
  __shared__ uint64_t bar_full[NST2], bar_empty[NST2], bar_hk[NKB], bar_rhk[NKB], bar_u, bar_r, bar_c;
  __shared__ uint32_t tmem_base_s;
  const AugruTcSeq& S = p.s[blockIdx.y];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank(), peer = rank ^ 1;
  const int m0 = (blockIdx.x >> 1) * TM;

  if (tid == 0) {
    for (int i = 0; i < NST2; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
    for (int i = 0; i < NKB; ++i) {          // owner CTA: 4 local warp arrivals; the other: 1 remote arrive + the copy's tx bytes
      int cnt = ((uint32_t)(i / 4) == cluster_rank()) ? 4 : 1;
      mbar_init(&bar_hk[i], cnt); mbar_init(&bar_rhk[i], cnt);
    }
    mbar_init(&bar_u, 1); mbar_init(&bar_r, 2); mbar_init(&bar_c, 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // peer's barriers are initialised before anyone arrives remotely
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 9) {
      // ===== TMA producer: this rank's half of the weight stream =====
      if (lane == 0) {
        int stage = 0; uint32_t phase = 0;
        const uint8_t* base = S.Wimg + (size_t)rank * STAGES_PER_STEP2 * STAGE2_BYTES;
        for (int t = 0; t < STEPS; ++t) {
          const uint8_t* src = base;
          for (int i = 0; i < STAGES_PER_STEP2; ++i, src += STAGE2_BYTES) {
            mbar_wait(&bar_empty[stage], phase ^ 1);
            mbar_expect_tx(&bar_full[stage], STAGE2_BYTES);
            bulk_g2s(sB + stage * STAGE2_BYTES, src, STAGE2_BYTES, &bar_full[stage]);
            if (++stage == NST2) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 10) {
      // ===== forwarder: pushes every K block this CTA produced to the peer (asynchronous bulk copies) =====
      if (lane == 0) {
        const uint32_t lAhi = smem_u32(sAhi), lAlo = smem_u32(sAlo);
        const uint32_t rAhi = mapa(lAhi, peer), rAlo = mapa(lAlo, peer);
        auto forward = [&](uint64_t* bars, uint32_t par) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int kb = 4 * rank + (i & 1) * 2 + (i >> 1);        // 4r+0, 4r+2, 4r+1, 4r+3
            mbar_wait_cl(&bars[kb], par);
            const uint32_t rbar = mapa(smem_u32(&bars[kb]), peer);
            const uint32_t off = (uint32_t)kb * A2_KB_BYTES;
            mbar_arrive_expect_tx_remote(rbar, 2 * A2_KB_BYTES);
            bulk_s2peer(rAhi + off, lAhi + off, A2_KB_BYTES, rbar);
            bulk_s2peer(rAlo + off, lAlo + off, A2_KB_BYTES, rbar);
          }
        };
        forward(bar_hk, 0);                                          // h0
        for (int t = 0; t < STEPS; ++t) {
          forward(bar_rhk, t & 1);
          forward(bar_hk, (t + 1) & 1);
        }
      }
    } else if (warp == 8) {
      // ===== MMA issuer =====
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc(TM, NH);
        const uint32_t aHi = smem_u32(sAhi), aLo = smem_u32(sAlo), bBase = smem_u32(sB);
        int stage = 0; uint32_t phase = 0;
        // one K block (32 wide) of one GEMM: 2 weight stages (hi, lo), 6 MMAs
        auto kblock = [&](uint32_t dcol, int kb, bool first) {
          mbar_wait(&bar_full[stage], phase);
          tc_fence_after();
          {
            uint32_t b = bBase + stage * STAGE2_BYTES;
#pragma unroll
            for (int j = 0; j < KB / 16; ++j) {
              uint64_t db = make_desc(b + j * 2 * LBO, LBO, B_SBO);
              uint32_t koff = kb * A2_KB_BYTES + j * 2 * LBO;
              mma_bf16(tbase + dcol, make_desc(aHi + koff, LBO, A2_SBO), db, idesc, (first && j == 0) ? 0u : 1u);
              mma_bf16(tbase + dcol, make_desc(aLo + koff, LBO, A2_SBO), db, idesc, 1u);
            }
          }
          umma_commit(&bar_empty[stage]);
          if (++stage == NST2) { stage = 0; phase ^= 1; }
          mbar_wait(&bar_full[stage], phase);
          tc_fence_after();
          {
            uint32_t b = bBase + stage * STAGE2_BYTES;
#pragma unroll
            for (int j = 0; j < KB / 16; ++j) {
              uint64_t db = make_desc(b + j * 2 * LBO, LBO, B_SBO);
              uint32_t koff = kb * A2_KB_BYTES + j * 2 * LBO;
              mma_bf16(tbase + dcol, make_desc(aHi + koff, LBO, A2_SBO), db, idesc, 1u);
            }
          }
          umma_commit(&bar_empty[stage]);
          if (++stage == NST2) { stage = 0; phase ^= 1; }
        };
        for (int t = 0; t < STEPS; ++t) {
          const uint32_t par = t & 1;
          const uint32_t ucol = T_U0 + (t & 1) * NH;
          for (int i = 0; i < NKB; ++i) {        // u chases the h K blocks of this step
            const int kb = kb_order(i);
            mbar_wait_cl(&bar_hk[kb], par);
            tc_fence_after();
            kblock(ucol, kb, i == 0);
          }
          umma_commit(&bar_u);
          for (int i = 0; i < NKB; ++i) kblock(T_R, kb_order(i), i == 0);
          umma_commit_pair(&bar_r);              // -> both CTAs: "u and r GEMMs of this CTA no longer read h"
          for (int i = 0; i < NKB; ++i) {        // c chases the r*h K blocks
            const int kb = kb_order(i);
            mbar_wait_cl(&bar_rhk[kb], par);
            tc_fence_after();
            kblock(T_C, kb, i == 0);
          }
          umma_commit_pair(&bar_c);
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ===== epilogue: thread = (row, 64-column quarter of the 256 columns) =====
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;
    const int lc0 = half * 64;                    // first local column (TMEM, within this CTA's 128)
    const int gc0 = rank * NH + lc0;              // first global column (A operand K index, XT column)
    int r = m0 + row;
    const bool valid = r < p.R;
    if (!valid) r = p.R - 1;
    const int ci = S.shared ? 0 : (p.row0 + r) / p.div;
    const float* xt = S.XT + ((size_t)(ci / TM) * STEPS) * XT_COLS * TM + (ci % TM);
    const float* st = S.scoresT + ((size_t)((m0 + row) / TM) * STEPS) * TM + row;
    const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
    const uint32_t a_row_off = (uint32_t)(row / 8) * A2_SBO + (uint32_t)(row % 8) * 16;
    const int kb_first = gc0 / KB;                // this thread's two K blocks: kb_first, kb_first + 1
    float h[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) h[i] = 0.f;

    // one 8-column group (hi, lo) into the LOCAL A operand (K-block-major)
    auto put8 = [&](int gcol, const uint4& hi, const uint4& lo) {
      uint32_t off = (uint32_t)(gcol / KB) * A2_KB_BYTES + a_row_off + (uint32_t)((gcol % KB) / 8) * LBO;
      *reinterpret_cast<uint4*>(sAhi + off) = hi;
      *reinterpret_cast<uint4*>(sAlo + off) = lo;
    };
    // this warp's 32 rows of K block kb are complete locally: release the local issuer and the forwarder
    auto publish = [&](uint64_t* bars, int kb) {
#ifndef R4_EXPERIMENT_NOFENCE
      proxy_fence();                               // my generic-proxy writes -> visible to the async proxy
#endif
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[kb]);       // cta-scope release: the consumers (issuer, forwarder) are local
    };
    // h0 = 0
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      put8(gc0 + g * 8, make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0));
      if ((g & 3) == 3) publish(bar_hk, kb_first + (g >> 2));
    }

    for (int t = 0; t < STEPS; ++t) {
      const uint32_t par = t & 1;
      const float* xs = xt + (size_t)t * XT_COLS * TM;
      const float one_minus_s = 1.0f - __ldg(st + (size_t)t * TM);
      const uint32_t ucol = T_U0 + (t & 1) * NH + lc0;
      const bool dbg = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0;
#define R4_STAMP(i) if (dbg) p.dbg[t * 16 + (i)] = clock64()
#define R4_LOADX(dst, colbase) _Pragma("unroll") for (int j = 0; j < 16; ++j) dst[j] = __ldg(xs + (size_t)((colbase) + j) * TM)
      // ---- phase U: u' = (1 - s) sigmoid(acc_u + Xu) -> back into TMEM (in place) ----
      {
        float x[4][16], a[2][16];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) R4_LOADX(x[c4], HID + gc0 + c4 * 16);
        R4_STAMP(0);
        mbar_wait(&bar_u, par);
        R4_STAMP(1);
        tc_fence_after();
        tmem_ld16(tlane + ucol, a[0]);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int cur = ch & 1, nxt = cur ^ 1;
          tmem_wait_ld();
          if (ch < 3) tmem_ld16(tlane + ucol + (ch + 1) * 16, a[nxt]);
#pragma unroll
          for (int j = 0; j < 16; ++j) a[cur][j] = one_minus_s * fast_sigmoid(a[cur][j] + x[ch][j]);
          tmem_st16(tlane + ucol + ch * 16, a[cur]);
        }
        tmem_wait_st();
      }
      // ---- phase R: r*h -> both A operands, K block by K block ----
      {
        float x[4][16], a[2][16];
        // all X loads of the phase are issued BEFORE the wait: no global load may be in flight at a
        // fence.proxy.async below (the fence waits for the thread's outstanding loads: ~2k cycles each)
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) R4_LOADX(x[c4], gc0 + c4 * 16);
        R4_STAMP(2);
        mbar_wait_cl(&bar_r, par);               // every u/r GEMM of the pair has finished reading h
        R4_STAMP(3);
        tc_fence_after();
        tmem_ld16(tlane + T_R + lc0, a[0]);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int cur = ch & 1, nxt = cur ^ 1;
          tmem_wait_ld();
          if (ch < 3) tmem_ld16(tlane + T_R + lc0 + (ch + 1) * 16, a[nxt]);
#pragma unroll
          for (int j = 0; j < 16; ++j) a[cur][j] = fast_sigmoid(a[cur][j] + x[ch][j]) * h[ch * 16 + j];
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 hi, lo;
            split8(a[cur] + g * 8, hi, lo);
            put8(gc0 + ch * 16 + g * 8, hi, lo);
          }
          if (ch == 1) R4_STAMP(8);
          if (ch == 3) R4_STAMP(10);
          if (ch & 1) publish(bar_rhk, kb_first + (ch >> 1));
          if (ch == 1) R4_STAMP(9);
          if (ch == 3) R4_STAMP(11);
        }
      }
      // ---- phase C: c = tanh(acc_c + Xc); h <- u' h + (1 - u') c -> both A operands ----
      {
        float x[4][16], a[2][16], u[2][16];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) R4_LOADX(x[c4], 2 * HID + gc0 + c4 * 16);
        R4_STAMP(4);
        mbar_wait_cl(&bar_c, par);               // every c GEMM of the pair has finished reading r*h
        R4_STAMP(5);
        tc_fence_after();
        tmem_ld16(tlane + T_C + lc0, a[0]);
        tmem_ld16(tlane + ucol, u[0]);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int cur = ch & 1, nxt = cur ^ 1;
          tmem_wait_ld();
          if (ch < 3) {
            tmem_ld16(tlane + T_C + lc0 + (ch + 1) * 16, a[nxt]);
            tmem_ld16(tlane + ucol + (ch + 1) * 16, u[nxt]);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float c = fast_tanh(a[cur][j] + x[ch][j]);
            float hn = fmaf(u[cur][j], h[ch * 16 + j] - c, c);
            h[ch * 16 + j] = hn;
            a[cur][j] = hn;
          }
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 hi, lo;
            split8(a[cur] + g * 8, hi, lo);
            put8(gc0 + ch * 16 + g * 8, hi, lo);
          }
          if (ch & 1) {
            tc_fence_before();
            publish(bar_hk, kb_first + (ch >> 1));
          }
        }
      }
      R4_STAMP(6);
#undef R4_LOADX
#undef R4_STAMP
    }
    if (valid) {
      float* o = S.out + (size_t)(m0 + row) * p.out_ld + gc0;
#pragma unroll
      for (int i = 0; i < 64; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(h[i], h[i + 1], h[i + 2], h[i + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // nobody exits while the peer may still write here
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(512));
}

}  // namespace r4tc2
