// r4_augru_tc.cuh -- the AUGRU recurrence (deepctr VecAttGRUCell, nets/utils.py:123-124) on the
// 5th-generation tensor cores: tcgen05.mma (kind::f16, bf16 operands, fp32 accumulators in TMEM),
// weights streamed by 1-D TMA (cp.async.bulk) into a shared-memory ring, gate math in epilogue warps.
//
// One CTA = 128 feature rows x 64 steps of one sequence.  Per step, with h the 256-wide state:
//     u = sigmoid(Xu_t + h Wu)         D[:,256:512]   (tensor core, K = 256)
//     r = sigmoid(Xr_t + h Wr)         D[:,  0:256]
//     c = tanh   (Xc_t + (r*h) Wc)     D[:,  0:256]   (after r has been consumed)
//     u' = (1 - score_t) u ;  h <- u' h + (1 - u') c
// fp32 parity on a bf16 tensor pipe: every fp32 operand x is split x = hi + lo (both bf16, lo =
// bf16(x - hi)) and each product is issued as hi*hi + lo*hi + hi*lo (3 MMAs, fp32 accumulate):
// relative error ~2^-16 per term, measured 1e-5..3e-5 of rms on the 64-step recurrence against
// f64 (tools/split_sim.py) -- inside the 1e-4 parity bound.  The state h itself stays fp32 in the
// epilogue threads' registers; only the MMA operand copies are rounded.
//
// Shared memory (216 KB): A operand = h (then r*h) as bf16 hi + lo, SWIZZLE_NONE K-major core
// matrices (8 rows x 16 B), 2 x 64 KB; B ring = 5 stages x 16 KB (one 32-wide K block of one
// weight split, pre-tiled on the host in exactly this layout so a stage is ONE contiguous bulk
// copy); mbarriers.  TMEM: 512 columns (r|c in 0..255, u in 256..511).
// Warp roles: 0-7 epilogue (thread = row x column half), 8 MMA issuer (one elected lane), 9 TMA producer,
// 10-11 idle (they complete the control warpgroup so setmaxnreg can move registers to the epilogue).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>

namespace r4tc {

constexpr int TM = 128;                 // rows per CTA tile (UMMA_M)
constexpr int HID = 256;                // AUGRU hidden = GEMM N and K
constexpr int STEPS = 64;
constexpr int KB = 32;                  // K elements per B stage
constexpr int NKB = HID / KB;           // 8 K blocks per matrix
constexpr int NST = 5;                  // B ring stages
constexpr int STAGE_BYTES = HID * KB * 2;            // 16384
constexpr int A_BYTES = TM * HID * 2;                // 65536 per split
constexpr int LBO = 128;                             // K-adjacent core matrices
constexpr int A_SBO = (HID / 8) * 128;               // 4096: 8-row groups of the A operand
constexpr int B_SBO = (KB / 8) * 128;                // 512:  8-row groups inside a B stage
constexpr int STAGES_PER_STEP = 3 * NKB * 2;         // u, r, c  x  8 K blocks  x  (hi, lo) = 48
constexpr int W_IMAGE_BYTES = STAGES_PER_STEP * STAGE_BYTES;   // 786432 per sequence
constexpr int XT_COLS = 3 * HID;                     // transposed input halves: [r | u | c] rows of 128 lanes
constexpr int SMEM_BYTES = 2 * A_BYTES + NST * STAGE_BYTES + 1024;
constexpr int NTHREADS = 384;               // 8 epilogue warps + one control warpgroup (MMA, TMA, 2 idle)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);   // version 1, SWIZZLE_NONE
}
__device__ __forceinline__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n"
               :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                 "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  const uint32_t* u = reinterpret_cast<const uint32_t*>(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               :: "r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]),
                  "r"(u[8]), "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MUFU-based gate functions: ex2.approx (rel. error 2^-22) + rcp.approx (1 ulp); saturate correctly
// (ex2 -> +inf gives rcp -> 0).  No slow-path calls (the IEEE __frcp_rn costs a CALL per element).
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_sigmoid(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float fast_tanh(float x) { return fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(2.8853900817779268f * x)), 1.0f); }

// split 8 consecutive fp32 into bf16 hi / lo vectors (16 B each)
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 hh = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
    float r0 = x[2 * i] - __bfloat162float(hh.x), r1 = x[2 * i + 1] - __bfloat162float(hh.y);
    __nv_bfloat162 ll = __floats2bfloat162_rn(r0, r1);
    h[i] = *reinterpret_cast<uint32_t*>(&hh);
    l[i] = *reinterpret_cast<uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// 16 consecutive columns (colbase % 16 == 0) of one lane of a lane-major tile-step in the quad layout
// [col / 4][lane][col % 4] (r4_gemm_tc.cuh: xt_index): four 128-bit loads.  `ts` = tile-step base, `ln4` = lane * 4.
#ifndef R4_LDG_NOALLOC
#define R4_LDG_NOALLOC 0     // 1: the input halves are read once per CTA -> ld.global.nc.L1::no_allocate (probe switch)
#endif
__device__ __forceinline__ float4 ldg_x4(const float* p) {
#if R4_LDG_NOALLOC
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
#else
  return __ldg(reinterpret_cast<const float4*>(p));
#endif
}
__device__ __forceinline__ void load_x16(float* dst, const float* ts, int colbase, int ln4) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float4 v = ldg_x4(ts + (size_t)(colbase + 4 * g) * TM + ln4);
    dst[4 * g] = v.x; dst[4 * g + 1] = v.y; dst[4 * g + 2] = v.z; dst[4 * g + 3] = v.w;
  }
}

struct AugruTcSeq {
  const float* XT;        // transposed input halves [n_tiles_cached, 64, 768 / 4, 128, 4]  (tile, step, column quad, lane, column % 4)
  const uint8_t* Wimg;    // pre-tiled bf16 hi/lo weight image, W_IMAGE_BYTES, stream order u, r, c
  const float* scoresT;   // [n_row_tiles, 64, 128]
  float* out;             // final state, row stride out_ld
  int shared;             // 1: every row reads cached sequence 0
};
struct AugruTcParams {
  AugruTcSeq s[2];
  int R, row0, div, out_ld;
  long long* dbg;         // optional: per-step phase timestamps of CTA 0 (development probe), else null
};

__global__ void __launch_bounds__(NTHREADS, 1) k_augru_tc(AugruTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sAhi = smem;
  uint8_t* sAlo = smem + A_BYTES;
  uint8_t* sB = smem + 2 * A_BYTES;
  __shared__ uint64_t bar_full[NST], bar_empty[NST], bar_h, bar_u, bar_r, bar_rh, bar_c;
  __shared__ uint32_t tmem_base_s;
  const AugruTcSeq& S = p.s[blockIdx.y];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * TM;

  if (tid == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
    mbar_init(&bar_h, 256); mbar_init(&bar_rh, 256);
    mbar_init(&bar_u, 1); mbar_init(&bar_r, 1); mbar_init(&bar_c, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  // Register budget: an SMSP hosts 2 epilogue warps + 1 control warp (16384 regs): 2*232 + 40 fits.
  if (warp >= 8) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 9) {
    // ===== TMA producer: the 48-stage weight stream of a step, repeated 64 times =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < STEPS; ++t) {
        const uint8_t* src = S.Wimg;
        for (int i = 0; i < STAGES_PER_STEP; ++i, src += STAGE_BYTES) {
          mbar_wait(&bar_empty[stage], phase ^ 1);
          mbar_expect_tx(&bar_full[stage], STAGE_BYTES);
          bulk_g2s(sB + stage * STAGE_BYTES, src, STAGE_BYTES, &bar_full[stage]);
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 8) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TM, HID);
      const uint32_t aHi = smem_u32(sAhi), aLo = smem_u32(sAlo), bBase = smem_u32(sB);
      int stage = 0; uint32_t phase = 0;
      auto gemm = [&](uint32_t dcol) {
        for (int kb = 0; kb < NKB; ++kb) {
          // hi split of this K block: A_hi*B_hi + A_lo*B_hi
          mbar_wait(&bar_full[stage], phase);
          tc_fence_after();
          {
            uint32_t b = bBase + stage * STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < KB / 16; ++j) {
              uint64_t db = make_desc(b + j * 2 * LBO, LBO, B_SBO);
              uint32_t koff = (kb * (KB / 16) + j) * 2 * LBO;
              mma_bf16(tbase + dcol, make_desc(aHi + koff, LBO, A_SBO), db, idesc, (kb | j) ? 1u : 0u);
              mma_bf16(tbase + dcol, make_desc(aLo + koff, LBO, A_SBO), db, idesc, 1u);
            }
          }
          umma_commit(&bar_empty[stage]);
          if (++stage == NST) { stage = 0; phase ^= 1; }
          // lo split: A_hi*B_lo
          mbar_wait(&bar_full[stage], phase);
          tc_fence_after();
          {
            uint32_t b = bBase + stage * STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < KB / 16; ++j) {
              uint64_t db = make_desc(b + j * 2 * LBO, LBO, B_SBO);
              uint32_t koff = (kb * (KB / 16) + j) * 2 * LBO;
              mma_bf16(tbase + dcol, make_desc(aHi + koff, LBO, A_SBO), db, idesc, 1u);
            }
          }
          umma_commit(&bar_empty[stage]);
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      };
      for (int t = 0; t < STEPS; ++t) {
        uint32_t par = t & 1;
        mbar_wait(&bar_h, par);          // h (hi/lo) of this step is in shared memory
        tc_fence_after();
        gemm(HID);                       // u
        umma_commit(&bar_u);
        gemm(0);                         // r
        umma_commit(&bar_r);
        mbar_wait(&bar_rh, par);         // r*h written, r accumulators consumed
        tc_fence_after();
        gemm(0);                         // c
        umma_commit(&bar_c);
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ===== epilogue warps: thread = (row, column half) =====
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;
    const int c0 = half * 128;
    int r = m0 + row;
    const bool valid = r < p.R;
    if (!valid) r = p.R - 1;
    const int ci = S.shared ? 0 : (p.row0 + r) / p.div;
    const float* xt = S.XT + ((size_t)(ci / TM) * STEPS) * XT_COLS * TM;
    const int ln4 = (ci % TM) * 4;
    const float* st = S.scoresT + ((size_t)((m0 + row) / TM) * STEPS) * TM + row;   // this CTA's tile
    const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
    const uint32_t a_row_off = (uint32_t)(row / 8) * A_SBO + (uint32_t)(row % 8) * 16;
    float h[128];
#pragma unroll
    for (int i = 0; i < 128; ++i) h[i] = 0.f;
    // h0 = 0 into the A operand
#pragma unroll
    for (int g = 0; g < 16; ++g) {
      uint32_t off = a_row_off + (uint32_t)((c0 + g * 8) / 8) * LBO;
      *reinterpret_cast<uint4*>(sAhi + off) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sAlo + off) = make_uint4(0, 0, 0, 0);
    }
    proxy_fence();
    mbar_arrive(&bar_h);

    for (int t = 0; t < STEPS; ++t) {
      const uint32_t par = t & 1;
      const float* xs = xt + (size_t)t * XT_COLS * TM;
      const float one_minus_s = 1.0f - __ldg(st + (size_t)t * TM);
      // Pull the NEXT step's input halves (768 columns x 128 lanes x 4 B = 3072 lines) from HBM into L2
      // now: each thread touches 12 lines; the demand loads one step later then see L2 latency.
      if (t + 1 < STEPS) {
        const float* xn = S.XT + (((size_t)(ci / TM) * STEPS + (t + 1)) * XT_COLS) * TM;
#pragma unroll
        for (int i = 0; i < 12; ++i)
          asm volatile("prefetch.global.L2 [%0];" :: "l"(xn + (size_t)(i * 256 + (tid & 255)) * 32));
      }
      // Each phase walks its 128 columns in 8 chunks of 16 with a 2-deep software pipeline: the TMEM
      // load and the coalesced X loads of chunk ch+1 are in flight while chunk ch is computed.
#define R4_LOADX(dst, colbase) load_x16(dst, xs, (colbase), ln4)
      // ---- phase U: u' = (1 - s) sigmoid(acc_u + Xu) -> back into TMEM ----
      {
        float x[2][16], a[2][16];
        R4_LOADX(x[0], HID + c0);
        mbar_wait(&bar_u, par);
        tc_fence_after();
        tmem_ld16(tlane + HID + c0, a[0]);
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const int cur = ch & 1, nxt = cur ^ 1;
          tmem_wait_ld();
          if (ch < 7) { R4_LOADX(x[nxt], HID + c0 + (ch + 1) * 16); tmem_ld16(tlane + HID + c0 + (ch + 1) * 16, a[nxt]); }
#pragma unroll
          for (int j = 0; j < 16; ++j) a[cur][j] = one_minus_s * fast_sigmoid(a[cur][j] + x[cur][j]);
          tmem_st16(tlane + HID + c0 + ch * 16, a[cur]);
        }
        tmem_wait_st();
      }
      // ---- phase R: r*h -> A operand ----
      {
        float x[2][16], a[2][16];
        R4_LOADX(x[0], c0);
        mbar_wait(&bar_r, par);
        tc_fence_after();
        tmem_ld16(tlane + c0, a[0]);
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const int cur = ch & 1, nxt = cur ^ 1;
          tmem_wait_ld();
          if (ch < 7) { R4_LOADX(x[nxt], c0 + (ch + 1) * 16); tmem_ld16(tlane + c0 + (ch + 1) * 16, a[nxt]); }
#pragma unroll
          for (int j = 0; j < 16; ++j) a[cur][j] = fast_sigmoid(a[cur][j] + x[cur][j]) * h[ch * 16 + j];
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 hi, lo;
            split8(a[cur] + g * 8, hi, lo);
            uint32_t off = a_row_off + (uint32_t)((c0 + ch * 16 + g * 8) / 8) * LBO;
            *reinterpret_cast<uint4*>(sAhi + off) = hi;
            *reinterpret_cast<uint4*>(sAlo + off) = lo;
          }
        }
      }
      tc_fence_before();
      proxy_fence();
      mbar_arrive(&bar_rh);
      // ---- phase C: c = tanh(acc_c + Xc); h <- u' h + (1 - u') c -> A operand ----
      {
        float x[2][16], a[2][16], u[2][16];
        R4_LOADX(x[0], 2 * HID + c0);
        mbar_wait(&bar_c, par);
        tc_fence_after();
        tmem_ld16(tlane + c0, a[0]);
        tmem_ld16(tlane + HID + c0, u[0]);
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const int cur = ch & 1, nxt = cur ^ 1;
          tmem_wait_ld();
          if (ch < 7) {
            R4_LOADX(x[nxt], 2 * HID + c0 + (ch + 1) * 16);
            tmem_ld16(tlane + c0 + (ch + 1) * 16, a[nxt]);
            tmem_ld16(tlane + HID + c0 + (ch + 1) * 16, u[nxt]);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float c = fast_tanh(a[cur][j] + x[cur][j]);
            float hn = fmaf(u[cur][j], h[ch * 16 + j] - c, c);          // u' h + (1 - u') c
            h[ch * 16 + j] = hn;
            a[cur][j] = hn;
          }
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 hi, lo;
            split8(a[cur] + g * 8, hi, lo);
            uint32_t off = a_row_off + (uint32_t)((c0 + ch * 16 + g * 8) / 8) * LBO;
            *reinterpret_cast<uint4*>(sAhi + off) = hi;
            *reinterpret_cast<uint4*>(sAlo + off) = lo;
          }
        }
      }
#undef R4_LOADX
      tc_fence_before();
      proxy_fence();
      mbar_arrive(&bar_h);
    }
    if (valid) {
      float* o = S.out + (size_t)(m0 + row) * p.out_ld + c0;
#pragma unroll
      for (int i = 0; i < 128; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(h[i], h[i + 1], h[i + 2], h[i + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(512));
}

// host: fp32 recurrent weights -> the pre-tiled bf16 hi/lo stream image (order u, r, c; 8 K blocks; hi, lo).
// Wg: [256][512] rows = h index, columns [r | u];  Wc: [256][256].
inline uint16_t host_bf16_bits(float x) {
  uint32_t u; memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)(u >> 16);
  u = (u + 0x7fffu + ((u >> 16) & 1u)) >> 16; return (uint16_t)u;
}
inline float host_bf16_val(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }
inline void build_weight_image(const float* Wg, const float* Wc, uint8_t* img) {
  for (int mat = 0; mat < 3; ++mat)
    for (int kb = 0; kb < NKB; ++kb)
      for (int sp = 0; sp < 2; ++sp) {
        uint8_t* st = img + (size_t)((mat * NKB + kb) * 2 + sp) * STAGE_BYTES;
        for (int n = 0; n < HID; ++n)
          for (int kk = 0; kk < KB; ++kk) {
            int k = kb * KB + kk;
            float w = mat == 0 ? Wg[(size_t)k * 2 * HID + HID + n] : (mat == 1 ? Wg[(size_t)k * 2 * HID + n] : Wc[(size_t)k * HID + n]);
            uint16_t hi = host_bf16_bits(w);
            uint16_t v = sp == 0 ? hi : host_bf16_bits(w - host_bf16_val(hi));
            memcpy(st + (n / 8) * B_SBO + (kk / 8) * LBO + (n % 8) * 16 + (kk % 8) * 2, &v, 2);
          }
      }
}

}  // namespace r4tc
