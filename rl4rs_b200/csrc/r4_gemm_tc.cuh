// r4_gemm_tc.cuh -- fp32-parity dense layers on tcgen05 (Keras Dense: nets/utils.py:50-53,
// dien.py:35; the input projections of GRU-1 / AUGRU / attention; the Embedding gather fused in).
//
//   C[M,N] = act(A[M,K] . W[K,N] + bias)      A, C, bias fp32 in HBM; W pre-tiled bf16 hi/lo image
//
// Numerics: fp32-exact in practice.  Every operand is split THREE ways, x = hi + mid + lo (bf16 each,
// 24 significand bits together), and the six products whose magnitude exceeds 2^-24 of the result are
// issued (hh, hm, mh, hl, lh, mm), fp32 accumulation in TMEM.  These layers are a few % of the step, so
// the 2x tensor work over the 2-way split of the recurrence buys parity margin for free.
// One CTA = a 128-row x (<=256)-column tile.
// Warp roles: 0-3 A producers (fp32 -> bf16 hi/lo core matrices, coalesced 32-byte reads per thread)
// and, at the end, the epilogue (TMEM -> bias/act -> global); 4 MMA issuer; 5 TMA producer for W.
// Ring: 3 stages x (A hi/mid/lo 24 KB + W hi/mid/lo <=48 KB), K block = 32.
#pragma once
#include "r4_augru_tc.cuh"

namespace r4tc {

constexpr int G_BM = 128, G_BK = 32, G_NST = 3, G_BNMAX = 256, G_SPLIT = 3;
constexpr int G_A_STAGE = G_BM * G_BK * 2;                 // 8 KB per split
constexpr int G_B_STAGE = G_BNMAX * G_BK * 2;              // 16 KB per split (max)
constexpr int G_STAGE_BYTES = G_SPLIT * G_A_STAGE + G_SPLIT * G_B_STAGE;   // 72 KB
constexpr int G_SMEM_BYTES = G_NST * G_STAGE_BYTES + 1024;
constexpr int G_THREADS = 192;
constexpr int G_A_SBO = (G_BK / 8) * 128;                  // 512
constexpr int G_B_SBO = (G_BK / 8) * 128;                  // 512

// host-side description of a pre-tiled weight
struct GemmImage {
  const uint8_t* img = nullptr;   // device
  int K = 0, N = 0, kblocks = 0;  // kblocks = ceil(K/32) (zero padded)
};

// image layout: n-tile nt (256 columns, last one narrower), K block kb, split {hi, lo}: [bn x 32] core matrices
inline size_t gemm_image_bytes(int K, int N) {
  int kb = (K + G_BK - 1) / G_BK;
  size_t tot = 0;
  for (int n0 = 0; n0 < N; n0 += G_BNMAX) tot += (size_t)kb * G_SPLIT * std::min(G_BNMAX, N - n0) * G_BK * 2;
  return tot;
}
inline void build_gemm_image(const float* W /*[K][N]*/, int K, int N, uint8_t* img) {
  int kbn = (K + G_BK - 1) / G_BK;
  size_t off = 0;
  for (int n0 = 0; n0 < N; n0 += G_BNMAX) {
    int bn = std::min(G_BNMAX, N - n0);
    for (int kb = 0; kb < kbn; ++kb)
      for (int sp = 0; sp < G_SPLIT; ++sp) {
        uint8_t* st = img + off;
        for (int n = 0; n < bn; ++n)
          for (int kk = 0; kk < G_BK; ++kk) {
            int k = kb * G_BK + kk;
            float w = k < K ? W[(size_t)k * N + n0 + n] : 0.f;
            uint16_t hi = host_bf16_bits(w);
            float r1 = w - host_bf16_val(hi);
            uint16_t mid = host_bf16_bits(r1);
            uint16_t lo = host_bf16_bits(r1 - host_bf16_val(mid));
            uint16_t v = sp == 0 ? hi : (sp == 1 ? mid : lo);
            memcpy(st + (n / 8) * G_B_SBO + (kk / 8) * LBO + (n % 8) * 16 + (kk % 8) * 2, &v, 2);
          }
        off += (size_t)bn * G_BK * 2;
      }
  }
}

struct GemmTcParams {
  const float* A; int lda; const int32_t* gather;
  const uint8_t* Wimg; const float* bias; float* C; int ldc;
  int M, N, K, act;     // act: 0 none, 1 ELU
  // "sequence" mode (tm_ns > 0): the M = tm_ns * 64 rows are (sequence n, step t) pairs enumerated t-major
  // (m = t * tm_ns + n; source row n * 64 + t), and the result is written in the lane-major tile layout the
  // recurrent kernels read: columns < ldT -> outT[((nabs/128) * 64 + t) * ldT + col) * 128 + nabs % 128],
  // columns >= ldT -> outK[(nabs * 64 + t) * (N - ldT) + col - ldT], nabs = cr_base + n.  Consecutive threads are
  // consecutive sequences of one step, so both reads and transposed writes stay coalesced.
  int tm_ns, cr_base, ldT; float* outT; float* outK;
};

__global__ void __launch_bounds__(G_THREADS, 1) k_gemm_tc(GemmTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_a[G_NST], bar_b[G_NST], bar_empty[G_NST], bar_done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * G_BM;
  const int n0 = blockIdx.y * G_BNMAX;
  const int bn = min(G_BNMAX, p.N - n0);
  const int kbn = (p.K + G_BK - 1) / G_BK;
  // start of this n-tile in the image: full tiles before it are 256 wide
  const uint8_t* wtile = p.Wimg + (size_t)blockIdx.y * kbn * G_SPLIT * G_BNMAX * G_BK * 2;
  const uint32_t b_split_bytes = (uint32_t)bn * G_BK * 2;

  if (tid == 0) {
    for (int i = 0; i < G_NST; ++i) { mbar_init(&bar_a[i], 128); mbar_init(&bar_b[i], 1); mbar_init(&bar_empty[i], 1); }
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  if (warp == 5) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < kbn; ++kb) {
        mbar_wait(&bar_empty[stage], phase ^ 1);
        uint8_t* dst = smem + stage * G_STAGE_BYTES + G_SPLIT * G_A_STAGE;
        mbar_expect_tx(&bar_b[stage], G_SPLIT * b_split_bytes);
#pragma unroll
        for (int sp = 0; sp < G_SPLIT; ++sp)
          bulk_g2s(dst + sp * G_B_STAGE, wtile + ((size_t)kb * G_SPLIT + sp) * b_split_bytes, b_split_bytes, &bar_b[stage]);
        if (++stage == G_NST) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 4) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(G_BM, bn);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < kbn; ++kb) {
        mbar_wait(&bar_a[stage], phase);
        mbar_wait(&bar_b[stage], phase);
        tc_fence_after();
        uint32_t sa = smem_u32(smem + stage * G_STAGE_BYTES);
        uint32_t sb = sa + G_SPLIT * G_A_STAGE;
#pragma unroll
        for (int j = 0; j < G_BK / 16; ++j) {
          uint64_t a[3], b[3];
#pragma unroll
          for (int sp = 0; sp < 3; ++sp) {
            a[sp] = make_desc(sa + sp * G_A_STAGE + j * 2 * LBO, LBO, G_A_SBO);
            b[sp] = make_desc(sb + sp * G_B_STAGE + j * 2 * LBO, LBO, G_B_SBO);
          }
          // smallest products first so they are not absorbed by a large partial sum
          mma_bf16(tbase, a[1], b[1], idesc, (kb | j) ? 1u : 0u);   // mid*mid
          mma_bf16(tbase, a[0], b[2], idesc, 1u);                   // hi*lo
          mma_bf16(tbase, a[2], b[0], idesc, 1u);                   // lo*hi
          mma_bf16(tbase, a[0], b[1], idesc, 1u);                   // hi*mid
          mma_bf16(tbase, a[1], b[0], idesc, 1u);                   // mid*hi
          mma_bf16(tbase, a[0], b[0], idesc, 1u);                   // hi*hi
        }
        umma_commit(&bar_empty[stage]);
        if (++stage == G_NST) { stage = 0; phase ^= 1; }
      }
      umma_commit(&bar_done);
    }
  } else {
    // ---- A producers: thread (r8 = tid/4, kc = tid%4) converts 8 consecutive K values of 4 rows per K block ----
    const int kc = tid & 3;
    const float* arow[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      int m = m0 + it * 32 + (tid >> 2);
      if (m >= p.M) m = p.M - 1;
      if (p.tm_ns > 0) m = (m % p.tm_ns) * 64 + (m / p.tm_ns);
      size_t src = p.gather ? (size_t)p.gather[m] : (size_t)m;
      arow[it] = p.A + src * p.lda + kc * 8;
    }
    int stage = 0; uint32_t phase = 0;
    float v[4][8], nx[4][8];
    auto load_block = [&](int kb, float (*dst)[8]) {
      const int k = kb * G_BK + kc * 8;
      const bool kok = kb < kbn && k + 8 <= p.K;        // K % 8 == 0 is required
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        if (kok) {
          float4 x0 = __ldg(reinterpret_cast<const float4*>(arow[it] + kb * G_BK));
          float4 x1 = __ldg(reinterpret_cast<const float4*>(arow[it] + kb * G_BK + 4));
          dst[it][0] = x0.x; dst[it][1] = x0.y; dst[it][2] = x0.z; dst[it][3] = x0.w;
          dst[it][4] = x1.x; dst[it][5] = x1.y; dst[it][6] = x1.z; dst[it][7] = x1.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[it][j] = 0.f;
        }
      }
    };
    load_block(0, v);
    for (int kb = 0; kb < kbn; ++kb) {
      load_block(kb + 1, nx);                           // next block's global loads fly during this one
      mbar_wait(&bar_empty[stage], phase ^ 1);
      uint8_t* sa = smem + stage * G_STAGE_BYTES;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        int row = it * 32 + (tid >> 2);
        uint4 hi, mid, lo;
        float r1[8];
        split8(v[it], hi, mid);                         // hi and bf16(x - hi)
        {
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&hi);
          const __nv_bfloat162* m2 = reinterpret_cast<const __nv_bfloat162*>(&mid);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            r1[2 * j] = (v[it][2 * j] - __bfloat162float(h2[j].x)) - __bfloat162float(m2[j].x);
            r1[2 * j + 1] = (v[it][2 * j + 1] - __bfloat162float(h2[j].y)) - __bfloat162float(m2[j].y);
          }
          uint32_t l[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 ll = __floats2bfloat162_rn(r1[2 * j], r1[2 * j + 1]);
            l[j] = *reinterpret_cast<uint32_t*>(&ll);
          }
          lo = make_uint4(l[0], l[1], l[2], l[3]);
        }
        uint32_t off = (uint32_t)(row / 8) * G_A_SBO + (uint32_t)kc * LBO + (uint32_t)(row % 8) * 16;
        *reinterpret_cast<uint4*>(sa + off) = hi;
        *reinterpret_cast<uint4*>(sa + G_A_STAGE + off) = mid;
        *reinterpret_cast<uint4*>(sa + 2 * G_A_STAGE + off) = lo;
      }
      proxy_fence();
      mbar_arrive(&bar_a[stage]);
      if (++stage == G_NST) { stage = 0; phase ^= 1; }
#pragma unroll
      for (int it = 0; it < 4; ++it)
#pragma unroll
        for (int j = 0; j < 8; ++j) v[it][j] = nx[it][j];
    }
    // ---- epilogue: thread = row ----
    mbar_wait(&bar_done, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    const int m = m0 + row;
    const uint32_t tlane = tbase + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < bn; c += 16) {
      float a[16];
      tmem_ld16(tlane + c, a);
      tmem_wait_ld();
      if (m < p.M && p.tm_ns > 0) {
        const int t = m / p.tm_ns, nabs = p.cr_base + m % p.tm_ns;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = n0 + c + j;
          float r = a[j] + (p.bias ? __ldg(p.bias + col) : 0.f);
          if (p.act == 1) r = r > 0.f ? r : expm1f(r);
          if (col < p.ldT) p.outT[(((size_t)(nabs / TM) * STEPS + t) * p.ldT + col) * TM + (nabs % TM)] = r;
          else p.outK[((size_t)nabs * STEPS + t) * (p.N - p.ldT) + (col - p.ldT)] = r;
        }
      } else if (m < p.M) {
        float* o = p.C + (size_t)m * p.ldc + n0 + c;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 b = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 r = make_float4(a[j] + b.x, a[j + 1] + b.y, a[j + 2] + b.z, a[j + 3] + b.w);
          if (p.act == 1) {
            r.x = r.x > 0.f ? r.x : expm1f(r.x); r.y = r.y > 0.f ? r.y : expm1f(r.y);
            r.z = r.z > 0.f ? r.z : expm1f(r.z); r.w = r.w > 0.f ? r.w : expm1f(r.w);
          }
          *reinterpret_cast<float4*>(o + j) = r;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(256));
}

}  // namespace r4tc
