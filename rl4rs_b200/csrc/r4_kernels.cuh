// r4_kernels.cuh -- sm_100a kernels of the RL4RS hot path.
//
// Every kernel cites the reference operation it replaces.  Arithmetic is fp32 with precise
// expf/tanhf/expm1f (parity: 1e-4 relative against the f32 CPU oracle); reward and kNN scores are
// f64 like the reference (slate.py:188,302).  Layouts are described in DESIGN.md section 3.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace r4 {

constexpr int EMB = 128;        // emb_size
constexpr int MAXLEN = 64;      // maxlen
constexpr int NCAT = 21;        // category_feature_num
constexpr int NDENSE = 432;     // dense_feature_num
constexpr int VEC = 40;         // item vector width (item_info.csv)
constexpr int UD = 32;          // user dense floats  (user_protrait[10:])
constexpr int UC = 10;          // user categorical ids (user_protrait[:10])
constexpr int PAGE = 9;         // page_items
constexpr int HU = 128;         // hidden_units (dense tower)
constexpr int AH1 = 64;         // attention MLP hidden 1
constexpr int AH2 = 16;         // attention MLP hidden 2
constexpr int AUH = 256;        // AUGRU hidden (emb_size * 2, nets/utils.py:123)
constexpr int OBSD = 256;       // simulator_obs width (dien.py:35)
constexpr int ALLF = 2 * AUH + HU + EMB + NCAT * EMB;   // 3456, dien.py:34
constexpr int XIN_LD = 3 * EMB;                          // GRU-1 input projection: [r|u|c] = 384
constexpr int XK_LD = 2 * AUH + AUH + AH1;               // AUGRU input proj [r|u (512) | c (256) | key (64)] = 832
constexpr int XK_C = 2 * AUH;                            // offset of candidate part
constexpr int XK_K = 3 * AUH;                            // offset of attention key part
constexpr int MAX_WORDS = 16;                            // action_size <= 512

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float eluf_(float x) { return x > 0.f ? x : expm1f(x); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// location_mask rows (slate.py:60-64): layer 0 = ids 1..39, 1 = 40..147, 2 = 148.., 3 = {0}
__device__ __forceinline__ bool loc_allowed(int layer, int a) {
  switch (layer) {
    case 0: return a >= 1 && a < 40;
    case 1: return a >= 40 && a < 148;
    case 2: return a >= 148;
    default: return a == 0;
  }
}

// ------------------------------------------------------------------------------------------
// K1 reset: SlateState.__init__ (slate.py:16-19): prev_actions = 0, masks = 1
// ------------------------------------------------------------------------------------------
__global__ void k_init_state(int B, int T, int A, int words, int32_t* prev_actions, uint32_t* amask,
                             uint8_t* sflag, uint8_t* mask_out /*[B,A] or null*/, int layer0) {
  int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (b >= B) return;
  for (int t = lane; t < T; t += 32) prev_actions[(size_t)b * T + t] = 0;
  for (int w = lane; w < words; w += 32) {
    int lo = w * 32;
    uint32_t bits = (A - lo >= 32) ? 0xffffffffu : ((A - lo) <= 0 ? 0u : ((1u << (A - lo)) - 1u));
    amask[(size_t)b * words + w] = bits;
  }
  if (lane == 0) sflag[b] = 0;
  if (mask_out)
    for (int a = lane; a < A; a += 32) mask_out[(size_t)b * A + a] = loc_allowed(layer0, a) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// K1 + K2  SlateState.act (slate.py:193-202) / SeqSlateState.act (seqslate.py:92-102,124-126)
// one warp per env row.  Conti mode resolves the action by masked kNN in f64 (slate.py:186-191):
// fill -2^31, first index wins ties.
// ------------------------------------------------------------------------------------------
struct ActParams {
  int B, T, P, A, words, seq, conti, emb_dim, cur_steps, act_f64;
};

__global__ void k_act(ActParams p, const void* __restrict__ action, const double* __restrict__ action_emb,
                      const uint8_t* __restrict__ special, int32_t* __restrict__ prev_actions,
                      uint32_t* __restrict__ amask, uint8_t* __restrict__ sflag,
                      int32_t* __restrict__ chosen_out, uint8_t* __restrict__ mask_out) {
  int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (b >= p.B) return;
  uint32_t* am = amask + (size_t)b * p.words;
  int32_t* pa = prev_actions + (size_t)b * p.T;
  int cur = p.cur_steps;
  int a_sel;
  if (p.conti) {
    int layer = p.seq ? (cur % p.P) / 3 : cur / 3;          // pre-act layer (Q4)
    bool sf = sflag[b] != 0;
    double best = -1e308;
    int besti = 0x7fffffff;
    const float* af = reinterpret_cast<const float*>(action) + (size_t)b * p.emb_dim;
    const double* ad = reinterpret_cast<const double*>(action) + (size_t)b * p.emb_dim;
    for (int a = lane; a < p.A; a += 32) {
      bool ok = ((am[a >> 5] >> (a & 31)) & 1u) && loc_allowed(layer, a) && !(sf && special[a]);
      double s;
      if (ok) {
        s = 0.0;
        const double* e = action_emb + (size_t)a * p.emb_dim;
        if (p.act_f64) { for (int k = 0; k < p.emb_dim; ++k) s += ad[k] * e[k]; }
        else { for (int k = 0; k < p.emb_dim; ++k) s += (double)af[k] * e[k]; }
      } else {
        s = -2147483648.0;
      }
      if (s > best) { best = s; besti = a; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      double s2 = __shfl_xor_sync(0xffffffffu, best, o);
      int i2 = __shfl_xor_sync(0xffffffffu, besti, o);
      if (s2 > best || (s2 == best && i2 < besti)) { best = s2; besti = i2; }
    }
    a_sel = besti;
  } else {
    a_sel = reinterpret_cast<const int32_t*>(action)[b];
    a_sel = a_sel < 0 ? 0 : (a_sel >= p.A ? p.A - 1 : a_sel);   // memory safety; host validates
  }
  if (lane == 0) {
    pa[cur] = a_sel;                                           // slate.py:198
    am[a_sel >> 5] &= ~(1u << (a_sel & 31));                   // slate.py:199 (also for a == 0, Q5)
    if (chosen_out) chosen_out[b] = a_sel;
  }
  __syncwarp();
  // slate.py:200-202: special mask closes once the WHOLE history holds a special item (Q6, Q8)
  bool any_sp = false;
  for (int t = lane; t < p.T; t += 32) any_sp |= special[pa[t]] != 0;
  any_sp = __any_sync(0xffffffffu, any_sp);
  int nxt = cur + 1;
  bool page_reset = p.seq && (nxt % p.P == 0);                 // seqslate.py:124-126
  if (page_reset) {
    any_sp = false;
    for (int w = lane; w < p.words; w += 32) {
      int lo = w * 32;
      am[w] = (p.A - lo >= 32) ? 0xffffffffu : ((1u << (p.A - lo)) - 1u);
    }
    __syncwarp();
  }
  if (lane == 0) sflag[b] = any_sp ? 1 : 0;
  if (mask_out) {                                              // slate.py:93-97 / seqslate.py:15-17
    int layer = p.seq ? (nxt % p.P) / 3 : nxt / 3;
    for (int a = lane; a < p.A; a += 32) {
      bool ok = ((am[a >> 5] >> (a & 31)) & 1u) && loc_allowed(layer, a) && !(any_sp && special[a]);
      mask_out[(size_t)b * p.A + a] = ok ? 1 : 0;
    }
  }
}

// unmasked kNN (slate.py:180-184), one warp per query row
__global__ void k_knn_plain(int n, int A, int emb_dim, int act_f64, const void* __restrict__ action,
                            const double* __restrict__ action_emb, int32_t* __restrict__ out) {
  int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (b >= n) return;
  const float* af = reinterpret_cast<const float*>(action) + (size_t)b * emb_dim;
  const double* ad = reinterpret_cast<const double*>(action) + (size_t)b * emb_dim;
  double best = -1e308;
  int besti = 0x7fffffff;
  for (int a = lane; a < A; a += 32) {
    double s = 0.0;
    const double* e = action_emb + (size_t)a * emb_dim;
    if (act_f64) { for (int k = 0; k < emb_dim; ++k) s += ad[k] * e[k]; }
    else { for (int k = 0; k < emb_dim; ++k) s += (double)af[k] * e[k]; }
    if (s > best) { best = s; besti = a; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    double s2 = __shfl_xor_sync(0xffffffffu, best, o);
    int i2 = __shfl_xor_sync(0xffffffffu, besti, o);
    if (s2 > best || (s2 == best && i2 < besti)) { best = s2; besti = i2; }
  }
  if (lane == 0) out[b] = besti;
}

// ------------------------------------------------------------------------------------------
// K3 feature assembly: the state rebuild of SlateState.act (slate.py:203-213) /
// SeqSlateState.act (seqslate.py:103-121) / get_complete_states (slate.py:117-131,
// seqslate.py:27-50) followed by FeatureUtil.feature_extraction's post-pad/truncate
// (datautil.py:52-65).  One warp per feature row; rows r0.. of the pass.
//   mode 0: initial state (slate.py:72-80)      rows = B
//   mode 1: state after act at step `step`      rows = B
//   mode 2: reward rows, env-row major          rows = B * rpe (rpe = T for Slate, P for SeqSlate)
// ------------------------------------------------------------------------------------------
struct AsmParams {
  int mode, B, T, P, seq, step, rpe, row0, nrows;
};

__global__ void k_assemble(AsmParams p, const int32_t* __restrict__ row_idx,
                           const int32_t* __restrict__ log_cat, const float* __restrict__ log_dense,
                           const float* __restrict__ item_vec, const int32_t* __restrict__ prev_actions,
                           int32_t* __restrict__ cat_out, float* __restrict__ dense_out) {
  int rl = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (rl >= p.nrows) return;
  int r = p.row0 + rl;
  int b = (p.mode == 2) ? r / p.rpe : r;
  int j = (p.mode == 2) ? r % p.rpe : 0;
  const int32_t* pa = prev_actions + (size_t)b * p.T;
  int64_t lr = row_idx[b];
  int W = 0, w0 = 0, sid = 0, cur_a = 0;
  if (p.mode != 0) {
    int st = (p.mode == 1) ? p.step : (p.seq ? p.step - p.P + j : j);   // step this row's state was built at
    if (p.seq) { w0 = st / p.P * p.P; W = p.P; sid = st / p.P + 1; }
    else { w0 = 0; W = p.T; sid = 1; }
    cur_a = pa[st];
  }
  int32_t* co = cat_out + (size_t)rl * NCAT;
  float* dn = dense_out + (size_t)rl * NDENSE;
  if (lane < NCAT) {
    int c = lane, v = 0;
    if (c < UC) v = log_cat[lr * UC + c];
    else if (p.mode != 0) {
      if (c == UC) v = sid;
      else if (c < UC + 1 + W) v = pa[w0 + c - UC - 1];
      else if (c == UC + 1 + W) v = cur_a;
    }
    co[c] = v;
  }
  for (int d = lane; d < NDENSE; d += 32) {
    float v = 0.f;
    if (d < UD) v = log_dense[lr * UD + d];
    else if (p.mode != 0) {
      int k = (d - UD) / VEC, e = (d - UD) % VEC;
      if (k < W) v = item_vec[(size_t)pa[w0 + k] * VEC + e];
      else if (k == W) v = item_vec[(size_t)cur_a * VEC + e];
    }
    dn[d] = v;
  }
}

// sequence ids: seq0 = user history (pre-padded in the log), seq1 = previous pages of the episode
// (seqslate.py:36-37,109-110 + pad_sequences at datautil.py:43-46) or zeros.
__global__ void k_seq_ids(int B, int T, int p0, const int32_t* __restrict__ row_idx,
                          const int32_t* __restrict__ log_seq, const int32_t* __restrict__ prev_actions,
                          int32_t* __restrict__ seq0 /*[B,64] or null*/, int32_t* __restrict__ seq1 /*[B,64] or null*/,
                          int32_t* __restrict__ seq_out /*[B,2,64] or null*/) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * MAXLEN) return;
  int b = i / MAXLEN, t = i % MAXLEN;
  int v0 = log_seq[(size_t)row_idx[b] * MAXLEN + t];
  int n = p0 < MAXLEN ? p0 : MAXLEN;                   // keep the last maxlen ids
  int v1 = 0;
  if (t >= MAXLEN - n) v1 = prev_actions[(size_t)b * T + (p0 - n) + (t - (MAXLEN - n))];
  if (seq0) seq0[i] = v0;
  if (seq1) seq1[i] = v1;
  if (seq_out) { seq_out[(size_t)b * 2 * MAXLEN + t] = v0; seq_out[(size_t)b * 2 * MAXLEN + MAXLEN + t] = v1; }
}

// ------------------------------------------------------------------------------------------
// Generic fp32 GEMM  C[M,N] = act(A[M,K] @ W[K,N] + bias)   (Keras Dense: nets/utils.py:50-53,
// dien.py:35) with optional row gather on A (Embedding lookup fused into the input projection,
// nets/utils.py:113).  128x128x8 tiles, 8x8 register micro-tiles.  N % 4 == 0, K % 8 == 0.
// ------------------------------------------------------------------------------------------
template <int ACT>   // 0 none, 1 ELU
__global__ void __launch_bounds__(256) k_gemm(int M, int N, int K, const float* __restrict__ A, int lda,
                                              const int32_t* __restrict__ gather, const float* __restrict__ W,
                                              const float* __restrict__ bias, float* __restrict__ C, int ldc) {
  __shared__ __align__(16) float As[8][128];
  __shared__ __align__(16) float Bs[8][128];
  int tid = threadIdx.x;
  int m0 = blockIdx.x * 128, n0 = blockIdx.y * 128;
  int tx = tid & 15, ty = tid >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  int arow = tid >> 1, akq = (tid & 1) * 4;
  int gm = m0 + arow;
  const float* aptr = nullptr;
  if (gm < M) {
    int64_t src = gather ? (int64_t)gather[gm] : (int64_t)gm;
    aptr = A + src * lda + akq;
  }
  int bk = tid >> 5, bn = (tid & 31) * 4;
  bool bok = (n0 + bn) < N;
  for (int k0 = 0; k0 < K; k0 += 8) {
    float4 av = aptr ? ldg4(aptr + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 bv = bok ? ldg4(W + (size_t)(k0 + bk) * N + n0 + bn) : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    As[akq + 0][arow] = av.x; As[akq + 1][arow] = av.y; As[akq + 2][arow] = av.z; As[akq + 3][arow] = av.w;
    *reinterpret_cast<float4*>(&Bs[bk][bn]) = bv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int n = n0 + (h ? 64 : 0) + tx * 4;
      if (n >= N) continue;
      float4 bb = bias ? ldg4(bias + n) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 v;
      v.x = acc[i][h * 4 + 0] + bb.x; v.y = acc[i][h * 4 + 1] + bb.y;
      v.z = acc[i][h * 4 + 2] + bb.z; v.w = acc[i][h * 4 + 3] + bb.w;
      if (ACT == 1) { v.x = eluf_(v.x); v.y = eluf_(v.y); v.z = eluf_(v.z); v.w = eluf_(v.w); }
      *reinterpret_cast<float4*>(C + (size_t)m * ldc + n) = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// K7 / K9 recurrence.  TF1 GRUCell (HID=128, nets/utils.py:120) and deepctr VecAttGRUCell
// (HID=256, AUGRU, nets/utils.py:123-124), h0 = 0, all 64 steps:
//   [r,u] = sigmoid(x_t Wgx + h Wgh + bg);  c = tanh(x_t Wcx + (r*h) Wch + bc)
//   AUGRU: u <- (1 - score_t) * u;          h <- u*h + (1-u)*c
// The input halves x_t Wgx + bg and x_t Wcx + bc are precomputed (X, row stride xld, per cached
// sequence); this kernel does the h-dependent halves.  One CTA = M rows x 64 steps; h lives in
// shared memory transposed (hT[k][row]) so a warp's operand reads are broadcasts; each thread owns
// an 8-row x 4-column tile of r, u and c, so the gate algebra never leaves registers.
// grid.y selects the sequence (two independent weight sets).
// ------------------------------------------------------------------------------------------
struct RecurSeq {
  const float* X;        // [n_cached, 64, xld]
  const float* Wgh;      // [HID, 2*HID]
  const float* Wch;      // [HID, HID]
  const float* scores;   // [R, 64] (AUGRU) or null
  float* out;            // STORE_ALL: [R, 64, HID]; else final state rows with stride out_ld
  int shared;            // 1: every row uses cached sequence 0
};
struct RecurParams {
  RecurSeq s[2];
  int R, row0, div, xld, xoff_g, xoff_c, out_ld;
};

template <int HID, bool AUGRU_, bool STORE_ALL>
__global__ void __launch_bounds__(256, 2) k_recur(RecurParams p) {
  constexpr int NCG = HID / 4;          // column groups of 4
  constexpr int NRG = 256 / NCG;        // row groups of 8
  constexpr int M = 8 * NRG;            // rows per CTA: 32 (HID=256) / 64 (HID=128)
  extern __shared__ __align__(16) float smem[];
  float* hT = smem;                     // [HID][M]
  float* rhT = smem + HID * M;          // [HID][M]
  float* sc = rhT + HID * M;            // [M][64] scores (AUGRU)
  __shared__ int cidx_s[M];             // cached-sequence index of each row
  const RecurSeq& S = p.s[blockIdx.y];
  int tid = threadIdx.x;
  int tx = tid % NCG, ty = tid / NCG;
  int m0 = blockIdx.x * M;
  int rows_here = p.R - m0 < M ? p.R - m0 : M;
  if (tid < M) {
    int r = m0 + tid;
    if (r >= p.R) r = p.R - 1;
    cidx_s[tid] = S.shared ? 0 : (p.row0 + r) / p.div;
  }
  const size_t seq_stride = (size_t)MAXLEN * p.xld;
  for (int i = tid; i < HID * M; i += 256) hT[i] = 0.f;
  if (AUGRU_) {
    for (int i = tid; i < M * MAXLEN; i += 256) {
      int rr = i / MAXLEN;
      int r = m0 + rr; if (r >= p.R) r = p.R - 1;
      sc[i] = S.scores[(size_t)r * MAXLEN + (i % MAXLEN)];
    }
  }
  float hreg[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) hreg[i][c] = 0.f;
  __syncthreads();

  for (int t = 0; t < MAXLEN; ++t) {
    float ar[8][4], au[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float* xp = S.X + cidx_s[ty * 8 + i] * seq_stride + (size_t)t * p.xld + p.xoff_g + tx * 4;
      float4 vr = ldg4(xp), vu = ldg4(xp + HID);
      ar[i][0] = vr.x; ar[i][1] = vr.y; ar[i][2] = vr.z; ar[i][3] = vr.w;
      au[i][0] = vu.x; au[i][1] = vu.y; au[i][2] = vu.z; au[i][3] = vu.w;
    }
    const float* wg = S.Wgh + tx * 4;
#pragma unroll 4
    for (int k = 0; k < HID; ++k) {
      float4 wr = ldg4(wg + (size_t)k * 2 * HID);
      float4 wu = ldg4(wg + (size_t)k * 2 * HID + HID);
      float4 h0 = *reinterpret_cast<const float4*>(&hT[k * M + ty * 8]);
      float4 h1 = *reinterpret_cast<const float4*>(&hT[k * M + ty * 8 + 4]);
      float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        ar[i][0] = fmaf(hv[i], wr.x, ar[i][0]); ar[i][1] = fmaf(hv[i], wr.y, ar[i][1]);
        ar[i][2] = fmaf(hv[i], wr.z, ar[i][2]); ar[i][3] = fmaf(hv[i], wr.w, ar[i][3]);
        au[i][0] = fmaf(hv[i], wu.x, au[i][0]); au[i][1] = fmaf(hv[i], wu.y, au[i][1]);
        au[i][2] = fmaf(hv[i], wu.z, au[i][2]); au[i][3] = fmaf(hv[i], wu.w, au[i][3]);
      }
    }
    // r * h -> rhT ; keep u
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float4 lo, hi;
      lo.x = sigmoidf_(ar[0][c]) * hreg[0][c]; lo.y = sigmoidf_(ar[1][c]) * hreg[1][c];
      lo.z = sigmoidf_(ar[2][c]) * hreg[2][c]; lo.w = sigmoidf_(ar[3][c]) * hreg[3][c];
      hi.x = sigmoidf_(ar[4][c]) * hreg[4][c]; hi.y = sigmoidf_(ar[5][c]) * hreg[5][c];
      hi.z = sigmoidf_(ar[6][c]) * hreg[6][c]; hi.w = sigmoidf_(ar[7][c]) * hreg[7][c];
      *reinterpret_cast<float4*>(&rhT[(tx * 4 + c) * M + ty * 8]) = lo;
      *reinterpret_cast<float4*>(&rhT[(tx * 4 + c) * M + ty * 8 + 4]) = hi;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) au[i][c] = sigmoidf_(au[i][c]);
    __syncthreads();
    // candidate
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 vc = ldg4(S.X + cidx_s[ty * 8 + i] * seq_stride + (size_t)t * p.xld + p.xoff_c + tx * 4);
      ar[i][0] = vc.x; ar[i][1] = vc.y; ar[i][2] = vc.z; ar[i][3] = vc.w;
    }
    const float* wc = S.Wch + tx * 4;
#pragma unroll 4
    for (int k = 0; k < HID; ++k) {
      float4 w = ldg4(wc + (size_t)k * HID);
      float4 h0 = *reinterpret_cast<const float4*>(&rhT[k * M + ty * 8]);
      float4 h1 = *reinterpret_cast<const float4*>(&rhT[k * M + ty * 8 + 4]);
      float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        ar[i][0] = fmaf(hv[i], w.x, ar[i][0]); ar[i][1] = fmaf(hv[i], w.y, ar[i][1]);
        ar[i][2] = fmaf(hv[i], w.z, ar[i][2]); ar[i][3] = fmaf(hv[i], w.w, ar[i][3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float scale = 1.f;
      if (AUGRU_) scale = 1.0f - sc[(ty * 8 + i) * MAXLEN + t];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float u = au[i][c];
        if (AUGRU_) u = scale * u;
        float cand = tanhf(ar[i][c]);
        hreg[i][c] = u * hreg[i][c] + (1.0f - u) * cand;
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      *reinterpret_cast<float4*>(&hT[(tx * 4 + c) * M + ty * 8]) =
          make_float4(hreg[0][c], hreg[1][c], hreg[2][c], hreg[3][c]);
      *reinterpret_cast<float4*>(&hT[(tx * 4 + c) * M + ty * 8 + 4]) =
          make_float4(hreg[4][c], hreg[5][c], hreg[6][c], hreg[7][c]);
    }
    if (STORE_ALL) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int rr = ty * 8 + i;
        if (rr < rows_here)
          *reinterpret_cast<float4*>(S.out + ((size_t)(m0 + rr) * MAXLEN + t) * HID + tx * 4) =
              make_float4(hreg[i][0], hreg[i][1], hreg[i][2], hreg[i][3]);
      }
    }
    __syncthreads();
  }
  if (!STORE_ALL) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int rr = ty * 8 + i;
      if (rr < rows_here)
        *reinterpret_cast<float4*>(S.out + (size_t)(m0 + rr) * p.out_ld + tx * 4) =
            make_float4(hreg[i][0], hreg[i][1], hreg[i][2], hreg[i][3]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// K8 DIN local-activation scores (deepctr AttentionSequencePoolingLayer(att_hidden_units=(64,16),
// return_score=True), weight_normalization=False; nets/utils.py:114-115,121-122):
//   q = mean_j E_s[cat[-10:]]                                  (reduce_mean, keepdims)
//   z1_t = sigmoid([q, k_t, q-k_t, q*k_t] W1 + b1) = sigmoid(q(Wq+Wd) + b1 + k_t(Wk-Wd) + (q*k_t)Wp)
//   z2_t = sigmoid(z1_t W2 + b2);  score_t = z2_t . kv + b        (raw, unbounded)
// The key-only term k_t(Wk-Wd) is cached with the sequence (XK[..., 768:832]).
// One CTA per (row, sequence): a 64x64x128 tile product with A = q*H staged in shared memory.
// ------------------------------------------------------------------------------------------
struct ScoreSeq {
  const float* H;     // [n_cached, 64, 128]  GRU-1 outputs
  const float* XK;    // [n_cached, 64, 832]
  const float* Wqd;   // [128, 64]  Wq + Wd
  const float* Wp;    // [128, 64]
  const float* b1;    // [64]
  const float* W2;    // [64, 16]
  const float* b2;    // [16]
  const float* kv;    // [16]
  float bk;
  float* scores;      // [R, 64]
  int shared;
};
struct ScoreParams {
  ScoreSeq s[2];
  int R, row0, div;
  int transposed;   // 1: scores[(r/128)*64 + t][r%128] (lane-major tiles for the tensor-core AUGRU kernel)
};

constexpr int SC_ALD = 132;   // padded stride of the A tile
constexpr int SC_SMEM_FLOATS = 128 * 64 + 64 * SC_ALD + 128 + 64 * 4 + 64 + 64 * 16 + 16 + 16;

__global__ void __launch_bounds__(256) k_scores(ScoreParams p, const int32_t* __restrict__ cat,
                                                const float* __restrict__ emb_seq) {
  extern __shared__ __align__(16) float smem[];
  float* Wp_s = smem;                         // [128][64]
  float* A_s = Wp_s + 128 * 64;               // [64][132]  (later z1)
  float* q_s = A_s + 64 * SC_ALD;             // [128]
  float* qa_part = q_s + 128;                 // [4][64]
  float* qa_s = qa_part + 256;                // [64]
  float* W2_s = qa_s + 64;                    // [64][16]
  float* b2_s = W2_s + 64 * 16;               // [16]
  float* kv_s = b2_s + 16;                    // [16]
  const ScoreSeq& S = p.s[blockIdx.y];
  int r = blockIdx.x, tid = threadIdx.x;
  size_t ci = S.shared ? 0 : (size_t)((p.row0 + r) / p.div);
  const int32_t* crow = cat + (size_t)r * NCAT;
  if (tid < 128) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 10; ++j) s += __ldg(emb_seq + (size_t)crow[NCAT - 10 + j] * EMB + tid);
    q_s[tid] = s / 10.0f;
  }
  for (int i = tid; i < 128 * 64 / 4; i += 256)
    reinterpret_cast<float4*>(Wp_s)[i] = __ldg(reinterpret_cast<const float4*>(S.Wp) + i);
  for (int i = tid; i < 64 * 16; i += 256) W2_s[i] = __ldg(S.W2 + i);
  if (tid < 16) { b2_s[tid] = __ldg(S.b2 + tid); kv_s[tid] = __ldg(S.kv + tid); }
  __syncthreads();
  {  // qa = q (Wq + Wd) + b1, split over 4 k-ranges
    int j = tid & 63, part = tid >> 6;
    float s = 0.f;
#pragma unroll 8
    for (int k = part * 32; k < part * 32 + 32; ++k) s = fmaf(q_s[k], __ldg(S.Wqd + k * 64 + j), s);
    qa_part[part * 64 + j] = s;
  }
  const float* Hrow = S.H + ci * MAXLEN * EMB;
  for (int i = tid; i < 64 * 32; i += 256) {
    int t = i >> 5, k4 = (i & 31) * 4;
    float4 h = ldg4(Hrow + (size_t)t * EMB + k4);
    float4 q = *reinterpret_cast<const float4*>(&q_s[k4]);
    *reinterpret_cast<float4*>(&A_s[t * SC_ALD + k4]) = make_float4(h.x * q.x, h.y * q.y, h.z * q.z, h.w * q.w);
  }
  __syncthreads();
  if (tid < 64) qa_s[tid] = qa_part[tid] + qa_part[64 + tid] + qa_part[128 + tid] + qa_part[192 + tid] + __ldg(S.b1 + tid);
  __syncthreads();
  int tx = tid & 15, ty = tid >> 4;
  float z[4][4];
  {
    const float* kp = S.XK + ci * MAXLEN * XK_LD + XK_K + tx * 4;
    float4 qa = *reinterpret_cast<const float4*>(&qa_s[tx * 4]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 kk = ldg4(kp + (size_t)(ty * 4 + i) * XK_LD);
      z[i][0] = qa.x + kk.x; z[i][1] = qa.y + kk.y; z[i][2] = qa.z + kk.z; z[i][3] = qa.w + kk.w;
    }
  }
#pragma unroll 4
  for (int k = 0; k < 128; ++k) {
    float4 b = *reinterpret_cast<const float4*>(&Wp_s[k * 64 + tx * 4]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a = A_s[(ty * 4 + i) * SC_ALD + k];
      z[i][0] = fmaf(a, b.x, z[i][0]); z[i][1] = fmaf(a, b.y, z[i][1]);
      z[i][2] = fmaf(a, b.z, z[i][2]); z[i][3] = fmaf(a, b.w, z[i][3]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(&A_s[(ty * 4 + i) * SC_ALD + tx * 4]) =
        make_float4(sigmoidf_(z[i][0]), sigmoidf_(z[i][1]), sigmoidf_(z[i][2]), sigmoidf_(z[i][3]));
  __syncthreads();
  {
    int t = tid >> 2, jq = tid & 3;
    float o[4] = {b2_s[jq * 4], b2_s[jq * 4 + 1], b2_s[jq * 4 + 2], b2_s[jq * 4 + 3]};
#pragma unroll 8
    for (int j = 0; j < 64; ++j) {
      float a = A_s[t * SC_ALD + j];
      float4 w = *reinterpret_cast<const float4*>(&W2_s[j * 16 + jq * 4]);
      o[0] = fmaf(a, w.x, o[0]); o[1] = fmaf(a, w.y, o[1]); o[2] = fmaf(a, w.z, o[2]); o[3] = fmaf(a, w.w, o[3]);
    }
    float s = sigmoidf_(o[0]) * kv_s[jq * 4] + sigmoidf_(o[1]) * kv_s[jq * 4 + 1] +
              sigmoidf_(o[2]) * kv_s[jq * 4 + 2] + sigmoidf_(o[3]) * kv_s[jq * 4 + 3];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (jq == 0) {
      size_t o = p.transposed ? ((size_t)(r >> 7) * MAXLEN + t) * 128 + (r & 127) : (size_t)r * MAXLEN + t;
      S.scores[o] = s + S.bk;
    }
  }
}

// ------------------------------------------------------------------------------------------
// K4 + K5 category features (nets/utils.py:16-25): E_c gather, tf.keras.layers.Attention()
// (softmax(Q K^T) V, no scale, no mask), GlobalAveragePooling1D, and the flattened embeddings.
// One warp per row; softmax rows are reduced with warp shuffles.  Writes straight into the
// concatenated head input: allf[r, 640:768] = pooled attention, allf[r, 768:3456] = flatten.
// mean_t(P emb) is evaluated as (mean_t P) emb -- same value up to fp32 rounding order.
// ------------------------------------------------------------------------------------------
constexpr int CAT_LD = 132;
__global__ void __launch_bounds__(128) k_cat_attn(int R, const int32_t* __restrict__ cat,
                                                  const float* __restrict__ emb_cat, float* __restrict__ allf) {
  extern __shared__ __align__(16) float smem[];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int r = blockIdx.x * 4 + w;
  float* e = smem + w * (NCAT * CAT_LD + NCAT * 24);     // [21][132]
  float* Ssm = e + NCAT * CAT_LD;                         // [21][24]
  if (r >= R) return;
  const int32_t* crow = cat + (size_t)r * NCAT;
  float* out = allf + (size_t)r * ALLF;
  for (int j = 0; j < NCAT; ++j) {
    float4 v = ldg4(emb_cat + (size_t)crow[j] * EMB + lane * 4);
    *reinterpret_cast<float4*>(&e[j * CAT_LD + lane * 4]) = v;
    *reinterpret_cast<float4*>(out + 2 * AUH + HU + EMB + j * EMB + lane * 4) = v;   // Flatten
  }
  __syncwarp();
  // S[t][j] = <e_t, e_j>, symmetric: 231 pairs spread over lanes
  for (int pidx = lane; pidx < NCAT * (NCAT + 1) / 2; pidx += 32) {
    int t = 0, rem = pidx;
    while (rem >= NCAT - t) { rem -= NCAT - t; ++t; }
    int j = t + rem;
    float s = 0.f;
    const float* a = e + t * CAT_LD;
    const float* b = e + j * CAT_LD;
#pragma unroll 8
    for (int k = 0; k < EMB; k += 4) {
      float4 x = *reinterpret_cast<const float4*>(a + k);
      float4 y = *reinterpret_cast<const float4*>(b + k);
      s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
    }
    Ssm[t * 24 + j] = s;
    Ssm[j * 24 + t] = s;
  }
  __syncwarp();
  float wcol = 0.f;      // lane j accumulates sum_t P[t][j]
  for (int t = 0; t < NCAT; ++t) {
    float v = lane < NCAT ? Ssm[t * 24 + lane] : -INFINITY;
    float m = v;
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float ex = lane < NCAT ? expf(v - m) : 0.f;
    float sum = ex;
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    wcol += ex / sum;
  }
  wcol = wcol / (float)NCAT;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = 0; j < NCAT; ++j) {
    float wj = __shfl_sync(0xffffffffu, wcol, j);
    float4 v = *reinterpret_cast<const float4*>(&e[j * CAT_LD + lane * 4]);
    acc.x = fmaf(wj, v.x, acc.x); acc.y = fmaf(wj, v.y, acc.y);
    acc.z = fmaf(wj, v.z, acc.z); acc.w = fmaf(wj, v.w, acc.w);
  }
  *reinterpret_cast<float4*>(out + 2 * AUH + HU + lane * 4) = acc;
}

// ------------------------------------------------------------------------------------------
// K10 tail: probs = softmax(obs Wr + br) (dien.py:36); click prob = probs[:,1] (slate.py:298).
// One warp per row, warp-shuffle reduction of the two 256-long dots.
// ------------------------------------------------------------------------------------------
__global__ void k_reward_head(int R, const float* __restrict__ obs, const float* __restrict__ Wr,
                              const float* __restrict__ br, float* __restrict__ p1 /*[R] or null*/,
                              float* __restrict__ probs /*[R,2] or null*/) {
  int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (r >= R) return;
  float z0 = 0.f, z1 = 0.f;
  for (int k = lane; k < OBSD; k += 32) {
    float o = obs[(size_t)r * OBSD + k];
    z0 = fmaf(o, __ldg(Wr + k * 2), z0);
    z1 = fmaf(o, __ldg(Wr + k * 2 + 1), z1);
  }
  for (int o = 16; o > 0; o >>= 1) {
    z0 += __shfl_xor_sync(0xffffffffu, z0, o);
    z1 += __shfl_xor_sync(0xffffffffu, z1, o);
  }
  if (lane == 0) {
    z0 += br[0]; z1 += br[1];
    float m = fmaxf(z0, z1);
    float e0 = expf(z0 - m), e1 = expf(z1 - m);
    float s = e0 + e1;
    if (p1) p1[r] = e1 / s;
    if (probs) { probs[(size_t)r * 2] = e0 / s; probs[(size_t)r * 2 + 1] = e1 / s; }
  }
}

// ------------------------------------------------------------------------------------------
// K11 violation (slate.py:133-147 / seqslate.py:52-69) and reward (slate.py:293-308 /
// seqslate.py:147-157).  One thread per env row.
// ------------------------------------------------------------------------------------------
__device__ inline int violation_row(const int32_t* pa, int cur, int P, int seq, const uint8_t* special) {
  int ok = 1;
  for (int s = 0; s < cur; ++s) {
    int layer = seq ? (s % P) / 3 : s / 3;
    ok &= loc_allowed(layer, pa[s]) ? 1 : 0;
  }
  int n1 = cur - 1 > 1 ? cur - 1 : 1;                       // range(max(cur-1, 1))  (Q7)
  for (int s = 0; s < n1; ++s) ok &= (pa[s] != pa[s + 1]) ? 1 : 0;
  int n2 = cur - 2 > 1 ? cur - 2 : 1;
  for (int s = 0; s < n2; ++s) ok &= (pa[s] != pa[s + 2]) ? 1 : 0;
  return ok;
}

__device__ inline int distinct_specials_gt1(const int32_t* w, int n, const uint8_t* special) {
  int first = -1;
  for (int k = 0; k < n; ++k) {
    int a = w[k];
    if (special[a]) {
      if (first < 0) first = a;
      else if (a != first) return 1;
    }
  }
  return 0;
}

// T_alloc = row stride of prev_actions; the dup checks read pa[s+1], pa[s+2] with s+2 <= max(cur-1,2)
// which stays inside the row for cur <= T and T >= 3.
__global__ void k_violation(int B, int T, int P, int seq, int cur, const int32_t* __restrict__ prev_actions,
                            const uint8_t* __restrict__ special, int32_t* __restrict__ out) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int32_t* pa = prev_actions + (size_t)b * T;
  int ok = violation_row(pa, cur, P, seq, special);
  if (seq) {
    int cp = cur % P;                                        // Q9: pages 0..cur%P only
    for (int j = 0; j <= cp; ++j) {
      int lo = P * j, hi = P * (j + 1);
      if (lo >= T) break;
      if (hi > T) hi = T;
      if (distinct_specials_gt1(pa + lo, hi - lo, special)) ok = 0;
    }
  } else {
    if (distinct_specials_gt1(pa, T, special)) ok = 0;
  }
  out[b] = ok;
}

__global__ void k_reward(int B, int T, int P, int seq, int cur /*after act*/, int zero_on_violation,
                         const int32_t* __restrict__ prev_actions, const uint8_t* __restrict__ special,
                         const double* __restrict__ price, const float* __restrict__ p1 /*[B*rpe]*/,
                         int rpe, double* __restrict__ reward, float* __restrict__ click_p) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int32_t* pa = prev_actions + (size_t)b * T;
  int w0 = seq ? cur - P : 0;
  double s = 0.0;
  for (int j = 0; j < rpe; ++j) {
    float pj = p1[(size_t)b * rpe + j];
    s += price[pa[w0 + j]] * (double)pj;
    if (click_p) click_p[(size_t)b * rpe + j] = pj;
  }
  if (zero_on_violation) {
    int ok = violation_row(pa, cur, P, seq, special);
    if (seq) {
      int cp = cur % P;
      for (int j = 0; j <= cp; ++j) {
        int lo = P * j, hi = P * (j + 1);
        if (lo >= T) break;
        if (hi > T) hi = T;
        if (distinct_specials_gt1(pa + lo, hi - lo, special)) ok = 0;
      }
    } else if (distinct_specials_gt1(pa, T, special)) ok = 0;
    if (!ok) s = 0.0;
  }
  reward[b] = s;
}

__global__ void k_fill_f64(int n, double v, double* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = v;
}
__global__ void k_fill_u8(int n, uint8_t v, uint8_t* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = v;
}

// d3rl 'masked_actions' (slate.py:98-104 / seqslate.py:18-23)
__global__ void k_masked_actions(int B, int T, int w0, int W, const int32_t* __restrict__ prev_actions,
                                 int32_t* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * W) return;
  int b = i / W, k = i % W;
  out[i] = prev_actions[(size_t)b * T + w0 + k];
}

// ------------------------------------------------------------------------------------------
// logged policy: SlateState.offline_action / offline_reward (slate.py:149-174, seqslate.py:71-86)
// ------------------------------------------------------------------------------------------
__global__ void k_offline_action(int B, int S, int cur, int T, int emb_dim, const int32_t* __restrict__ row_idx,
                                 const int32_t* __restrict__ log_items, const double* __restrict__ action_emb,
                                 int32_t* __restrict__ items, double* __restrict__ emb) {
  int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (b >= B) return;
  int a = (cur < T && cur < S) ? log_items[(size_t)row_idx[b] * S + cur] : 0;
  if (items && lane == 0) items[b] = a;
  if (emb)
    for (int k = lane; k < emb_dim; k += 32) emb[(size_t)b * emb_dim + k] = action_emb[(size_t)a * emb_dim + k];
}

__global__ void k_offline_reward(int B, int S, int lo, int hi, const int32_t* __restrict__ row_idx,
                                 const int32_t* __restrict__ log_items, const uint8_t* __restrict__ fb,
                                 const double* __restrict__ price, double* __restrict__ out) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  size_t base = (size_t)row_idx[b] * S;
  double s = 0.0;
  for (int k = lo; k < hi; ++k) s += price[log_items[base + k]] * (double)fb[base + k];
  out[b] = s;
}

}  // namespace r4
