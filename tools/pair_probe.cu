// pair_probe.cu -- development probe for tcgen05.mma.cta_group::2 (one MMA spanning a 2-CTA cluster):
//   (1) M=128 (64 rows of A per CTA), N=256 (128 rows of B per CTA), bf16, SWIZZLE_NONE K-major operands;
//       where do the accumulators land in each CTA's TMEM?  (hypothesis: lanes 0-63 = columns 0-127 of the
//       CTA's 64 rows, lanes 64-127 = columns 128-255: the "2x2" atom of cute's tmem_frg_2sm)
//   (2) tcgen05.alloc / commit (multicast) / dealloc in cta_group::2 form;
//   (3) issue rate of back-to-back pair MMAs;  (4) round-trip latency of a remote mbarrier arrive.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/build/pair_probe tools/pair_probe.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int M = 128, N = 256, K = 64;
constexpr int LBO = 128, SBO = (K / 8) * 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma2_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit2_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
               "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t laddr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(laddr), "r"(rank)); return r;
}
__device__ __forceinline__ void arrive_remote(uint32_t caddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(caddr) : "memory");
}
__device__ __forceinline__ void arrive_remote_relaxed(uint32_t caddr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" :: "r"(caddr) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                 "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
               : "r"(taddr) : "memory");
}

// A: [128][K] fp32 (bf16-exact), B: [256][K]; out: [2 ranks][128 lanes][128 cols]; cyc[0] = cycles of 256 MMAs,
// cyc[1] = cycles of 64 remote-arrive round trips.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(const float* A, const float* B, float* out, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                       // 64 rows x K bf16 = 8 KB
  uint8_t* sB = smem + 64 * K * 2;          // 128 rows x K bf16 = 16 KB
  __shared__ uint64_t bar_done, bar_t, bar_ping, bar_many;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank();
  for (int i = tid; i < 64 * K; i += 128) {
    int r = i / K, k = i % K;
    __nv_bfloat16 v = __float2bfloat16(A[(size_t)(64 * rank + r) * K + k]);
    *reinterpret_cast<__nv_bfloat16*>(sA + (r / 8) * SBO + (k / 8) * LBO + (r % 8) * 16 + (k % 8) * 2) = v;
  }
  for (int i = tid; i < 128 * K; i += 128) {
    int n = i / K, k = i % K;
    __nv_bfloat16 v = __float2bfloat16(B[(size_t)(128 * rank + n) * K + k]);
    *reinterpret_cast<__nv_bfloat16*>(sB + (n / 8) * SBO + (k / 8) * LBO + (n % 8) * 16 + (k % 8) * 2) = v;
  }
  if (tid == 0) {
    mbar_init(&bar_done, 1); mbar_init(&bar_t, 1); mbar_init(&bar_ping, 1); mbar_init(&bar_many, 1 << 20);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  const uint32_t idesc = make_idesc(M, N);
  if (rank == 0 && tid == 0) {
    for (int k = 0; k < K / 16; ++k)
      mma2_bf16(tbase, make_desc(smem_u32(sA) + k * 2 * LBO, LBO, SBO), make_desc(smem_u32(sB) + k * 2 * LBO, LBO, SBO), idesc, k ? 1u : 0u);
    commit2_mc(&bar_done, 3);
  }
  mbar_wait_cluster(&bar_done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int ch = 0; ch < 8; ++ch) {
    float v[16];
    tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + ch * 16, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[((size_t)rank * 128 + warp * 32 + lane) * 128 + ch * 16 + j] = v[j];
  }
  // (3) issue rate: 256 pair MMAs into columns 128.. (scratch), one commit
  if (rank == 0 && tid == 0) {
    long long t0 = clock64();
    for (int it = 0; it < 64; ++it)
      for (int k = 0; k < K / 16; ++k)
        mma2_bf16(tbase + 128, make_desc(smem_u32(sA) + k * 2 * LBO, LBO, SBO), make_desc(smem_u32(sB) + k * 2 * LBO, LBO, SBO), idesc, 1u);
    commit2_mc(&bar_t, 3);
    mbar_wait_cluster(&bar_t, 0);
    cyc[0] = clock64() - t0;
  } else if (tid == 0) {
    mbar_wait_cluster(&bar_t, 0);
  }
  // (4) remote arrive ping-pong: rank 0 arrives on rank 1's barrier, rank 1 answers on rank 0's.
  __syncthreads();
  cluster_sync_all();
  if (tid == 0) {
    const uint32_t remote = mapa(smem_u32(&bar_ping), rank ^ 1);
    long long t0 = clock64();
    for (int it = 0; it < 64; ++it) {
      if (rank == 0) { arrive_remote(remote); mbar_wait_cluster(&bar_ping, it & 1); }
      else { mbar_wait_cluster(&bar_ping, it & 1); arrive_remote(remote); }
    }
    if (rank == 0) cyc[1] = clock64() - t0;
  }
  // (5) issue cost of 64 back-to-back remote arrives (nobody waits): release vs relaxed
  __syncthreads();
  cluster_sync_all();
  if (tid == 0 && rank == 1) {
    const uint32_t remote = mapa(smem_u32(&bar_many), 0);
    long long t0 = clock64();
    for (int it = 0; it < 64; ++it) arrive_remote(remote);
    long long t1 = clock64();
    for (int it = 0; it < 64; ++it) arrive_remote_relaxed(remote);
    long long t2 = clock64();
    cyc[2] = t1 - t0; cyc[3] = t2 - t1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(512));
}

int main() {
  std::vector<float> A(M * K), B(N * K), D((size_t)M * N);
  srand(5);
  for (auto& x : A) x = (float)((rand() % 31) - 15) / 8.0f;
  for (auto& x : B) x = (float)((rand() % 29) - 14) / 16.0f;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k];
    D[(size_t)m * N + n] = (float)s;
  }
  float *dA, *dB, *dO; long long* dC;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dO, 2 * 128 * 128 * 4)); CK(cudaMalloc(&dC, 32));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0xff, 2 * 128 * 128 * 4)); CK(cudaMemset(dC, 0, 32));
  int smem = 64 * K * 2 + 128 * K * 2;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe<<<2, 128, smem>>>(dA, dB, dO, dC);
  CK(cudaDeviceSynchronize());
  std::vector<float> O(2 * 128 * 128); long long cyc[4];
  CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(cyc, dC, 32, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int r = 0; r < 2; ++r) for (int l = 0; l < 128; ++l) for (int c = 0; c < 128; ++c) {
    float exp = D[(size_t)(64 * r + (l % 64)) * N + (l / 64) * 128 + c];
    if (O[((size_t)r * 128 + l) * 128 + c] != exp) ++bad;
  }
  printf("hypothesis (lane<64: cols 0-127, lane>=64: cols 128-255 of the CTA's 64 rows): %d / %d mismatches\n", bad, 2 * 128 * 128);
  if (bad) {
    // brute-force where a few samples came from
    for (int r = 0; r < 2; ++r) for (int l : {0, 1, 31, 32, 63, 64, 96, 127}) for (int c : {0, 1, 64, 127}) {
      float v = O[((size_t)r * 128 + l) * 128 + c];
      int fm = -1, fn = -1, cnt = 0;
      for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) if (D[(size_t)m * N + n] == v) { if (!cnt) { fm = m; fn = n; } ++cnt; }
      printf("  rank %d lane %3d col %3d = %9.4f  <- D[%d][%d] (%d candidates)\n", r, l, c, v, fm, fn, cnt);
    }
  }
  printf("256 pair MMAs (128x256x16): %lld cycles = %.1f / MMA;  remote-arrive round trip: %.1f cycles\n", cyc[0], cyc[0] / 256.0, cyc[1] / 64.0);
  printf("64 back-to-back remote arrives: release.cluster %.1f cycles each, relaxed.cluster %.1f cycles each\n", cyc[2] / 64.0, cyc[3] / 64.0);
  return 0;
}
