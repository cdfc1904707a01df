// tc_probe.cu -- development probe for the tcgen05 building blocks used by the AUGRU kernel:
//   (1) SWIZZLE_NONE K-major core-matrix operands written by threads, tcgen05.mma kind::f16 (bf16),
//       accumulator in TMEM, tcgen05.ld 32x32b epilogue;       D[128,256] = A[128,K] . B[256,K]^T
//   (2) B operand brought in by cp.async.bulk (1-D TMA) from a pre-tiled global image;
//   (3) issue-rate timing of back-to-back MMAs.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/build/tc_probe tools/tc_probe.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int M = 128, N = 256, K = 64;
constexpr int LBO = 128;                 // bytes between K-adjacent core matrices
constexpr int SBO = (K / 8) * 128;       // bytes between M/N-adjacent core matrices (8-row groups)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;                // version = 1 (sm100)
  // base_offset = 0, lbo_mode = 0, layout_type = 0 (SWIZZLE_NONE)
  return d;
}

// instruction descriptor: c=f32, a=b=bf16, K-major both, N, M
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
  uint32_t d = 0;
  d |= 1u << 4;              // c_format = F32
  d |= 1u << 7;              // a_format = BF16
  d |= 1u << 10;             // b_format = BF16
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n"
      :: "r"(smem_u32(bar)), "r"(parity) : "memory");
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

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}

// mode 0: B loaded by threads; mode 1: B by cp.async.bulk from the pre-tiled image; reps: MMA repetitions for timing
__global__ void __launch_bounds__(192, 1) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, const uint8_t* Btiled,
                                                float* D, int mode, int reps, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                       // M*K*2 = 16 KB
  uint8_t* sB = smem + M * K * 2;           // N*K*2 = 32 KB
  __shared__ uint64_t bar_mma, bar_tma;
  __shared__ uint32_t tmem_base;
  int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&bar_mma, 1); mbar_init(&bar_tma, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // operands -> core-matrix layout
  for (int i = tid; i < M * K / 8; i += blockDim.x) {
    int m = i / (K / 8), kc = i % (K / 8);
    uint4 v = *reinterpret_cast<const uint4*>(A + (size_t)m * K + kc * 8);
    *reinterpret_cast<uint4*>(sA + (m / 8) * SBO + kc * LBO + (m % 8) * 16) = v;
  }
  if (mode == 0) {
    for (int i = tid; i < N * K / 8; i += blockDim.x) {
      int n = i / (K / 8), kc = i % (K / 8);
      uint4 v = *reinterpret_cast<const uint4*>(B + (size_t)n * K + kc * 8);
      *reinterpret_cast<uint4*>(sB + (n / 8) * SBO + kc * LBO + (n % 8) * 16) = v;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tbase = tmem_base;
  if (mode == 1 && tid == 0) {
    mbar_expect_tx(&bar_tma, N * K * 2);
    bulk_g2s(sB, Btiled, N * K * 2, &bar_tma);
  }
  if (warp == 4 && lane == 0) {
    if (mode == 1) mbar_wait(&bar_tma, 0);
    uint32_t idesc = make_idesc(M, N);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int k16 = 0; k16 < K / 16; ++k16) {
        uint64_t da = make_desc(smem_u32(sA) + k16 * 2 * LBO, LBO, SBO);
        uint64_t db = make_desc(smem_u32(sB) + k16 * 2 * LBO, LBO, SBO);
        mma_bf16(tbase, da, db, idesc, (r > 0 || k16 > 0) ? 1u : 0u);
      }
    }
    umma_commit(&bar_mma);
    mbar_wait(&bar_mma, 0);
    long long t1 = clock64();
    if (cycles) *cycles = t1 - t0;
  }
  if (warp < 4) {
    mbar_wait(&bar_mma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(256));
}

int main() {
  std::vector<__nv_bfloat16> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K);
  srand(1);
  for (int i = 0; i < M * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < N * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
  std::vector<uint8_t> hBt(N * K * 2);
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k)
    *reinterpret_cast<__nv_bfloat16*>(&hBt[(n / 8) * SBO + (k / 8) * LBO + (n % 8) * 16 + (k % 8) * 2]) = hB[n * K + k];
  std::vector<double> ref(M * N);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)fA[m * K + k] * fB[n * K + k]; ref[m * N + n] = s; }
  __nv_bfloat16 *dA, *dB; uint8_t* dBt; float* dD; long long* dC;
  CK(cudaMalloc(&dA, M * K * 2)); CK(cudaMalloc(&dB, N * K * 2)); CK(cudaMalloc(&dBt, N * K * 2));
  CK(cudaMalloc(&dD, M * N * 4)); CK(cudaMalloc(&dC, 8));
  CK(cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBt, hBt.data(), N * K * 2, cudaMemcpyHostToDevice));
  int smem_bytes = M * K * 2 + N * K * 2 + 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  std::vector<float> hD(M * N);
  for (int mode = 0; mode < 2; ++mode) {
    CK(cudaMemset(dD, 0, M * N * 4));
    probe<<<1, 192, smem_bytes>>>(dA, dB, dBt, dD, mode, 1, dC);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
    double worst = 0; int bad = 0;
    for (int i = 0; i < M * N; ++i) { double e = fabs(hD[i] - ref[i]); if (e > worst) worst = e; if (e > 1e-3) ++bad; }
    printf("mode %d: max abs err %.3g, bad %d / %d  -> %s   (D[0]=%f ref %f, D[last]=%f ref %f)\n", mode, worst, bad, M * N,
           bad == 0 ? "PASS" : "FAIL", hD[0], ref[0], hD[M * N - 1], ref[M * N - 1]);
  }
  for (int reps : {1, 16, 64, 256}) {
    probe<<<1, 192, smem_bytes>>>(dA, dB, dBt, dD, 0, reps, dC);
    CK(cudaDeviceSynchronize());
    long long c; CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
    printf("reps %d: %lld cycles for %d MMAs (128x256x16) -> %.1f cycles/MMA\n", reps, c, reps * K / 16, (double)c / (reps * K / 16));
  }
  return 0;
}
