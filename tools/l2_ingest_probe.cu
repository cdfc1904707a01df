// l2_ingest_probe.cu -- how many bytes per clock can ONE SM pull out of L2?  (the AUGRU kernels all sit at ~34 B/clk/SM)
// Each CTA (1 per SM: 200 KB of dynamic shared memory) streams its own L2-resident region into a shared-memory ring with
// cp.async.bulk (TMA, mode 0), or with 128-bit LDGs into registers (mode 1), or both at once (mode 2).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/build/l2_ingest_probe tools/l2_ingest_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}\n" :: "r"(smem_u32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int STAGE = 16384, NST = 8;

__global__ void __launch_bounds__(256, 1) k_ingest(const uint8_t* base, size_t region, int shared_region, int iters, int mode, float* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[NST];
  const uint8_t* src = base + (shared_region ? 0 : (size_t)blockIdx.x * region);
  const int per = (int)(region / STAGE);
  if (threadIdx.x == 0) { for (int i = 0; i < NST; ++i) mbar_init(&full[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  float acc = 0.f;
  if (threadIdx.x == 0 && mode != 1) {
    // one thread: keep NST stages in flight; a stage is "consumed" as soon as it has landed
    for (int i = 0; i < NST; ++i) { mbar_expect_tx(&full[i], STAGE); bulk_g2s(smem + i * STAGE, src + (size_t)(i % per) * STAGE, STAGE, &full[i]); }
    for (int it = 0; it < iters; ++it) {
      const int s = it % NST; const uint32_t par = (it / NST) & 1;
      mbar_wait(&full[s], par);
      const int nx = it + NST;
      if (nx < iters) { mbar_expect_tx(&full[s], STAGE); bulk_g2s(smem + s * STAGE, src + (size_t)(nx % per) * STAGE, STAGE, &full[s]); }
    }
  } else if (threadIdx.x >= 32 && mode != 0) {
    // 7 warps of 128-bit loads: iters * STAGE bytes in total per CTA (mode 1), or the same on top of the TMA stream (mode 2)
    const int t = threadIdx.x - 32, nt = 224;
    const float4* p = reinterpret_cast<const float4*>(src);
    const size_t n16 = (size_t)iters * STAGE / 16, r16 = region / 16;
    for (size_t i = t; i < n16; i += (size_t)nt * 4) {
      float4 a = __ldg(p + (i % r16)), b = __ldg(p + ((i + nt) % r16)), c = __ldg(p + ((i + 2 * nt) % r16)), d = __ldg(p + ((i + 3 * nt) % r16));
      acc += a.x + b.y + c.z + d.w;
    }
  }
  if (acc == 1234.5f) sink[0] = acc;
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
  const int sms = pr.multiProcessorCount;
  const size_t region = 393216;                       // the AUGRU weight image of one CTA rank
  uint8_t* buf; CK(cudaMalloc(&buf, region * sms)); CK(cudaMemset(buf, 1, region * sms));
  float* sink; CK(cudaMalloc(&sink, 4));
  const int smem = NST * STAGE + 65536;                // 192 KB: one CTA per SM
  CK(cudaFuncSetAttribute(k_ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 24 * 256;                          // 256 "steps" of 24 stages = 100 MB per CTA
  for (int mode = 0; mode < 3; ++mode)
    for (int sh = 0; sh < 2; ++sh)
      for (int ctas : {sms, 128, 64, 16, 1}) {
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
          cudaEventRecord(e0);
          k_ingest<<<ctas, 256, smem>>>(buf, region, sh, iters, mode, sink);
          cudaEventRecord(e1); CK(cudaDeviceSynchronize());
          float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double bytes = (double)iters * STAGE * (mode == 2 ? 2 : 1);
        printf("mode %d (%s) %s regions, %3d CTAs: %.3f ms -> %.1f B/clk/SM @%d MHz (nominal), %.2f TB/s total\n", mode,
               mode == 0 ? "TMA bulk" : (mode == 1 ? "LDG.128" : "TMA + LDG"), sh ? "ONE shared" : "per-CTA", ctas, best,
               bytes / (best * 1e-3 * clk * 1e3), clk / 1000, bytes * ctas / (best * 1e-3) / 1e12);
      }
  return 0;
}
