// gemm_probe.cu -- timing + per-role stall counters of k_gemm_tc on the two shapes that matter:
//   (a) the AUGRU/attention input projection: M = ns*64 (sequence mode), K = 128, N = 832, lane-major output;
//   (b) the observation head: M = 4096, K = 3456, N = 256.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/build/gemm_probe tools/gemm_probe.cu
#include "../rl4rs_b200/csrc/r4_gemm_tc.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace r4tc;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static void run(const char* name, int M, int N, int K, int tm_ns, int ldT, int bnt = G_BNMAX) {
  std::vector<float> W((size_t)K * N);
  for (auto& w : W) w = (rand() % 2001 - 1000) / 4000.f;
  std::vector<uint8_t> img(gemm_image_bytes(K, N, bnt));
  build_gemm_image(W.data(), K, N, img.data(), bnt);
  float *dA, *dC, *dT = nullptr, *dK = nullptr, *dB; uint8_t* dimg; long long* ddbg;
  CK(cudaMalloc(&dA, (size_t)M * K * 4)); CK(cudaMemset(dA, 0, (size_t)M * K * 4));
  CK(cudaMalloc(&dC, (size_t)M * N * 4)); CK(cudaMalloc(&dB, N * 4)); CK(cudaMemset(dB, 0, N * 4));
  CK(cudaMalloc(&dimg, img.size())); CK(cudaMemcpy(dimg, img.data(), img.size(), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&ddbg, 16 * 8));
  if (tm_ns) { CK(cudaMalloc(&dT, (size_t)M * ldT * 4)); CK(cudaMalloc(&dK, (size_t)M * (N - ldT) * 4)); }
  GemmTcParams p{dA, K, nullptr, dimg, dB, dC, N, M, N, K, 0, tm_ns, 0, ldT, dT, dK};
  p.dbg = ddbg; p.bnt = bnt;
  CK(cudaFuncSetAttribute(k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_BYTES));
  int tiles = ((M + G_BM - 1) / G_BM) * ((N + bnt - 1) / bnt);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) {
    CK(cudaMemset(ddbg, 0, 16 * 8));
    cudaEventRecord(e0);
    k_gemm_tc<<<std::min(tiles, 148), G_THREADS, G_SMEM_BYTES>>>(p);
    cudaEventRecord(e1); CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long d[16]; CK(cudaMemcpy(d, ddbg, sizeof(d), cudaMemcpyDeviceToHost));
    if (it == 2)
      printf("%s: %.3f ms, %d tiles; CTA0: %lld tiles in %lld cycles | MMA waits: acc_empty %lld, A %lld, W %lld | producer waits: cp.async %lld, "
             "stage empty %lld | epilogue: wait acc_full %lld, drain+store %lld (tmem %lld, fill %lld, proxy fence %lld, wait+bar+issue %lld)\n",
             name, ms, tiles, d[4], d[0], d[1], d[2], d[3], d[5], d[6], d[7], d[8], d[9], d[10], d[11], d[12]);
  }
  cudaFree(dA); cudaFree(dC); cudaFree(dB); cudaFree(dimg); cudaFree(ddbg); if (dT) cudaFree(dT); if (dK) cudaFree(dK);
}

int main() {
  run("input projection (seq mode) M=262144 K=128 N=832", 4096 * 64, 832, 128, 4096, 768);
  run("same shape, plain row-major output           ", 4096 * 64, 832, 128, 0, 0);
  run("head M=4096 K=3456 N=256                      ", 4096, 256, 3456, 0, 0);
  run("head M=36864 K=3456 N=256                     ", 36864, 256, 3456, 0, 0);
  run("head M=4096, n-tiles of 128                   ", 4096, 256, 3456, 0, 0, 128);
  run("head M=36864, n-tiles of 128                  ", 36864, 256, 3456, 0, 0, 128);
  return 0;
}
