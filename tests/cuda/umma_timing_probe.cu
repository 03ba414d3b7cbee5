// umma_timing_probe.cu — issue/complete timing of tcgen05.mma chains in the interleaved (SWIZZLE_NONE) layout:
// n back-to-back MMAs (M=128 or 64, N, K=16 each) + commit, cycles from first issue to (a) last issue, (b) mbarrier wake.
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include "../../brief_pytorch_b200/csrc/brief_umma.cuh"
using namespace brief::umma;

__global__ void __launch_bounds__(128) probe(int M, int N, int n_mma, int mn_major, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar; __shared__ uint32_t tmem_base;
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (t == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  fence_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_base, a0 = smem_u32(smem), b0 = a0 + 32768;
  const uint32_t idesc = (1u << 4) | ((mn_major ? 1u : 0u) << 15) | ((mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  uint32_t phase = 0;
  for (int rep = 0; rep < 5; ++rep) {
    long long c0 = 0, c1 = 0, c2 = 0;
    __syncthreads();
    if (warp == 0 && elect_one()) {
      c0 = clock64();
      for (int k = 0; k < n_mma; ++k) {
        const int kk = k & 3;
        if (mn_major) mma_f16(tm, make_desc(a0 + kk * 256, 128, 2048), make_desc(b0 + kk * 256, 128, 2048), idesc, k > 0);
        else mma_f16(tm, make_desc(a0 + kk * 4096, 2048, 128), make_desc(b0 + kk * 2 * (N / 8) * 128, (N / 8) * 128, 128), idesc, k > 0);
      }
      commit(&bar);
      c1 = clock64();
    }
    mbar_wait(&bar, phase); phase ^= 1;
    c2 = clock64();
    if (t == 0 && rep == 4) { out[0] = c1 - c0; out[1] = c2 - c0; }
    tc_fence_after();
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16); long long h[2];
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int cfgs[][4] = {{128, 64, 1, 0}, {128, 64, 4, 0}, {128, 64, 8, 0}, {128, 64, 16, 0}, {128, 64, 64, 0}, {128, 128, 16, 0}, {128, 256, 16, 0},
                         {128, 16, 16, 0}, {128, 32, 16, 0}, {64, 64, 8, 1}, {64, 64, 64, 1}, {64, 16, 8, 1}, {128, 64, 64, 1}};
  for (auto& c : cfgs) {
    probe<<<1, 128, 64 * 1024>>>(c[0], c[1], c[2], c[3], d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("M=%3d N=%3d n_mma=%2d %s : issue %5lld cyc, issue->wake %5lld cyc  (%.1f cyc/MMA)\n", c[0], c[1], c[2], c[3] ? "MN-major" : "K-major ", h[0], h[1], (double)h[1] / c[2]);
  }
  return 0;
}
