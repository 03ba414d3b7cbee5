// umma_dual_probe.cu — does the tcgen05.mma issue rate scale with the number of issuing warps?  W warps (1, 2, 4)
// each issue n back-to-back 128 x N x 16 SS MMAs into their own accumulator columns; reports cycles per MMA overall.
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include "../../brief_pytorch_b200/csrc/brief_umma.cuh"
using namespace brief::umma;

__global__ void __launch_bounds__(128) probe(int N, int n_mma, int issuers, int M, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar[4]; __shared__ uint32_t tmem_base;
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003800u + (i * 2654435761u >> 28);
  if (t == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  fence_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_base, a0 = smem_u32(smem) + warp * 16384, b0 = smem_u32(smem) + 65536 + (warp & 1) * 16384;
  const uint32_t idesc = make_idesc(M, N, M == 64, M == 64);
  long long c0 = 0, c2 = 0;
  __syncthreads();
  c0 = clock64();
  if (warp < issuers) {
    if (elect_one()) {
      for (int k = 0; k < n_mma; ++k) {
        const int kk = k & 3;
        if (M == 64) mma_f16(tm + warp * 128, make_desc(a0 + kk * 256, 128, 2048), make_desc(b0 + kk * 256, 128, 2048), idesc, k > 0);
        else mma_f16(tm + warp * 128, make_desc(a0 + kk * 4096, 2048, 128), make_desc(b0 + kk * 2 * (N / 8) * 128, (N / 8) * 128, 128), idesc, k > 0);
      }
      commit(&bar[warp]);
    }
    __syncwarp();
    mbar_wait(&bar[warp], 0);
  }
  __syncthreads();
  c2 = clock64();
  if (t == 0) out[0] = c2 - c0;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16); long long h;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int M : {128, 64})
    for (int N : {16, 64, 128})
      for (int issuers : {1, 2, 4}) {
        const int n = 64;
        probe<<<1, 128, 100 * 1024>>>(N, n, issuers, M, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("M=%3d N=%3d issuing warps=%d x %d MMAs: %6lld cyc total -> %.1f cyc per MMA overall\n", M, N, issuers, n, h, (double)h / (n * issuers));
      }
  return 0;
}
