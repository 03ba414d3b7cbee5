// umma_mixed_probe.cu — does tcgen05.mma kind::f16 accept DIFFERENT operand formats (A fp16, B bf16 and vice versa)?
// Non-integer operands chosen so that a format mix-up (bits misread) cannot pass by accident.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "../../brief_pytorch_b200/csrc/brief_umma.cuh"
using namespace brief::umma;
constexpr int F = 64;
__host__ __device__ constexpr uint32_t idesc_fmt(int M, int N, int afmt, int bfmt) {
  return (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__global__ void __launch_bounds__(128) probe(const float* A_in, const float* W_in, float* D, int afmt, int bfmt) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sA = smem; unsigned char* sW = sA + 128 * F * 2;
  __shared__ uint64_t bar; __shared__ uint32_t tmem_base;
  const int t = threadIdx.x, warp = t >> 5;
  for (int c = 0; c < F; ++c) {
    void* p = sA + chunk_off(t, c >> 3, 128) + (c & 7) * 2;
    if (afmt == 0) *reinterpret_cast<__half*>(p) = __float2half(A_in[t * F + c]); else *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16(A_in[t * F + c]);
  }
  if (t < F) for (int c = 0; c < F; ++c) {
    void* p = sW + chunk_off(t, c >> 3, F) + (c & 7) * 2;
    if (bfmt == 0) *reinterpret_cast<__half*>(p) = __float2half(W_in[t * F + c]); else *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16(W_in[t * F + c]);
  }
  if (t == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 64);
  fence_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_base;
  if (t == 0) {
    for (int k = 0; k < F / 16; ++k)
      mma_f16(tm, make_desc(smem_u32(sA) + k * 4096, 2048, 128), make_desc(smem_u32(sW) + k * 2 * (F / 8) * 128, (F / 8) * 128, 128),
               idesc_fmt(128, F, afmt, bfmt), k > 0);
    commit(&bar);
  }
  mbar_wait(&bar, 0); tc_fence_after();
  float v[16];
  for (int c0 = 0; c0 < F; c0 += 16) {
    tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, v); tmem_ld_wait();
    for (int i = 0; i < 16; ++i) D[t * F + c0 + i] = v[i];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 64);
}
int main() {
  std::vector<float> A(128 * F), W(F * F), d(128 * F);
  srand(3);
  for (auto& x : A) x = (rand() % 2001 - 1000) / 1024.0f;      // 11 significant bits: exact in fp16, NOT in bf16
  for (auto& x : W) x = (rand() % 255 - 127) / 4096.0f;        // 8 significant bits: exact in both
  float *dA, *dW, *o; cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&o, d.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  for (int afmt = 0; afmt < 2; ++afmt) for (int bfmt = 0; bfmt < 2; ++bfmt) {
    probe<<<1, 128, 128 * F * 2 + F * F * 2>>>(dA, dW, o, afmt, bfmt);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("afmt %d bfmt %d: CUDA error %s\n", afmt, bfmt, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(d.data(), o, d.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int s = 0; s < 128; ++s) for (int n = 0; n < F; ++n) {
      double r = 0; for (int k = 0; k < F; ++k) r += (double)A[s * F + k] * W[n * F + k];
      maxerr = fmax(maxerr, fabs(r - d[s * F + n])); maxref = fmax(maxref, fabs(r));
    }
    printf("A %s x B %s : max abs err %.3e (max |ref| %.3f)\n", afmt ? "bf16" : "fp16", bfmt ? "bf16" : "fp16", maxerr, maxref);
  }
  return 0;
}
