// umma_ts_probe.cu — tcgen05.mma with the A operand in TENSOR MEMORY (fp16 pairs, lane = row) against SS mode:
// (a) timing of n back-to-back 128 x N x 16 MMAs, (b) a numerical check of the A-in-TMEM layout
// (A[r][k] = f(r, k) written with tcgen05.st.32x32b, B = identity-like weights).
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include "../../brief_pytorch_b200/csrc/brief_umma.cuh"
using namespace brief::umma;

__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

__global__ void __launch_bounds__(128) probe(int N, int n_mma, int ts, long long* out, float* check) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar; __shared__ uint32_t tmem_base;
  const int t = threadIdx.x, warp = t >> 5;
  // B = [N x 64] K-major interleaved (R = N): B[n][k] = (n == k) ? 1 : 0  -> D[r][n] = A[r][n] for n < 64
  for (int i = t; i < 64 * 1024 / 2; i += 128) reinterpret_cast<__half*>(smem)[i] = __float2half(0.f);
  __syncthreads();
  for (int i = t; i < N * 64; i += 128) {
    const int n = i / 64, k = i % 64;
    const size_t off = ((size_t)(k >> 3) * (N >> 3) + (n >> 3)) * 128 + (n & 7) * 16 + (k & 7) * 2;
    *reinterpret_cast<__half*>(smem + 32768 + off) = __float2half(n == k ? 1.f : 0.f);
  }
  // A in SMEM too (for the SS timing): [128 x 64] interleaved, A[r][k] = r + k/64
  for (int i = t; i < 128 * 64; i += 128) {
    const int r = i / 64, k = i % 64;
    const size_t off = ((size_t)(k >> 3) * 16 + (r >> 3)) * 128 + (r & 7) * 16 + (k & 7) * 2;
    *reinterpret_cast<__half*>(smem + off) = __float2half((float)(r % 32) + k / 64.f);
  }
  if (t == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  fence_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_base, a0 = smem_u32(smem), b0 = a0 + 32768;
  // A in TMEM: columns [256, 256+32): thread t = row t, 64 fp16 = 32 words
  {
    const uint32_t my = tm + ((uint32_t)(32 * warp) << 16) + 256;
    for (int c = 0; c < 4; ++c) {
      uint32_t w[8];
      for (int j = 0; j < 8; ++j) {
        const int k = 16 * c + 2 * j;
        w[j] = pack_f16x2((float)(t % 32) + k / 64.f, (float)(t % 32) + (k + 1) / 64.f);
      }
      tmem_st8(my + 8 * c, w);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t idesc = make_idesc(128, N, false, false);
  uint32_t phase = 0;
  for (int rep = 0; rep < 3; ++rep) {
    long long c0 = 0, c1 = 0, c2 = 0;
    __syncthreads();
    if (warp == 0 && elect_one()) {
      c0 = clock64();
      for (int k = 0; k < n_mma; ++k) {
        const int kk = k & 3;
        const uint64_t bd = make_desc(b0 + kk * 2 * (N / 8) * 128, (N / 8) * 128, 128);
        if (ts) mma_f16_ts(tm, tm + 256 + 8 * kk, bd, idesc, kk > 0);
        else mma_f16(tm, make_desc(a0 + kk * 4096, 2048, 128), bd, idesc, kk > 0);
      }
      commit(&bar);
      c1 = clock64();
    }
    mbar_wait(&bar, phase); phase ^= 1;
    c2 = clock64();
    if (t == 0 && rep == 2) { out[0] = c1 - c0; out[1] = c2 - c0; }
    tc_fence_after();
  }
  // numerical check of the last 4-MMA group: D[r][n] should equal A[r][n] (n < 64)
  if (n_mma % 4 == 0) {
    float v[16];
    tmem_ld16(tm + ((uint32_t)(32 * warp) << 16), v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) check[t * 16 + i] = v[i];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16); long long h[2];
  float* chk; cudaMalloc(&chk, 128 * 16 * 4); float hc[128 * 16];
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {64, 128})
      for (int n : {4, 16, 64}) {
        probe<<<1, 128, 64 * 1024>>>(N, n, ts, d, chk);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s (ts=%d N=%d)\n", cudaGetErrorString(e), ts, N); return 1; }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        cudaMemcpy(hc, chk, sizeof hc, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int r = 0; r < 128; ++r) for (int i = 0; i < 16; ++i) {
          const double want = (r % 32) + i / 64.0;
          maxerr = fmax(maxerr, fabs(hc[r * 16 + i] - want));
        }
        printf("%s N=%3d n_mma=%2d : issue %5lld cyc, issue->wake %5lld cyc (%.1f cyc/MMA)  layout check max|err| = %.4f\n",
               ts ? "A in TMEM" : "A in SMEM", N, n, h[0], h[1], (double)h[1] / n, maxerr);
      }
  return 0;
}
