// tmem_probe.cu — tcgen05.ld throughput per SM: W warps each issue `reps` x tcgen05.ld.32x32b.{x16,x32,x64} (+ wait::ld
// every `batch` loads); reports bytes per clock.  Also the same with one MMA stream running concurrently.
#include <cstdio>
#include <cstdlib>
#include "../../brief_pytorch_b200/csrc/brief_umma.cuh"
using namespace brief::umma;

template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, float* v) {
  if constexpr (X == 16) tmem_ld16(taddr, v);
  else if constexpr (X == 32) tmem_ld32(taddr, v);
}

template <int X>
__global__ void __launch_bounds__(1024) probe(int reps, int batch, int with_mma, long long* out, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar; __shared__ uint32_t tmem_base;
  const int t = threadIdx.x, warp = t >> 5, nw = blockDim.x >> 5;
  for (int i = t; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (t == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  fence_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_base;
  const bool mma_warp = with_mma && warp == nw - 1;
  const uint32_t my = tm + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * X) % 256;
  float acc = 0.f;
  __syncthreads();
  const long long c0 = clock64();
  if (mma_warp) {
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 32768;
    const uint32_t idesc = make_idesc(128, 64, false, false);
    if (elect_one()) {
      for (int k = 0; k < reps; ++k) mma_f16(tm + 256, make_desc(a0 + (k & 3) * 4096, 2048, 128), make_desc(b0 + (k & 3) * 2048, 1024, 128), idesc, 1);
      commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
  } else {
    float v[X];
    for (int k = 0; k < reps; k += batch) {
      for (int b = 0; b < batch; ++b) ld<X>(my, v);
      tmem_ld_wait();
      acc += v[0] + v[X - 1];
    }
  }
  const long long c1 = clock64();
  __syncthreads();
  const long long c2 = clock64();
  if (t == 0) { out[0] = c1 - c0; out[1] = c2 - c0; }
  if (mma_warp && (t & 31) == 0) out[2] = c1 - c0;
  if (acc == 123.456f) sink[t] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 32); float* sink; cudaMalloc(&sink, 4096); long long h[3];
  cudaFuncSetAttribute(probe<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(probe<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int reps = 2048;
  for (int with_mma = 0; with_mma < 2; ++with_mma)
    for (int x : {16, 32})
      for (int warps : {4, 8, 16, 32})
        for (int batch : {1, 4}) {
          const int threads = 32 * warps + (with_mma ? 32 : 0);
          if (threads > 1024) continue;
          if (x == 16) probe<16><<<1, threads, 64 * 1024>>>(reps, batch, with_mma, d, sink);
          else probe<32><<<1, threads, 64 * 1024>>>(reps, batch, with_mma, d, sink);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
          const double bytes = (double)warps * reps * x * 32 * 4;
          printf("x%-2d warps=%2d batch=%d mma=%d : %8lld cyc total, %6.1f B/clk, %5.1f cyc per warp-load%s\n", x, warps, batch, with_mma,
                 h[1], bytes / h[1], (double)h[1] / reps, with_mma ? "" : "");
          if (with_mma) printf("      concurrent MMA stream: %lld cyc for %d MMAs = %.1f cyc/MMA\n", h[2], reps, (double)h[2] / reps);
        }
  return 0;
}
