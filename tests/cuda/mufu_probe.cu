// mufu_probe.cu — special-function-unit throughput on this GPU: sin.approx / cos.approx / ex2.approx per clock per SM
// with all SMs busy (the roofline denominator of the SIREN epilogues; SURVEY.md section 8d assumed 16/clk/SM).
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void __launch_bounds__(1024) k(float* out, int iters, long long* cyc) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
  const long long c0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) x[i] = __sinf(x[i]);
      else if (OP == 1) x[i] = __cosf(x[i]);
      else if (OP == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      else if (OP == 3) x[i] = fmaf(x[i], 1.0001f, 0.5f);
      else { float s = __sinf(x[i]); x[i] = fmaf(s, 0.5f, x[i]); }
    }
  }
  const long long c1 = clock64();
  float acc = 0;
  for (int i = 0; i < 8; ++i) acc += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = c1 - c0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 2 * 1024 * 4); cudaMalloc(&cyc, 8);
  const char* names[] = {"__sinf (FMUL+MUFU.SIN)", "__cosf (FMUL+MUFU.COS)", "ex2.approx (MUFU.EX2)", "fmaf", "sin + fma"};
  const int iters = 2000;
  for (int threads : {256, 1024}) for (int op = 0; op < 5; ++op) {
    long long h;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (op == 0) k<0><<<148 * (2048 / threads), threads>>>(out, iters, cyc);
      if (op == 1) k<1><<<148 * (2048 / threads), threads>>>(out, iters, cyc);
      if (op == 2) k<2><<<148 * (2048 / threads), threads>>>(out, iters, cyc);
      if (op == 3) k<3><<<148 * (2048 / threads), threads>>>(out, iters, cyc);
      if (op == 4) k<4><<<148 * (2048 / threads), threads>>>(out, iters, cyc);
      cudaEventRecord(e1);
      cudaDeviceSynchronize();
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // whole grid: 148 SMs * 2048 threads * 8 * iters ops in ms; block 0 ran h cycles -> clock = h / ms
    const double ops = 148.0 * 2048.0 * 8 * iters;
    printf("threads/CTA %4d  %-26s : %.3f ms, %.2f Tops/s, block0 %lld cyc (%.2f GHz if it spans the kernel) -> %.1f ops/clk/SM at 1.9 GHz\n",
           threads, names[op], ms, ops / ms / 1e9, h, h / ms / 1e6, ops / ms / 1e-3 / 148.0 / 1.9e9);
  }
  return 0;
}
