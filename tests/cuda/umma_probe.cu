// umma_probe.cu — stand-alone check of the operand layout / descriptor conventions of csrc/brief_umma.cuh on a
// real B200: the three contraction forms the SIREN kernels use (forward K-major x K-major, dX K-major x MN-major,
// dW MN-major x MN-major with M padded to 128), with small-integer operands so that bf16 is exact.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tests/cuda/umma_probe.cu && ./umma_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "../../brief_pytorch_b200/csrc/brief_umma.cuh"

using namespace brief::umma;

template <int F>
__global__ void __launch_bounds__(128) probe(const float* A_in, const float* G_in, const float* W_in, float* D1,
                                             float* D2, float* D3, float* D4) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sG = smem;                    // [128 x F] interleaved (G first: dW reads 2x its size as garbage rows)
  unsigned char* sA = sG + 128 * F * 2;        // [128 x F]
  unsigned char* sW = sA + 128 * F * 2;        // [F x F]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int t = threadIdx.x, warp = t >> 5;
  for (int c = 0; c < F; ++c) {
    *reinterpret_cast<__nv_bfloat16*>(sA + chunk_off(t, c >> 3, 128) + (c & 7) * 2) = __float2bfloat16(A_in[t * F + c]);
    *reinterpret_cast<__nv_bfloat16*>(sG + chunk_off(t, c >> 3, 128) + (c & 7) * 2) = __float2bfloat16(G_in[t * F + c]);
  }
  if (t < F)
    for (int c = 0; c < F; ++c)
      *reinterpret_cast<__nv_bfloat16*>(sW + chunk_off(t, c >> 3, F) + (c & 7) * 2) = __float2bfloat16(W_in[t * F + c]);
  if (t == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base;
  const uint32_t aA = smem_u32(sA), aG = smem_u32(sG), aW = smem_u32(sW);
  if (t == 0) {
    // 1) forward: D1[s][o] = sum_k A[s][k] W[o][k]
    for (int k = 0; k < F / 16; ++k)
      mma_f16(tm + 0, make_desc(aA + k * 2 * 2048, 2048, 128), make_desc(aW + k * 2 * (F / 8) * 128, (F / 8) * 128, 128),
               make_idesc(128, F, false, false, true), k > 0);
    // 2) dX: D2[s][k] = sum_o A[s][o] W[o][k]   (B = W seen MN-major: N = k, K = o)
    for (int k = 0; k < F / 16; ++k)
      mma_f16(tm + F, make_desc(aA + k * 2 * 2048, 2048, 128), make_desc(aW + k * 2 * 128, 128, (F / 8) * 128),
               make_idesc(128, F, false, true, true), k > 0);
    // 3) dW: D3[o][k] = sum_s G[s][o] A[s][k]   (both MN-major, K = 128 samples, M padded to 128)
    for (int k = 0; k < 128 / 16; ++k)
      mma_f16(tm + 2 * F, make_desc(aG + k * 2 * 128, 128, 2048), make_desc(aA + k * 2 * 128, 128, 2048),
               make_idesc(128, F, true, true, true), k > 0);
    // 4) dW with M = 64 (no padded rows): where do the 64 accumulator rows land in TMEM?
    for (int k = 0; k < 128 / 16; ++k)
      mma_f16(tm + 3 * F, make_desc(aG + k * 2 * 128, 128, 2048), make_desc(aA + k * 2 * 128, 128, 2048),
               make_idesc(64, F, true, true, true), k > 0);
    // 5) a second M = 64 accumulator in the SAME columns as 4), 16 lanes higher (the unused half of every quadrant):
    //    D5[o][k] = sum_s A[s][o] G[s][k]  (operands swapped so that the result differs from 4)
    for (int k = 0; k < 128 / 16; ++k)
      mma_f16(tm + 3 * F + (16u << 16), make_desc(aA + k * 2 * 128, 128, 2048), make_desc(aG + k * 2 * 128, 128, 2048),
               make_idesc(64, F, true, true, true), k > 0);
    commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const uint32_t lane_addr = tm + ((uint32_t)(warp * 32) << 16);
  float v[16];
  for (int part = 0; part < 4; ++part) {
    float* D = part == 0 ? D1 : part == 1 ? D2 : part == 2 ? D3 : D4;
    for (int c0 = 0; c0 < F; c0 += 16) {
      tmem_ld16(lane_addr + part * F + c0, v);
      tmem_ld_wait();
      for (int i = 0; i < 16; ++i) D[t * F + c0 + i] = v[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int F>
int run() {
  std::vector<float> A(128 * F), G(128 * F), W(F * F), d1(128 * F), d2(128 * F), d3(128 * F), d4(128 * F);
  srand(7 + F);
  for (auto& x : A) x = (float)(rand() % 7 - 3);
  for (auto& x : G) x = (float)(rand() % 5 - 2);
  for (auto& x : W) x = (float)(rand() % 9 - 4);
  float *dA, *dG, *dW, *o1, *o2, *o3, *o4;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dG, G.size() * 4); cudaMalloc(&dW, W.size() * 4);
  cudaMalloc(&o1, d1.size() * 4); cudaMalloc(&o2, d1.size() * 4); cudaMalloc(&o3, d1.size() * 4); cudaMalloc(&o4, d1.size() * 4); cudaMemset(o4, 0, d1.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dG, G.data(), G.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  const int smem = 2 * 128 * F * 2 + F * F * 2 + 32768;  // + slack read by the padded dW rows
  cudaFuncSetAttribute(probe<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<F><<<1, 128, smem>>>(dA, dG, dW, o1, o2, o3, o4);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("F=%d CUDA error: %s\n", F, cudaGetErrorString(e)); return 1; }
  cudaMemcpy(d1.data(), o1, d1.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(d2.data(), o2, d2.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(d3.data(), o3, d3.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(d4.data(), o4, d4.size() * 4, cudaMemcpyDeviceToHost);
  int bad1 = 0, bad2 = 0, bad3 = 0;
  for (int s = 0; s < 128; ++s)
    for (int o = 0; o < F; ++o) {
      float r1 = 0, r2 = 0;
      for (int k = 0; k < F; ++k) { r1 += A[s * F + k] * W[o * F + k]; r2 += A[s * F + k] * W[k * F + o]; }
      bad1 += d1[s * F + o] != r1;
      bad2 += d2[s * F + o] != r2;
    }
  for (int o = 0; o < F; ++o)
    for (int k = 0; k < F; ++k) {
      float r3 = 0;
      for (int s = 0; s < 128; ++s) r3 += G[s * F + o] * A[s * F + k];
      bad3 += d3[o * F + k] != r3;
    }
  // M=64: for each accumulator row find the TMEM lane that holds it
  std::vector<float> ref3(F * F);
  for (int o = 0; o < F; ++o)
    for (int k = 0; k < F; ++k) { float r = 0; for (int s2 = 0; s2 < 128; ++s2) r += G[s2 * F + o] * A[s2 * F + k]; ref3[o * F + k] = r; }
  printf("F=%d M=64 row->lane:", F);
  for (int o = 0; o < F && o < 64; ++o) {
    int found = -1;
    for (int lane = 0; lane < 128; ++lane) {
      bool ok = true;
      for (int k = 0; k < F; ++k) ok = ok && d4[lane * F + k] == ref3[o * F + k];
      if (ok) { found = lane; break; }
    }
    printf(" %d", found);
  }
  printf("\n");
  {  // 5) rows of the swapped product must sit at lane 32*(o/16) + 16 + o%16, and 4) must be intact
    int bad5 = 0, bad4 = 0;
    for (int o = 0; o < F && o < 64; ++o)
      for (int k = 0; k < F; ++k) {
        float r5 = 0; for (int s2 = 0; s2 < 128; ++s2) r5 += A[s2 * F + o] * G[s2 * F + k];
        bad5 += d4[(32 * (o / 16) + 16 + o % 16) * F + k] != r5;
        bad4 += d4[(32 * (o / 16) + o % 16) * F + k] != ref3[o * F + k];
      }
    printf("F=%d lane-offset-16 accumulator: mismatches %d (lower half intact: %d mismatches)\n", F, bad5, bad4);
  }
  printf("F=%d forward mismatches %d, dX mismatches %d, dW mismatches %d  (d1[0..3]= %g %g %g %g)\n", F, bad1, bad2, bad3,
         d1[0], d1[1], d1[2], d1[3]);
  return bad1 + bad2 + bad3;
}

int main(int argc, char** argv) {
  int f = argc > 1 ? atoi(argv[1]) : 64;
  int bad = f == 64 ? run<64>() : f == 32 ? run<32>() : f == 48 ? run<48>() : run<16>();
  printf(bad ? "UMMA PROBE FAILED\n" : "UMMA PROBE OK\n");
  return bad ? 1 : 0;
}
