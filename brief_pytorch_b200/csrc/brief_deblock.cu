// brief_deblock.cu — the reference's deblocking post-filter (deblock.cpp, its only native component: a single-threaded
// triple loop over block seams, deblock.cpp:279-319) as one CUDA launch.
//
// What is kept bit for bit: the H.264-style 6-tap read / 4-tap write filter in the reference's integer arithmetic
// (C truncating divisions, float clip, uint16 wrap-around of the float -> uint16 conversion, deblock.cpp:33-71), the seam
// list with its sticky duplicate flags (:244-276, built on the host by brief_deblock) and — because seams are filtered IN
// PLACE and cross each other — the reference's traversal ORDER: block by block, z by z, (left, right, down, up).
// Seams of different z never touch the same voxel, so one CTA owns one z-slice and walks that slice's seams in the
// reference's order, with a CTA barrier between seams; the pixels along one seam are independent (each reads and writes
// only its own row / column) and are spread over the CTA's threads.  Integer / byte work: no tensor cores, the kernel is
// latency-bound on the barrier chain (blocks x 4 seams per slice), not on HBM — the seams are a tiny fraction of the volume.
#include "brief_kernels.h"

namespace brief {

__device__ __forceinline__ int cdiv_trunc(int a, int b) { return a / b; }  // C semantics: truncation toward zero

// one seam pixel: taps p2 p1 p0 | q0 q1 q2 at ptr[-3*s] .. ptr[2*s]
__device__ __forceinline__ void deblock_pixel(unsigned short* ptr, long long s, float alpha, float beta, int thres) {
  const int p2 = ptr[-3 * s], p1 = ptr[-2 * s], p0 = ptr[-s], q0 = ptr[0], q1 = ptr[s], q2 = ptr[2 * s];
  if ((p1 + p0 + q0 + q1) / 4 > thres) return;  // bright areas are left alone
  if (!((float)abs(p0 - q0) < alpha && (float)abs(p1 - p0) < beta && (float)abs(q1 - q0) < beta)) return;
  float d0 = (float)cdiv_trunc(4 * (q0 - p0) + (p1 - q1) + 4, 8);
  float dp1 = (float)cdiv_trunc(p2 + (p0 + q0 + 1) / 2 - 2 * p1, 2);
  float dq1 = (float)cdiv_trunc(q2 + (q0 + p0 + 1) / 2 - 2 * q1, 2);
  const float c1 = 20.f;
  float c0 = 20.f;
  if ((float)abs(p2 - p0) < beta) c0 += 1.f;
  if ((float)abs(q2 - q0) < beta) c0 += 1.f;
  d0 = fminf(fmaxf(d0, -c0), c0);
  dp1 = fminf(fmaxf(dp1, -c1), c1);
  dq1 = fminf(fmaxf(dq1, -c1), c1);
  // uint16 += float as the reference's compiler does it: float -> int32 (truncation), low 16 bits
  ptr[-2 * s] = (unsigned short)(int)((float)p1 + dp1);
  ptr[-s] = (unsigned short)(int)((float)p0 + d0);
  ptr[0] = (unsigned short)(int)((float)q0 - d0);
  ptr[s] = (unsigned short)(int)((float)q1 + dq1);
}

__global__ void __launch_bounds__(256) deblock_kernel(unsigned short* img, int D, int H, int W, const DeblockBlock* blocks,
                                                      int n_blocks, float alpha, float beta, int thres) {
  const int z = blockIdx.x;
  unsigned short* slice = img + (long long)z * H * W;
  for (int b = 0; b < n_blocks; ++b) {
    const DeblockBlock k = blocks[b];
    if (z < k.z1 || z > k.z2) continue;  // CTA-uniform
    for (int s = 0; s < 4; ++s) {
      if (!((k.mask >> s) & 1)) continue;
      const int l = s == 1 ? k.x2 : k.x1, r = s == 0 ? k.x1 : k.x2;
      const int d = s == 3 ? k.y2 : k.y1, u = s == 2 ? k.y1 : k.y2;
      if (l == r && (l - 3 < 0 || l + 3 > W - 1)) continue;
      else if (d == u && (d - 3 < 0 || d + 3 > H - 1)) continue;
      if (l == r) {  // vertical seam at column l: rows d..u, taps along x
        for (int y = d + threadIdx.x; y <= u; y += blockDim.x) deblock_pixel(slice + (long long)y * W + l, 1, alpha, beta, thres);
      } else if (d == u) {  // horizontal seam at row d: columns l..r, taps along y
        for (int x = l + threadIdx.x; x <= r; x += blockDim.x) deblock_pixel(slice + (long long)d * W + x, W, alpha, beta, thres);
      }
      __syncthreads();  // the next seam may read what this one wrote
    }
  }
}

cudaError_t launch_deblock(unsigned short* img, int D, int H, int W, const DeblockBlock* dev_blocks, int n_blocks,
                           float alpha, float beta, int thres, cudaStream_t st) {
  if (D < 1 || n_blocks < 1) return cudaSuccess;
  deblock_kernel<<<D, 256, 0, st>>>(img, D, H, W, dev_blocks, n_blocks, alpha, beta, thres);
  return cudaGetLastError();
}

}  // namespace brief
