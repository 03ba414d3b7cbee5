// brief_deblock.cu — the reference's deblocking post-filter (deblock.cpp, its only native component: a single-threaded
// triple loop over block seams, deblock.cpp:279-319) as one CUDA launch.
//
// What is kept bit for bit: the H.264-style 6-tap read / 4-tap write filter in the reference's integer arithmetic
// (C truncating divisions, float clip, uint16 wrap-around of the float -> uint16 conversion, deblock.cpp:33-71), the seam
// list with its sticky duplicate flags (:244-276, built on the host by brief_deblock) and — because seams are filtered IN
// PLACE and cross each other — the reference's traversal ORDER: block by block, z by z, (left, right, down, up).
// Seams of different z never touch the same voxel, so one CTA owns one z-slice; seams whose footprints do not touch are
// independent, so the host turns the reference's order into waves of mutually independent seams (see deblock_kernel).
// Integer / byte work: no tensor cores; algorithmic bytes 12 B read + 8 B written per seam pixel, a tiny fraction of the
// volume, so the launch is bound by the few barriers per slice and the strided column accesses, not by HBM.
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

// One CTA per z slice.  The seams of the slice are filtered in WAVES: the host orders the reference's seam list and gives
// every seam the wave 1 + max(wave of the earlier seams whose footprint it touches), so the seams of one wave read and
// write disjoint voxels — any order inside a wave, and the reference's order between conflicting seams, give the
// reference's bits.  One CTA barrier per wave (a regular 8 x 8 block grid needs a handful) instead of one per seam; inside a
// wave every warp takes a seam and spreads its pixels over the lanes.
__global__ void __launch_bounds__(256) deblock_kernel(unsigned short* img, int D, int H, int W, const DeblockSeam* seams,
                                                      const int* wave_off, int n_waves, float alpha, float beta, int thres) {
  const int z = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  unsigned short* slice = img + (long long)z * H * W;
  for (int w = 0; w < n_waves; ++w) {
    const int lo = wave_off[w], hi = wave_off[w + 1];
    for (int i = lo + warp; i < hi; i += n_warps) {
      const DeblockSeam k = seams[i];
      if (z < k.z1 || z > k.z2) continue;
      if (k.l == k.r) {  // vertical seam at column l: rows d..u, taps along x
        for (int y = k.d + lane; y <= k.u; y += 32) deblock_pixel(slice + (long long)y * W + k.l, 1, alpha, beta, thres);
      } else {           // horizontal seam at row d: columns l..r, taps along y
        for (int x = k.l + lane; x <= k.r; x += 32) deblock_pixel(slice + (long long)k.d * W + x, W, alpha, beta, thres);
      }
    }
    __syncthreads();  // the next wave may read what this one wrote
  }
}

cudaError_t launch_deblock(unsigned short* img, int D, int H, int W, const DeblockSeam* dev_seams, const int* dev_wave_off,
                           int n_waves, float alpha, float beta, int thres, cudaStream_t st) {
  if (D < 1 || n_waves < 1) return cudaSuccess;
  deblock_kernel<<<D, 256, 0, st>>>(img, D, H, W, dev_seams, dev_wave_off, n_waves, alpha, beta, thres);
  return cudaGetLastError();
}

}  // namespace brief
