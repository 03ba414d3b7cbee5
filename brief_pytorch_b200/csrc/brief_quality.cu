// brief_quality.cu — quality metrics of a decoded volume on the device: sum of squared errors (-> MSE, PSNR) and the
// reference's SSIM (mean over z-slices of the 2-D 11-tap Gaussian-window SSIM, sigma 1.5, valid region, K = (0.01,
// 0.03)) in ONE pass over the two volumes.
//
// Reference work replaced: eval_performance / cal_psnr / cal_ssim (utils/misc.py:447-499) and utils/ssim.py:9-150 —
// float32 copies of both volumes plus five conv2d passes per slice, run at every checkpoint (40x per fit).
// Here each CTA owns a 16 x 16 patch of output pixels of one slice: the (16+10)^2 halo of both volumes is staged in
// shared memory as fp32, every thread forms the five windowed moments of its pixel (mu_x, mu_y, E[xx], E[yy], E[xy]) with
// the separable window (rows first, into registers/shared memory), evaluates the SSIM map value in fp32 like the
// reference and the CTA adds one fp64 partial.  Algorithmic bytes: both volumes read once (2 x sizeof(dtype) per voxel);
// the 2.6x halo re-reads come out of L2.
#include "brief_kernels.h"

namespace brief {

constexpr int kQT = 16;          // output patch edge
constexpr int kQW = 11;          // window
constexpr int kQH = kQT + kQW - 1;  // staged edge (26)

struct QualityArgs {
  const void* a;
  const void* b;
  int dtype, D, H, W;
  float win[kQW];
  float c1, c2;
  double* out;  // [0] sum of squared differences, [1] sum of the SSIM map over all valid pixels
};

template <typename T>
__device__ __forceinline__ float q_load(const void* p, long long i) { return (float)__ldg(reinterpret_cast<const T*>(p) + i); }

template <typename T>
__global__ void __launch_bounds__(kQT * kQT) quality_kernel(QualityArgs q) {
  __shared__ float sx[kQH][kQH + 1], sy[kQH][kQH + 1];
  __shared__ float rx[kQH][kQT + 1], ry[kQH][kQT + 1], rxx[kQH][kQT + 1], ryy[kQH][kQT + 1], rxy[kQH][kQT + 1];
  __shared__ double s_red[2][kQT * kQT / 32];
  const int z = blockIdx.z, y0 = blockIdx.y * kQT, x0 = blockIdx.x * kQT;
  const int t = threadIdx.x;
  const long long base = (long long)z * q.H * q.W;
  // ---- stage the halo; the squared error is summed over the patch's OWN 16 x 16 input pixels (each voxel once)
  double se = 0.0;
  for (int i = t; i < kQH * kQH; i += kQT * kQT) {
    const int r = i / kQH, c = i - r * kQH;
    const int y = y0 + r, x = x0 + c;
    float va = 0.f, vb = 0.f;
    if (y < q.H && x < q.W) {
      va = q_load<T>(q.a, base + (long long)y * q.W + x);
      vb = q_load<T>(q.b, base + (long long)y * q.W + x);
      if (r < kQT && c < kQT) { const double d = (double)va - (double)vb; se += d * d; }
    }
    sx[r][c] = va;
    sy[r][c] = vb;
  }
  __syncthreads();
  // ---- rows: 26 rows x 16 output columns, five moments
  for (int i = t; i < kQH * kQT; i += kQT * kQT) {
    const int r = i / kQT, c = i - r * kQT;
    float mx = 0.f, my = 0.f, xx = 0.f, yy = 0.f, xy = 0.f;
#pragma unroll
    for (int k = 0; k < kQW; ++k) {
      const float w = q.win[k], a = sx[r][c + k], b = sy[r][c + k];
      mx = fmaf(w, a, mx); my = fmaf(w, b, my);
      xx = fmaf(w, a * a, xx); yy = fmaf(w, b * b, yy); xy = fmaf(w, a * b, xy);
    }
    rx[r][c] = mx; ry[r][c] = my; rxx[r][c] = xx; ryy[r][c] = yy; rxy[r][c] = xy;
  }
  __syncthreads();
  // ---- columns + SSIM map for this thread's output pixel (valid region only)
  const int oy = t / kQT, ox = t - oy * kQT;
  double ss = 0.0;
  if (y0 + oy + kQW - 1 < q.H && x0 + ox + kQW - 1 < q.W) {
    float mx = 0.f, my = 0.f, xx = 0.f, yy = 0.f, xy = 0.f;
#pragma unroll
    for (int k = 0; k < kQW; ++k) {
      const float w = q.win[k];
      mx = fmaf(w, rx[oy + k][ox], mx); my = fmaf(w, ry[oy + k][ox], my);
      xx = fmaf(w, rxx[oy + k][ox], xx); yy = fmaf(w, ryy[oy + k][ox], yy); xy = fmaf(w, rxy[oy + k][ox], xy);
    }
    const float mxx = mx * mx, myy = my * my, mxy = mx * my;
    const float s1 = xx - mxx, s2 = yy - myy, s12 = xy - mxy;
    const float cs = (2.f * s12 + q.c2) / (s1 + s2 + q.c2);
    ss = (double)(((2.f * mxy + q.c1) / (mxx + myy + q.c1)) * cs);
  }
  // ---- CTA reduction -> one fp64 atomic per quantity
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    se += __shfl_down_sync(0xffffffffu, se, off);
    ss += __shfl_down_sync(0xffffffffu, ss, off);
  }
  if ((t & 31) == 0) { s_red[0][t >> 5] = se; s_red[1][t >> 5] = ss; }
  __syncthreads();
  if (t == 0) {
    for (int k = 1; k < kQT * kQT / 32; ++k) { se += s_red[0][k]; ss += s_red[1][k]; }
    atomicAdd(q.out, se);
    atomicAdd(q.out + 1, ss);
  }
}

cudaError_t launch_quality(const void* a, const void* b, int dtype, int D, int H, int W, const float* win11, float c1, float c2,
                           double* dev_out, cudaStream_t st) {
  QualityArgs q{};
  q.a = a; q.b = b; q.dtype = dtype; q.D = D; q.H = H; q.W = W; q.c1 = c1; q.c2 = c2; q.out = dev_out;
  for (int k = 0; k < kQW; ++k) q.win[k] = win11[k];
  // every input pixel must belong to exactly one patch's own 16 x 16 square (squared error), so the grid covers H x W
  dim3 grid((W + kQT - 1) / kQT, (H + kQT - 1) / kQT, D);
  if (dtype == 0) quality_kernel<unsigned char><<<grid, kQT * kQT, 0, st>>>(q);
  else if (dtype == 1) quality_kernel<unsigned short><<<grid, kQT * kQT, 0, st>>>(q);
  else quality_kernel<float><<<grid, kQT * kQT, 0, st>>>(q);
  return cudaGetLastError();
}

}  // namespace brief
