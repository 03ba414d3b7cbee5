// brief_tc_common.cuh — pieces shared by the tensor-core translation units (brief_tc.cu: decode + narrow fit kernel,
// brief_tc_wide.cu: wide fit kernel): tile constants, MMA issue helpers for the interleaved operand layout, the
// operand-row stores, sine / cosine on the special-function unit, work lookup, optional stage timing.
#pragma once
#include <cuda_fp16.h>
#include <cstdlib>

#include "brief_common.cuh"
#include "brief_image.cuh"
#include "brief_kernels.h"
#include "brief_umma.cuh"

namespace brief {

using namespace umma;

// ---- optional stage timing (tools/tc_stage_timing.py / wide_timing.py build a second library with -DBRIEF_TC_TIMING);
//      every translation unit has its own counter array (no relocatable device code) ------------------------------------
#ifdef BRIEF_TC_TIMING
static __device__ unsigned long long g_tc_timing[64];
#define TT(var) const long long var = clock64()
#define TACC(slot, expr) do { if (blockIdx.x == 0 && lane == 0 && tslot >= 0) \
    atomicAdd(&g_tc_timing[tslot + (slot)], (unsigned long long)(expr)); } while (0)
#else
#define TT(var)
#define TACC(slot, expr)
#endif

constexpr int kTile = 128;
constexpr float kGradScale = 1.0f / 256.0f;  // dy' = kGradScale * w * (yhat - y)
constexpr uint32_t kActLBO = (kTile / 8) * 128;  // 2048: feature-group stride of a [128 x F] operand buffer

__host__ __device__ constexpr int tmem_cols_pow2(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

// ---- packed image: brief_image.cuh (written by pack_kernel, brief_opt.cu) --------------------------------------------
__device__ __forceinline__ int tc_find_work(const int* __restrict__ prefix, int n, int b) {
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(prefix + mid) <= b) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ void tc_load_net(NetDev& dst, const NetDev& src) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&src);
  uint32_t* d = reinterpret_cast<uint32_t*>(&dst);
  for (int i = threadIdx.x; i < (int)(sizeof(NetDev) / 4); i += blockDim.x) d[i] = __ldg(s + i);
}

// sin / cos on the special-function unit (MUFU after the 1/2pi pre-scale); abs error ~1e-6 for |theta| < 64, far
// below the fp16 rounding of the activation it feeds
#ifdef BRIEF_EXP_NOSIN  // experiment only (tools/exp_variant.py): what does the kernel cost without the SFU work?
__device__ __forceinline__ float fast_sin(float x) { return x * 0.159f; }
__device__ __forceinline__ float fast_cos(float x) { return x * 0.161f; }
#else
__device__ __forceinline__ float fast_sin(float x) { return __sinf(x); }
__device__ __forceinline__ float fast_cos(float x) { return __cosf(x); }
#endif

// Sines of one 16-column chunk of a row whose first column is `col0`, for a network of width f (F_PAD >= f + 2):
// columns < f take the special-function unit; columns f, f+1 are the constant-one bias columns (sin(pi/2), written as
// the literal 1.0 — the same fp16 operand); the zero pads beyond stay zero.  f is CTA-uniform, so a chunk of real
// columns runs the plain unrolled loop and only the boundary chunk is predicated: at f = 56 (F_PAD = 64) this is
// 56 instead of 64 MUFU per row and layer.
__device__ __forceinline__ void sin_chunk16(float* v, int col0, int f) {
  if (col0 + 16 <= f) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fast_sin(v[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int col = col0 + i;
      v[i] = col < f ? fast_sin(v[i]) : (col < f + 2 ? 1.0f : 0.0f);
    }
  }
}

// dz = dX * scale * cos(theta) for one 16-column chunk; columns >= f (bias columns: cos(pi/2), pads: dX = 0) are zero
__device__ __forceinline__ void cos_mul_chunk16(float* z, const float* x, float scale, int col0, int f) {
  if (col0 + 16 <= f) {
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = x[i] * scale * fast_cos(z[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = col0 + i < f ? x[i] * scale * fast_cos(z[i]) : 0.0f;
  }
}

// first layer for 8 consecutive features of one sample -> 4 packed f16x2 words (optionally the raw z)
template <bool WITH_Z>
__device__ __forceinline__ uint4 first_layer8(const float4* __restrict__ w0b, int c0, float x0, float x1, float x2,
                                              float w0, float* zout) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 w = w0b[c0 + i];
    float z = w.w;
    z = fmaf(w.x, x0, z); z = fmaf(w.y, x1, z); z = fmaf(w.z, x2, z);
    if (WITH_Z) zout[i] = z;
    a[i] = fast_sin(w0 * z);
  }
  return make_uint4(pack_f16x2(a[0], a[1]), pack_f16x2(a[2], a[3]), pack_f16x2(a[4], a[5]), pack_f16x2(a[6], a[7]));
}

// issue z = act[128 x F] * W^T  (A K-major, B K-major) into TMEM columns [d, d+F)
template <int F>
__device__ __forceinline__ void issue_forward(uint32_t d, uint32_t act, uint32_t w) {
  constexpr uint32_t idesc = make_idesc(128, F, false, false);
#pragma unroll
  for (int k = 0; k < F / 16; ++k)
    mma_f16(d, make_desc(act + k * 2 * kActLBO, kActLBO, 128), make_desc(w + k * 2 * (F / 8) * 128, (F / 8) * 128, 128),
            idesc, k > 0);
}
// issue dX = dz[128 x F] * W  (A K-major, B = W seen MN-major: N = in feature, K = out feature)
template <int F>
__device__ __forceinline__ void issue_dx(uint32_t d, uint32_t dz, uint32_t w) {
  constexpr uint32_t idesc = make_idesc(128, F, false, true);
#pragma unroll
  for (int k = 0; k < F / 16; ++k)
    mma_f16(d, make_desc(dz + k * 2 * kActLBO, kActLBO, 128), make_desc(w + k * 2 * 128, 128, (F / 8) * 128), idesc, k > 0);
}
// the same two contractions with the A operand (activations / dz, fp16) in tensor memory
template <int F>
__device__ __forceinline__ void issue_forward_ts(uint32_t d, uint32_t a_tmem, uint32_t w) {
  constexpr uint32_t idesc = make_idesc(128, F, false, false);
#pragma unroll
  for (int k = 0; k < F / 16; ++k)
    mma_f16_ts(d, a_tmem + 8 * k, make_desc(w + k * 2 * (F / 8) * 128, (F / 8) * 128, 128), idesc, k > 0);
}
template <int F>
__device__ __forceinline__ void issue_dx_ts(uint32_t d, uint32_t a_tmem, uint32_t w) {
  constexpr uint32_t idesc = make_idesc(128, F, false, true);
#pragma unroll
  for (int k = 0; k < F / 16; ++k)
    mma_f16_ts(d, a_tmem + 8 * k, make_desc(w + k * 2 * 128, 128, (F / 8) * 128), idesc, k > 0);
}
// issue dW[64 x N] (+)= A^T B over the 128 samples of the tile (both MN-major, K = samples)
template <int N>
__device__ __forceinline__ void issue_dw(uint32_t d, uint32_t a_buf, uint32_t b_buf, bool accumulate) {
  constexpr uint32_t idesc = make_idesc(64, N, true, true);
#pragma unroll
  for (int k = 0; k < kTile / 16; ++k)
    mma_f16(d, make_desc(a_buf + k * 2 * 128, 128, kActLBO), make_desc(b_buf + k * 2 * 128, 128, kActLBO), idesc,
            (accumulate || k > 0) ? 1u : 0u);
}

// K-steps [k0, k1) of the same contraction (the MMA-issue warp of the fit kernel splits a dW into two halves so that a
// forward batch that becomes ready in between does not wait for all eight)
template <int N>
__device__ __forceinline__ void issue_dw_range(uint32_t d, uint32_t a_buf, uint32_t b_buf, bool accumulate, int k0, int k1) {
  constexpr uint32_t idesc = make_idesc(64, N, true, true);
  for (int k = k0; k < k1; ++k)
    mma_f16(d, make_desc(a_buf + k * 2 * 128, 128, kActLBO), make_desc(b_buf + k * 2 * 128, 128, kActLBO), idesc,
            (accumulate || k > 0) ? 1u : 0u);
}

// ==================================================================================================================
// operand-row stores shared by both kernels (thread = row r, chunk = 16 columns = two 16-byte core-matrix rows)
// ==================================================================================================================
// 16 fp32 -> 8 packed f16x2 words -> the operand row in shared memory and (optionally) the A-operand row in TMEM
template <bool SAT>
__device__ __forceinline__ void store_chunk16_both(unsigned char* buf, int r, int cg, const float* v, bool to_tmem,
                                                   uint32_t taddr) {
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = SAT ? pack_f16x2_sat(v[2 * i], v[2 * i + 1]) : pack_f16x2(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(buf + chunk_off(r, 2 * cg, kTile)) = make_uint4(w[0], w[1], w[2], w[3]);
  *reinterpret_cast<uint4*>(buf + chunk_off(r, 2 * cg + 1, kTile)) = make_uint4(w[4], w[5], w[6], w[7]);
  if (to_tmem) tmem_st8(taddr, w);
}
__device__ __forceinline__ void store_chunk16(unsigned char* buf, int r, int cg, const float* v) {
  *reinterpret_cast<uint4*>(buf + chunk_off(r, 2 * cg, kTile)) =
      make_uint4(pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]), pack_f16x2(v[4], v[5]), pack_f16x2(v[6], v[7]));
  *reinterpret_cast<uint4*>(buf + chunk_off(r, 2 * cg + 1, kTile)) =
      make_uint4(pack_f16x2(v[8], v[9]), pack_f16x2(v[10], v[11]), pack_f16x2(v[12], v[13]), pack_f16x2(v[14], v[15]));
}


// host side of the two translation units
size_t tc_wide_smem(int F, int L);
bool tc_wide_supported(int f, int L, int in_dim, int out_dim);
cudaError_t launch_tc_fit_wide(const FitArgs& a, int F_PAD, int L_max, int n_blocks, cudaStream_t st);

}  // namespace brief
