// brief_image.cuh — layout of the fp16 operand image of one tensor-core network (written by pack_kernel in
// brief_opt.cu after every optimiser step, staged into shared memory by ONE bulk copy per CTA in brief_tc.cu).
//
// F = F_PAD (multiple of 16, >= f + 2), NH = L - 2 hidden contractions.  All fp16 blocks use the UMMA "interleaved"
// (SWIZZLE_NONE, K-major) layout of brief_umma.cuh: element (row, k) of an R-row matrix at
//     ((k / 8) * (R / 8) + row / 8) * 128 + (row % 8) * 16 + (k % 8) * 2.
//
//   hidden [NH][F x F]   W'[o][k] = w_h * W[o][k]                       (o, k < f)
//                        W'[o][f] = hi(w_h * b[o]),  W'[o][f+1] = lo(w_h * b[o])   — the bias rides on two constant-one
//                        W'[f][f] = W'[f+1][f] = hi(pi/2), W'[.][f+1] = lo(pi/2)     activation columns f, f+1, which
//                        regenerate themselves: sin(pi/2) = 1.  (hi/lo: fp16 pair carrying ~22 bits.)
//   layer 0 [F x 16]     B operand of theta_0 = w_0 * (W0 x + b0) against the A row
//                        [x_hi(3) 1 x_lo(3) 1 x_hi(3) 0 0 0 0 0]:  cols 0-2 hi(w0 W0), 3 hi(w0 b0), 4-6 hi(w0 W0),
//                        7 lo(w0 b0), 8-10 lo(w0 W0); rows f, f+1: pi/2 in cols 3 / 7.  fp32-grade: the dropped
//                        lo*lo term is 2^-22 relative.
//   last  [16 x F]       row 0 = hi(Wlast), row 1 = lo(Wlast), column f carries blast: y = acc[0] + acc[1]
//   side (fp32)          float4 (W0x, W0y, W0z, b0) x F | w_h * b_l [NH][F] | Wlast [F] | blast, 0, 0, 0
#pragma once
#include <stddef.h>

namespace brief {

__host__ __device__ constexpr size_t img_hidden_bytes(int F, int NH) { return (size_t)NH * F * F * 2; }
__host__ __device__ constexpr size_t img_l0_off(int F, int NH) { return img_hidden_bytes(F, NH); }
__host__ __device__ constexpr size_t img_l0_bytes(int F) { return (size_t)F * 16 * 2; }
__host__ __device__ constexpr size_t img_last_off(int F, int NH) { return img_l0_off(F, NH) + img_l0_bytes(F); }
__host__ __device__ constexpr size_t img_last_bytes(int F) { return (size_t)16 * F * 2; }
__host__ __device__ constexpr size_t img_side_off(int F, int NH) { return img_last_off(F, NH) + img_last_bytes(F); }
__host__ __device__ constexpr size_t img_side_floats(int F, int NH) { return (size_t)4 * F + (size_t)NH * F + F + 4; }
__host__ __device__ constexpr size_t img_bytes(int F, int NH) { return img_side_off(F, NH) + img_side_floats(F, NH) * 4; }
__host__ __device__ constexpr size_t img_bytes_padded(int F, int NH) { return (img_bytes(F, NH) + 127) & ~(size_t)127; }
// byte offset of element (row, k) in an R-row interleaved K-major matrix
__host__ __device__ constexpr size_t img_elem_off(int row, int k, int R) {
  return ((size_t)(k >> 3) * (R >> 3) + (row >> 3)) * 128 + (row & 7) * 16 + (k & 7) * 2;
}

}  // namespace brief
