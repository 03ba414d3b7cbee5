// brief_opt.cu — per-network fused optimiser step for the whole group in ONE launch, plus the
// refresh of the fp16 operand image used by the tcgen05 kernels.
//
// Reference work replaced: torch.optim.Adamax / Adam / SGD .step() (configure_optimizer,
// utils/misc.py:174-183; main.py:399) — ~6 tiny ATen ops x 2L tensors per network per step — and
// optimizer.zero_grad() (main.py:387; gradients here are written, never accumulated).
//
// The arithmetic follows torch's single-tensor CPU kernels operation by operation:
//   Adamax : m = fma(1-b1, g-m, m)          (lerp_, vectorised form)
//            u = max(u*b2, |g|+eps)
//            p = p + ((-clr*m)/u)           (addcdiv_, clr = lr/(1-b1^t) computed in double on the host)
//   Adam   : m as above;  v = v*b2 + ((1-b2)*g)*g ;  den = sqrt(v)/sqrt(1-b2^t) + eps ;  p = p + ((-clr*m)/den)
//   SGD    : p = fma(-lr, g, p)
// When `partials` is given the kernel first reduces the per-slice gradient slots IN SLICE ORDER
// (deterministic, independent of how many networks share the GPU) — 28 B/param of optimiser traffic
// plus 4 B/param/slice of partial traffic; HBM/L2-bound and tiny next to the fit kernel.
#include "brief_common.cuh"
#include "brief_kernels.h"
#include "brief_image.cuh"
#include <cuda_fp16.h>

namespace brief {

constexpr int kOptThreads = 256;

__device__ __forceinline__ __half hi16(float x) { return __float2half_rn(x); }
__device__ __forceinline__ __half lo16(float x) { return __float2half_rn(__fsub_rn(x, __half2float(__float2half_rn(x)))); }

// Refresh the entries of the fp16 operand image (brief_image.cuh) that depend on parameter i (padded device layout) —
// the same values pack_kernel writes, so the optimiser step leaves the image current and the fit loop needs no
// separate pack launch.  Constant entries (pi/2 rows, zero pads) are written once by pack_kernel.
__device__ __forceinline__ void image_scatter(const NetDev& n, unsigned char* img, int i, float p) {
  const int F = n.F_PAD, NH = n.L - 2, f = n.f, F4 = n.F4;
  auto put = [&](size_t off, __half h) { *reinterpret_cast<__half*>(img + off) = h; };
  float* side = reinterpret_cast<float*>(img + img_side_off(F, NH));
  if (i < 4 * F4) {  // W0 [F4][4]
    const int o = i >> 2, c = i & 3;
    if (o >= f || c >= 3) return;
    const float v = __fmul_rn(n.w0, p);
    const size_t base = img_l0_off(F, NH);
    put(base + img_elem_off(o, c, F), hi16(v));
    put(base + img_elem_off(o, c + 4, F), hi16(v));
    put(base + img_elem_off(o, c + 8, F), lo16(v));
    side[4 * o + c] = p;
  } else if (i < 5 * F4) {  // b0
    const int o = i - 4 * F4;
    if (o >= f) return;
    const float v = __fmul_rn(n.w0, p);
    const size_t base = img_l0_off(F, NH);
    put(base + img_elem_off(o, 3, F), hi16(v));
    put(base + img_elem_off(o, 7, F), lo16(v));
    side[4 * o + 3] = p;
  } else if (i < dl_Wlast(n)) {  // hidden layers
    const int per = F4 * F4 + F4;
    const int j = i - 5 * F4, l = j / per, r = j - l * per;
    const size_t base = (size_t)l * F * F * 2;
    const float v = __fmul_rn(n.wh, p);
    if (r < F4 * F4) {
      const int o = r / F4, k = r - o * F4;
      if (o < f && k < f) put(base + img_elem_off(o, k, F), hi16(v));
    } else {
      const int o = r - F4 * F4;
      if (o >= f) return;
      put(base + img_elem_off(o, f, F), hi16(v));
      put(base + img_elem_off(o, f + 1, F), lo16(v));
      side[4 * F + l * F + o] = v;
    }
  } else {  // Wlast [F4], blast
    const int k = i < dl_blast(n) ? i - dl_Wlast(n) : (i == dl_blast(n) ? f : -1);
    if (k < 0 || (i < dl_blast(n) && k >= f)) return;
    const size_t base = img_last_off(F, NH);
    put(base + img_elem_off(0, k, 16), hi16(p));
    put(base + img_elem_off(1, k, 16), lo16(p));
    if (k < f) side[4 * F + NH * F + k] = p; else side[4 * F + NH * F + F] = p;
  }
}

__global__ void __launch_bounds__(kOptThreads) opt_kernel(OptArgs a) {
  // locate network: blk_prefix has n_nets+1 entries
  int lo = 0, hi = a.n_nets;
  const int b = blockIdx.x;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(a.blk_prefix + mid) <= b) lo = mid; else hi = mid;
  }
  const NetDev& n = a.nets[lo];
  const int blk = b - a.blk_prefix[lo];
  const int i = blk * kOptThreads + threadIdx.x;

  // per-network loss: ordered reduction of the slice partials by warp 0 of the network's first block
  if (blk == 0 && threadIdx.x < 32 && a.partials && a.loss_out) {
    float acc = 0.f;
    for (int s = threadIdx.x; s < n.n_slices; s += 32) acc += a.loss_partials[n.slice_off + s];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
    if (threadIdx.x == 0) a.loss_out[lo] = acc;
  }
  if (i >= n.P_dev) return;
  const long long gi = n.param_off + i;
  float g;
  if (a.partials) {
    const float* src = a.partials + n.part_off + i;
    g = 0.f;
    int s = 0;
    for (; s + 4 <= n.n_slices; s += 4) {
      const float g0 = src[(long long)(s + 0) * n.P_dev], g1 = src[(long long)(s + 1) * n.P_dev];
      const float g2 = src[(long long)(s + 2) * n.P_dev], g3 = src[(long long)(s + 3) * n.P_dev];
      g = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(g, g0), g1), g2), g3);
    }
    for (; s < n.n_slices; ++s) g = __fadd_rn(g, src[(long long)s * n.P_dev]);
    a.grads[gi] = g;
  } else {
    g = a.grads[gi];
  }
  if (!a.apply) return;
  const float neg_clr = a.state ? a.state->neg_clr : a.neg_clr;
  const float bc2_sqrt = a.state ? a.state->bc2_sqrt : a.bc2_sqrt;
  float p = a.params[gi];
  if (a.kind == 2) {  // SGD: param.add_(grad, alpha=-lr)
    p = fmaf(neg_clr, g, p);
  } else {
    float m = a.m[gi], v = a.v[gi];
    m = fmaf(a.w1, __fsub_rn(g, m), m);
    float den;
    if (a.kind == 0) {  // Adamax
      v = fmaxf(__fmul_rn(v, a.beta2), __fadd_rn(fabsf(g), a.eps));
      den = v;
    } else {  // Adam
      v = __fadd_rn(__fmul_rn(v, a.beta2), __fmul_rn(__fmul_rn(a.w2, g), g));
      den = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), a.eps);
    }
    p = __fadd_rn(p, __fdiv_rn(__fmul_rn(neg_clr, m), den));
    a.m[gi] = m;
    a.v[gi] = v;
  }
  a.params[gi] = p;
  if (a.wpack && n.eval_tc) image_scatter(n, a.wpack + n.wpack_off, i, p);
}

cudaError_t launch_opt(const OptArgs& a, int n_blocks, cudaStream_t st) {
  opt_kernel<<<n_blocks, kOptThreads, 0, st>>>(a);
  return cudaGetLastError();
}

// ---- fp16 operand image for the tcgen05 kernels (layout: brief_image.cuh) -----------------------------------------

__global__ void pack_kernel(const NetDev* nets, int n_nets, const float* __restrict__ params, unsigned char* wpack) {
  constexpr float kHalfPi = 1.57079632679f;
  for (int net_id = blockIdx.y; net_id < n_nets; net_id += gridDim.y) {
  const NetDev& n = nets[net_id];
  if (!n.eval_tc) continue;
  const int F = n.F_PAD, NH = n.L - 2, f = n.f, F4 = n.F4;
  const float* P = params + n.param_off;
  unsigned char* img = wpack + n.wpack_off;
  const int n_hidden = NH * F * F, n_l0 = F * 16, n_last = 16 * F;
  const int n_side = (int)img_side_floats(F, NH);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_hidden + n_l0 + n_last + n_side; i += gridDim.x * blockDim.x) {
    if (i < n_hidden) {
      const int l = i / (F * F), r = i - l * F * F;
      const int o = r / F, k = r - o * F;
      // hidden weights are stored pre-multiplied by the hidden omega: the MMA then yields the sine argument directly
      // (bias included, through the two constant-one columns), and the dX contraction yields w * dz W
      __half h = __float2half_rn(0.f);
      if (o < f) {
        if (k < f) h = hi16(__fmul_rn(n.wh, P[dl_W(n, l + 1) + o * F4 + k]));
        else if (k == f) h = hi16(__fmul_rn(n.wh, P[dl_b(n, l + 1) + o]));
        else if (k == f + 1) h = lo16(__fmul_rn(n.wh, P[dl_b(n, l + 1) + o]));
      } else if (o <= f + 1) {
        if (k == f) h = hi16(kHalfPi);
        else if (k == f + 1) h = lo16(kHalfPi);
      }
      *reinterpret_cast<__half*>(img + (size_t)l * F * F * 2 + img_elem_off(o, k, F)) = h;
    } else if (i < n_hidden + n_l0) {
      const int r = i - n_hidden, o = r >> 4, c = r & 15;
      __half h = __float2half_rn(0.f);
      if (o < f) {
        const int d = c & 3;
        const float v = d < 3 ? __fmul_rn(n.w0, P[dl_W0(n) + 4 * o + d]) : __fmul_rn(n.w0, P[dl_b0(n) + o]);
        if (c < 3 || c == 3 || (c >= 4 && c < 7)) h = hi16(v);
        else if (c == 7 || (c >= 8 && c < 11)) h = lo16(v);
      } else if (o <= f + 1) {
        if (c == 3) h = hi16(kHalfPi);
        else if (c == 7) h = lo16(kHalfPi);
      }
      *reinterpret_cast<__half*>(img + img_l0_off(F, NH) + img_elem_off(o, c, F)) = h;
    } else if (i < n_hidden + n_l0 + n_last) {
      const int q = i - n_hidden - n_l0, r = q / F, k = q - r * F;
      __half h = __float2half_rn(0.f);
      if (r < 2 && k <= f) {
        const float v = k < f ? P[dl_Wlast(n) + k] : P[dl_blast(n)];
        h = r == 0 ? hi16(v) : lo16(v);
      }
      *reinterpret_cast<__half*>(img + img_last_off(F, NH) + img_elem_off(r, k, 16)) = h;
    } else {
      const int j = i - n_hidden - n_l0 - n_last;
      float* side = reinterpret_cast<float*>(img + img_side_off(F, NH));
      float val;
      if (j < 4 * F) {
        const int o = j >> 2, c = j & 3;
        val = (o < f) ? (c < 3 ? P[dl_W0(n) + 4 * o + c] : P[dl_b0(n) + o])
                      : (o <= f + 1 && c == 3 ? __fdiv_rn(kHalfPi, n.w0) : 0.f);
      } else if (j < 4 * F + NH * F) {
        const int q = j - 4 * F, l = q / F, o = q - l * F;
        val = (o < f) ? __fmul_rn(n.wh, P[dl_b(n, l + 1) + o]) : (o <= f + 1 ? kHalfPi : 0.f);
      } else if (j < 4 * F + NH * F + F) {
        const int k = j - 4 * F - NH * F;
        val = (k < f) ? P[dl_Wlast(n) + k] : 0.f;
      } else {
        const int c = j - (4 * F + NH * F + F);
        val = c == 0 ? P[dl_blast(n)] : 0.f;
      }
      side[j] = val;
    }
  }
  }
}

cudaError_t launch_pack(const NetDev* nets, int n_nets, const float* params, unsigned char* wpack, cudaStream_t st) {
  dim3 grid(8, n_nets < 32768 ? n_nets : 32768);
  pack_kernel<<<grid, 256, 0, st>>>(nets, n_nets, params, wpack);
  return cudaGetLastError();
}

}  // namespace brief
