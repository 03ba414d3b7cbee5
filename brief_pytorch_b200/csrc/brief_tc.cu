// brief_tc.cu — the tensor-core path (BRIEF_PREC_F16): grouped, fully fused SIREN kernels on tcgen05.
//
// Reference work replaced (file:line relative to the reference root):
//   SIREN.forward / reconstruct_flattened / invnormalize_data   utils/Networks.py:269-271, utils/misc.py:59-92,
//                                                               utils/io.py:136-147            -> tc_eval_kernel
//   sampler gather + forward + datal2 + backward               main.py:126-163, 176-182, 385-396 -> tc_fit_kernel
//
// Shape of the computation.  One CTA = 128 threads = one tile of 128 samples (UMMA M = 128): thread t owns sample
// row t == TMEM lane t.  The first layer (K = 3) and the last layer (N = 1) run on CUDA cores in fp32; every
// hidden layer z_l = a_{l-1} W_l^T is a tcgen05.mma (fp16 operands from shared memory, fp32 accumulator in TMEM),
// followed by an epilogue that reads the accumulator row with tcgen05.ld, applies sin(w*z + w*b) and writes the
// fp16 activations straight back into the shared-memory operand buffer of the next layer.  The network's weights
// are staged ONCE per CTA by one bulk (TMA) copy of the packed fp16 image and stay resident for all its tiles.
//
// Why fp16 and not bf16: activations are sines (|a| <= 1) and weights are << 1, so fp16's 11-bit significand is
// usable without range problems and gives 8x smaller rounding error than bf16 at the same tensor rate.  A and B of
// one tcgen05.mma must share a format (mixing is an illegal instruction), so the backward operands are fp16 too:
// dz is carried as dz * kGradScale * (B*C)/2 — i.e. w*(yhat-y)/256 at the output — which keeps it inside fp16's
// normal range with orders of magnitude to spare on both sides; conversions saturate instead of overflowing and
// the scale is removed in fp32 when the accumulated dW leaves TMEM.
//
// Operand layout: brief_umma.cuh ("interleaved" 8x8 cores).  Width f is padded to F = F_PAD (multiple of 16,
// F > f); pad weights are zero.  Column f of every activation buffer is the constant 1 (its packed "bias" makes
// the sine argument pi/2), so bias gradients fall out of the dW contractions as column f.
//
// Fit kernel, per tile:   forward  NH x [ MMA z_l ; sin ]            (activations a_0..a_NH stay in shared memory)
//                         backward NH x [ MMA dW_l += dz_l^T a_{l-1} ; MMA z_{l-1} (recomputed) ; MMA dX = dz_l W_l ;
//                                         dz_{l-1} = dX * w * cos(w z_{l-1}) ]
// dW_l accumulate ACROSS the tiles of a slice in TMEM (M = 64 accumulators, F columns per layer) and are written
// once per slice to the slice's gradient-partial slot; the optimiser kernel reduces slots in fixed order.
//
// Roofline: per sample a hidden layer costs 2*F*F tensor FLOPs and F special-function ops (2F in the fit).  At
// F <= 64 the MUFU pipe (16 ops/clk/SM), not the tensor pipe, is the binding unit.
#include <cuda_fp16.h>

#include "brief_common.cuh"
#include "brief_kernels.h"
#include "brief_umma.cuh"

namespace brief {

using namespace umma;

constexpr int kTcThreads = 128;
constexpr int kTile = 128;
constexpr int kEvalTilesPerBlock = 64;     // must match kTcEvalTilesPerBlock in brief_capi.cu
constexpr float kGradScale = 1.0f / 256.0f;  // dy' = kGradScale * w * (yhat - y)
constexpr uint32_t kActLBO = (kTile / 8) * 128;  // 2048: feature-group stride of a [128 x F] operand buffer

__host__ __device__ constexpr int tmem_cols_pow2(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

// ---- packed image (written by pack_kernel, brief_opt.cu) ------------------------------------------------------
//   [NH][F*F] fp16 hidden weights (interleaved, R = F)  |  fp32 side block:
//   float4 (W0x, W0y, W0z, b0) x F | w_hidden * b_l [NH][F] | Wlast [F] | blast, 0, 0, 0
__host__ __device__ constexpr size_t img_hidden_bytes(int F, int NH) { return (size_t)NH * F * F * 2; }
__host__ __device__ constexpr size_t img_side_floats(int F, int NH) { return (size_t)4 * F + (size_t)NH * F + F + 4; }
__host__ __device__ constexpr size_t img_bytes(int F, int NH) { return img_hidden_bytes(F, NH) + img_side_floats(F, NH) * 4; }
__host__ __device__ constexpr size_t img_bytes_padded(int F, int NH) { return (img_bytes(F, NH) + 127) & ~(size_t)127; }

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
__device__ __forceinline__ float fast_sin(float x) { return __sinf(x); }
__device__ __forceinline__ float fast_cos(float x) { return __cosf(x); }

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
// issue dW[64 x N] (+)= A^T B over the 128 samples of the tile (both MN-major, K = samples)
template <int N>
__device__ __forceinline__ void issue_dw(uint32_t d, uint32_t a_buf, uint32_t b_buf, bool accumulate) {
  constexpr uint32_t idesc = make_idesc(64, N, true, true);
#pragma unroll
  for (int k = 0; k < kTile / 16; ++k)
    mma_f16(d, make_desc(a_buf + k * 2 * 128, 128, kActLBO), make_desc(b_buf + k * 2 * 128, 128, kActLBO), idesc,
            (accumulate || k > 0) ? 1u : 0u);
}

// ==================================================================================================================
// forward / decompress
// ==================================================================================================================
template <int F>
__global__ void __launch_bounds__(kTcThreads) tc_eval_kernel(EvalArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ NetDev sn;
  __shared__ __align__(8) uint64_t bar_w, bar_mma;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) unsigned short s_out[kTile];

  const int t = threadIdx.x, warp = t >> 5;
  int net_id;
  long long chunk;
  if (a.single_net >= 0) {
    net_id = a.single_net;
    chunk = blockIdx.x;
  } else {
    const int wi = tc_find_work(a.work_prefix, a.n_work, blockIdx.x);
    net_id = a.work_net[wi];
    chunk = blockIdx.x - a.work_prefix[wi];
  }
  tc_load_net(sn, a.nets[net_id]);
  if (t == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  constexpr int TCOLS = tmem_cols_pow2(F);
  if (warp == 0) tmem_alloc<TCOLS>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const NetDev& n = sn;
  const int NH = n.L - 2;
  unsigned char* sW = smem;                                             // packed image (hidden weights + side)
  const float* side = reinterpret_cast<const float*>(sW + img_hidden_bytes(F, NH));
  const float4* s_w0b = reinterpret_cast<const float4*>(side);          // [F]
  const float* s_wb = side + 4 * F;                                     // [NH][F]
  const float* s_wl = s_wb + NH * F;                                    // [F]
  const float* s_bl = s_wl + F;                                         // [4]
  unsigned char* sAct = sW + img_bytes_padded(F, NH);                   // [128 x F] fp16 interleaved
  if (t == 0) {
    const uint32_t bytes = (uint32_t)img_bytes(F, NH);
    mbar_expect_tx(&bar_w, bytes);
    bulk_g2s(sW, a.wpack + n.wpack_off, bytes, &bar_w);
  }
  mbar_wait(&bar_w, 0);

  const uint32_t tm = tmem_base_s;
  const uint32_t lane_addr = tm + ((uint32_t)(warp * 32) << 16);
  const uint32_t aAct = smem_u32(sAct), aW = smem_u32(sW);
  const long long total = a.coords ? a.n_coords : n.n_vox;
  const long long n_tiles = (total + kTile - 1) / kTile;
  const long long tile_end = min(n_tiles, (chunk + 1) * (long long)kEvalTilesPerBlock);
  uint32_t phase = 0;
  const float wh = n.wh;

  for (long long tile = chunk * kEvalTilesPerBlock; tile < tile_end; ++tile) {
    const long long s = tile * kTile + t;
    const bool valid = s < total;
    float x0 = 0.f, x1 = 0.f, x2 = 0.f;
    if (valid) {
      if (a.coords) {
        x0 = a.coords[s * n.in_dim];
        x1 = a.coords[s * n.in_dim + 1];
        x2 = n.in_dim == 3 ? a.coords[s * n.in_dim + 2] : 0.f;
      } else {
        brief_coords(n, a.axes, s, x0, x1, x2);
      }
    }
    // ---- layer 0 on CUDA cores -> fp16 operand rows
#pragma unroll
    for (int cg = 0; cg < F / 8; ++cg) {
      float z[8];
      uint4 pk;
      if (a.layers_out) {
        pk = first_layer8<true>(s_w0b, cg * 8, x0, x1, x2, n.w0, z);
        if (valid)
          for (int i = 0; i < 8; ++i)
            if (cg * 8 + i < n.f) a.layers_out[s * n.f + cg * 8 + i] = z[i];
      } else {
        pk = first_layer8<false>(s_w0b, cg * 8, x0, x1, x2, n.w0, z);
      }
      *reinterpret_cast<uint4*>(sAct + chunk_off(t, cg, kTile)) = pk;
    }
    float y = s_bl[0];
    // ---- hidden layers on the tensor core
    for (int l = 1; l <= NH; ++l) {
      fence_async_smem();
      __syncthreads();
      if (t == 0) {
        tc_fence_after();
        issue_forward<F>(tm, aAct, aW + (uint32_t)(l - 1) * F * F * 2);
        commit(&bar_mma);
      }
      mbar_wait(&bar_mma, phase);
      phase ^= 1;
      tc_fence_after();
      const float* wb = s_wb + (l - 1) * F;
      const bool last = l == NH;
      float* zdump = a.layers_out ? a.layers_out + (long long)l * total * n.f + s * n.f : nullptr;
      float v[2][16];
      tmem_ld16(lane_addr, v[0]);
#pragma unroll
      for (int ci = 0; ci < F / 16; ++ci) {
        tmem_ld_wait();
        if (ci + 1 < F / 16) tmem_ld16(lane_addr + (ci + 1) * 16, v[(ci + 1) & 1]);
        float* vv = v[ci & 1];
        float act[16];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(wb + ci * 16 + i);
          const float th0 = fmaf(vv[i], wh, b4.x), th1 = fmaf(vv[i + 1], wh, b4.y);
          const float th2 = fmaf(vv[i + 2], wh, b4.z), th3 = fmaf(vv[i + 3], wh, b4.w);
          if (zdump && valid) {
            const float th[4] = {th0, th1, th2, th3};
            for (int j = 0; j < 4; ++j)
              if (ci * 16 + i + j < n.f) zdump[ci * 16 + i + j] = th[j] / wh;
          }
          act[i] = fast_sin(th0); act[i + 1] = fast_sin(th1); act[i + 2] = fast_sin(th2); act[i + 3] = fast_sin(th3);
        }
        if (!last) {
          *reinterpret_cast<uint4*>(sAct + chunk_off(t, 2 * ci, kTile)) =
              make_uint4(pack_f16x2(act[0], act[1]), pack_f16x2(act[2], act[3]), pack_f16x2(act[4], act[5]),
                         pack_f16x2(act[6], act[7]));
          *reinterpret_cast<uint4*>(sAct + chunk_off(t, 2 * ci + 1, kTile)) =
              make_uint4(pack_f16x2(act[8], act[9]), pack_f16x2(act[10], act[11]), pack_f16x2(act[12], act[13]),
                         pack_f16x2(act[14], act[15]));
        } else {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(s_wl + ci * 16 + i);
            y = fmaf(w4.x, act[i], y); y = fmaf(w4.y, act[i + 1], y);
            y = fmaf(w4.z, act[i + 2], y); y = fmaf(w4.w, act[i + 3], y);
          }
        }
      }
      tc_fence_before();
    }
    // ---- epilogue: inverse normalisation + truncating cast + store
    if (a.out_f32) {
      if (valid) a.out_f32[s] = y;
    } else {
      void* dst = a.out_ptrs[net_id];
      if (a.out_dtype == 2) {
        if (valid) reinterpret_cast<float*>(dst)[s] = y;
      } else {
        const float vden = brief_denorm(n, y);
        const bool full = tile * kTile + kTile <= total;
        if (a.out_dtype == 1) {
          if (full) {  // stage the tile's 128 values, 16 threads store 16 B each (256 B contiguous)
            s_out[t] = (unsigned short)(int)vden;
            __syncthreads();
            if (t < 16)
              reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(dst) + tile * kTile)[t] =
                  reinterpret_cast<const uint4*>(s_out)[t];
          } else if (valid) {
            reinterpret_cast<unsigned short*>(dst)[s] = (unsigned short)(int)vden;
          }
        } else {
          if (full) {
            reinterpret_cast<unsigned char*>(s_out)[t] = (unsigned char)(int)vden;
            __syncthreads();
            if (t < 8)
              reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(dst) + tile * kTile)[t] =
                  reinterpret_cast<const uint4*>(s_out)[t];
          } else if (valid) {
            reinterpret_cast<unsigned char*>(dst)[s] = (unsigned char)(int)vden;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TCOLS>(tm);
}

// ==================================================================================================================
// fit: gather + forward + weighted L2 + backward -> per-slice gradient partials
// ==================================================================================================================
// shared memory (dynamic):  sDz[2] | sAct[NS] | sX | sDY | image        (NS = L-1 sine layers)
//   sDz first: the M = 64 dW contractions read 8 feature groups (16 KB) from the start of a dz buffer whatever F is.
// TMEM columns: Z [0,F) | X [F,2F) | dW_l [2F + (l-1)F, +F) l=1..NH | dW0 [.., +16) | dWlast [.., +16)
template <int F>
__global__ void __launch_bounds__(kTcThreads, 1) tc_fit_kernel(FitArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ NetDev sn;
  __shared__ __align__(8) uint64_t bar_w, bar_mma;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_red[4];

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int wi = tc_find_work(a.work_prefix, a.n_work, blockIdx.x);
  const int net_id = a.work_net[wi];
  const int slice = blockIdx.x - a.work_prefix[wi];
  tc_load_net(sn, a.nets[net_id]);
  if (t == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const NetDev& n = sn;
  const int NH = n.L - 2, NS = n.L - 1, f = n.f;
  constexpr uint32_t BUF = kTile * F * 2;  // one [128 x F] fp16 buffer
  unsigned char* sDz = smem;
  unsigned char* sAct = sDz + 2 * BUF;
  unsigned char* sX = sAct + (size_t)NS * BUF;
  unsigned char* sDY = sX + kTile * 16 * 2;
  unsigned char* sW = sDY + kTile * 16 * 2;
  const float* side = reinterpret_cast<const float*>(sW + img_hidden_bytes(F, NH));
  const float4* s_w0b = reinterpret_cast<const float4*>(side);
  const float* s_wb = side + 4 * F;
  const float* s_wl = s_wb + NH * F;
  const float* s_bl = s_wl + F;
  if (t == 0) {
    const uint32_t bytes = (uint32_t)img_bytes(F, NH);
    mbar_expect_tx(&bar_w, bytes);
    bulk_g2s(sW, a.wpack + n.wpack_off, bytes, &bar_w);
  }
  // second column group of the two 16-column blocks is constant zero
  *reinterpret_cast<uint4*>(sX + chunk_off(t, 1, kTile)) = make_uint4(0, 0, 0, 0);
  *reinterpret_cast<uint4*>(sDY + chunk_off(t, 1, kTile)) = make_uint4(0, 0, 0, 0);
  mbar_wait(&bar_w, 0);

  const uint32_t tm = tmem_base_s;
  const uint32_t lane_addr = tm + ((uint32_t)(warp * 32) << 16);
  const uint32_t TZ = tm, TX = tm + F, TDW = tm + 2 * F, TDW0 = TDW + NH * F, TDWL = TDW0 + 16;
  const uint32_t aDz = smem_u32(sDz), aAct = smem_u32(sAct), aX = smem_u32(sX), aDY = smem_u32(sDY), aW = smem_u32(sW);
  uint32_t phase = 0;
  const float wh = n.wh, w0 = n.w0;
  const long long s_begin = (long long)slice * n.slice_len;
  const long long s_end = min((long long)n.batch, s_begin + n.slice_len);
  float loss_acc = 0.f;
  int it = 0;

  for (long long tile0 = s_begin; tile0 < s_end; tile0 += kTile, ++it) {
    const long long s = tile0 + t;
    const bool valid = s < s_end;
    float x0 = 0.f, x1 = 0.f, x2 = 0.f, yv = 0.f, wv = 0.f;
    if (valid) {
      long long idx;
      if (n.mode == 0) idx = s;
      else if (a.idx) idx = a.idx[n.idx_off + s];
      else idx = brief_sample_index(a.seed, a.step, (uint32_t)net_id, (uint64_t)s, (uint64_t)n.n_vox);
      brief_coords(n, a.axes, idx, x0, x1, x2);
      const float raw = brief_raw_value(n, idx);
      yv = brief_normalize(n, raw);
      wv = brief_weight(n, idx, raw);
    }
    {  // B operand of the dW0 contraction: [x_hi(3), 1, x_lo(3), 0]  (hi/lo split keeps fp32-grade coordinates)
      const float h0 = __half2float(__float2half_rn(x0)), h1 = __half2float(__float2half_rn(x1)),
                  h2 = __half2float(__float2half_rn(x2));
      *reinterpret_cast<uint4*>(sX + chunk_off(t, 0, kTile)) =
          make_uint4(pack_f16x2(h0, h1), pack_f16x2(h2, 1.0f), pack_f16x2(x0 - h0, x1 - h1), pack_f16x2(x2 - h2, 0.f));
    }
    // ---- forward, layer 0 (CUDA cores)
#pragma unroll
    for (int cg = 0; cg < F / 8; ++cg) {
      float z[8];
      *reinterpret_cast<uint4*>(sAct + chunk_off(t, cg, kTile)) = first_layer8<false>(s_w0b, cg * 8, x0, x1, x2, w0, z);
    }
    // ---- forward, hidden layers (tensor core); a_l -> sAct[l]
    float y = s_bl[0];
    for (int l = 1; l <= NH; ++l) {
      tc_fence_before();
      fence_async_smem();
      __syncthreads();
      if (t == 0) {
        tc_fence_after();
        issue_forward<F>(TZ, aAct + (uint32_t)(l - 1) * BUF, aW + (uint32_t)(l - 1) * F * F * 2);
        commit(&bar_mma);
      }
      mbar_wait(&bar_mma, phase);
      phase ^= 1;
      tc_fence_after();
      const float* wb = s_wb + (l - 1) * F;
      unsigned char* dstA = sAct + (size_t)l * BUF;
      const bool last = l == NH;
      float v[2][16];
      tmem_ld16(lane_addr, v[0]);
#pragma unroll
      for (int ci = 0; ci < F / 16; ++ci) {
        tmem_ld_wait();
        if (ci + 1 < F / 16) tmem_ld16(lane_addr + (ci + 1) * 16, v[(ci + 1) & 1]);
        float* vv = v[ci & 1];
        float act[16];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(wb + ci * 16 + i);
          act[i] = fast_sin(fmaf(vv[i], wh, b4.x)); act[i + 1] = fast_sin(fmaf(vv[i + 1], wh, b4.y));
          act[i + 2] = fast_sin(fmaf(vv[i + 2], wh, b4.z)); act[i + 3] = fast_sin(fmaf(vv[i + 3], wh, b4.w));
        }
        *reinterpret_cast<uint4*>(dstA + chunk_off(t, 2 * ci, kTile)) =
            make_uint4(pack_f16x2(act[0], act[1]), pack_f16x2(act[2], act[3]), pack_f16x2(act[4], act[5]),
                       pack_f16x2(act[6], act[7]));
        *reinterpret_cast<uint4*>(dstA + chunk_off(t, 2 * ci + 1, kTile)) =
            make_uint4(pack_f16x2(act[8], act[9]), pack_f16x2(act[10], act[11]), pack_f16x2(act[12], act[13]),
                       pack_f16x2(act[14], act[15]));
        if (last) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(s_wl + ci * 16 + i);
            y = fmaf(w4.x, act[i], y); y = fmaf(w4.y, act[i + 1], y);
            y = fmaf(w4.z, act[i + 2], y); y = fmaf(w4.w, act[i + 3], y);
          }
        }
      }
    }
    // ---- loss (datal2, main.py:176-182) and scaled output gradient
    float dys = 0.f;
    if (valid) {
      const float e = y - yv;
      const float wt = (n.tau != 0.f && y <= n.tau) ? 1.0f : wv;
      loss_acc = fmaf(wt * e, e, loss_acc);
      dys = kGradScale * wt * e;
    }
    *reinterpret_cast<uint4*>(sDY + chunk_off(t, 0, kTile)) = make_uint4(pack_f16x2_sat(dys, 0.f), 0, 0, 0);
    // dz_NH = dy * Wlast * w * cos(w z_NH): second pass over the accumulator that is still in TMEM
    {
      const float* wb = s_wb + (NH - 1) * F;
      float v[2][16];
      tmem_ld16(lane_addr, v[0]);
#pragma unroll
      for (int ci = 0; ci < F / 16; ++ci) {
        tmem_ld_wait();
        if (ci + 1 < F / 16) tmem_ld16(lane_addr + (ci + 1) * 16, v[(ci + 1) & 1]);
        float* vv = v[ci & 1];
        float dz[16];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(wb + ci * 16 + i);
          const float4 w4 = *reinterpret_cast<const float4*>(s_wl + ci * 16 + i);
          dz[i] = dys * w4.x * wh * fast_cos(fmaf(vv[i], wh, b4.x));
          dz[i + 1] = dys * w4.y * wh * fast_cos(fmaf(vv[i + 1], wh, b4.y));
          dz[i + 2] = dys * w4.z * wh * fast_cos(fmaf(vv[i + 2], wh, b4.z));
          dz[i + 3] = dys * w4.w * wh * fast_cos(fmaf(vv[i + 3], wh, b4.w));
        }
        *reinterpret_cast<uint4*>(sDz + chunk_off(t, 2 * ci, kTile)) =
            make_uint4(pack_f16x2_sat(dz[0], dz[1]), pack_f16x2_sat(dz[2], dz[3]), pack_f16x2_sat(dz[4], dz[5]),
                       pack_f16x2_sat(dz[6], dz[7]));
        *reinterpret_cast<uint4*>(sDz + chunk_off(t, 2 * ci + 1, kTile)) =
            make_uint4(pack_f16x2_sat(dz[8], dz[9]), pack_f16x2_sat(dz[10], dz[11]), pack_f16x2_sat(dz[12], dz[13]),
                       pack_f16x2_sat(dz[14], dz[15]));
      }
    }
    // ---- backward through the hidden layers
    int cur = 0;
    for (int l = NH; l >= 1; --l) {
      tc_fence_before();
      fence_async_smem();
      __syncthreads();
      if (t == 0) {
        tc_fence_after();
        const uint32_t dzb = aDz + (uint32_t)cur * BUF;
        if (l == NH) issue_dw<16>(TDWL, aAct + (uint32_t)NH * BUF, aDY, it > 0);       // dWlast, dblast
        issue_dw<F>(TDW + (uint32_t)(l - 1) * F, dzb, aAct + (uint32_t)(l - 1) * BUF, it > 0);  // dW_l, db_l
        if (l >= 2) issue_forward<F>(TZ, aAct + (uint32_t)(l - 2) * BUF, aW + (uint32_t)(l - 2) * F * F * 2);  // z_{l-1}
        issue_dx<F>(TX, dzb, aW + (uint32_t)(l - 1) * F * F * 2);                      // dX_{l-1}
        commit(&bar_mma);
      }
      mbar_wait(&bar_mma, phase);
      phase ^= 1;
      tc_fence_after();
      unsigned char* dstZ = sDz + (size_t)(cur ^ 1) * BUF;
      if (l >= 2) {
        const float* wb = s_wb + (l - 2) * F;
        float vz[16], vx[16];
#pragma unroll
        for (int ci = 0; ci < F / 16; ++ci) {
          tmem_ld16(lane_addr + ci * 16, vz);
          tmem_ld16(lane_addr + F + ci * 16, vx);
          tmem_ld_wait();
          float dz[16];
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(wb + ci * 16 + i);
            dz[i] = vx[i] * wh * fast_cos(fmaf(vz[i], wh, b4.x));
            dz[i + 1] = vx[i + 1] * wh * fast_cos(fmaf(vz[i + 1], wh, b4.y));
            dz[i + 2] = vx[i + 2] * wh * fast_cos(fmaf(vz[i + 2], wh, b4.z));
            dz[i + 3] = vx[i + 3] * wh * fast_cos(fmaf(vz[i + 3], wh, b4.w));
          }
          *reinterpret_cast<uint4*>(dstZ + chunk_off(t, 2 * ci, kTile)) =
              make_uint4(pack_f16x2_sat(dz[0], dz[1]), pack_f16x2_sat(dz[2], dz[3]), pack_f16x2_sat(dz[4], dz[5]),
                         pack_f16x2_sat(dz[6], dz[7]));
          *reinterpret_cast<uint4*>(dstZ + chunk_off(t, 2 * ci + 1, kTile)) =
              make_uint4(pack_f16x2_sat(dz[8], dz[9]), pack_f16x2_sat(dz[10], dz[11]), pack_f16x2_sat(dz[12], dz[13]),
                         pack_f16x2_sat(dz[14], dz[15]));
        }
      } else {  // l == 1: dz_0 = dX_0 * w0 * cos(w0 z_0), z_0 recomputed on CUDA cores
        float vx[16];
#pragma unroll
        for (int ci = 0; ci < F / 16; ++ci) {
          tmem_ld16(lane_addr + F + ci * 16, vx);
          tmem_ld_wait();
          float dz[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 w = s_w0b[ci * 16 + i];
            float z = w.w;
            z = fmaf(w.x, x0, z); z = fmaf(w.y, x1, z); z = fmaf(w.z, x2, z);
            dz[i] = vx[i] * w0 * fast_cos(w0 * z);
          }
          *reinterpret_cast<uint4*>(dstZ + chunk_off(t, 2 * ci, kTile)) =
              make_uint4(pack_f16x2_sat(dz[0], dz[1]), pack_f16x2_sat(dz[2], dz[3]), pack_f16x2_sat(dz[4], dz[5]),
                         pack_f16x2_sat(dz[6], dz[7]));
          *reinterpret_cast<uint4*>(dstZ + chunk_off(t, 2 * ci + 1, kTile)) =
              make_uint4(pack_f16x2_sat(dz[8], dz[9]), pack_f16x2_sat(dz[10], dz[11]), pack_f16x2_sat(dz[12], dz[13]),
                         pack_f16x2_sat(dz[14], dz[15]));
        }
      }
      cur ^= 1;
    }
    // ---- dW0 += dz_0^T [x_hi, 1, x_lo]; wait so that the next tile may overwrite sX / sDz / sAct
    tc_fence_before();
    fence_async_smem();
    __syncthreads();
    if (t == 0) {
      tc_fence_after();
      issue_dw<16>(TDW0, aDz + (uint32_t)cur * BUF, aX, it > 0);
      commit(&bar_mma);
    }
    mbar_wait(&bar_mma, phase);
    phase ^= 1;
    tc_fence_after();
  }

  // ---- slice epilogue: loss partial + gradient partials (TMEM -> global), scale removed in fp32
  const float inv_count = 1.0f / ((float)n.batch * (float)n.out_dim);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) loss_acc += __shfl_down_sync(0xffffffffu, loss_acc, off);
  if (lane == 0) s_red[warp] = loss_acc;
  float* part = a.partials + n.part_off + (long long)slice * n.P_dev;
  for (int i = t; i < n.P_dev; i += kTcThreads) part[i] = 0.f;
  __syncthreads();
  if (t == 0) a.loss_partials[n.slice_off + slice] = (((s_red[0] + s_red[1]) + s_red[2]) + s_red[3]) * inv_count;
  if (it > 0) {
    const float unscale = 2.0f * inv_count / kGradScale;
    const int o = warp * 16 + lane;  // accumulator row held by this thread (M = 64 layout: lanes 0..15 of each warp)
    const bool row_ok = lane < 16 && o < F;
    const int F4 = n.F4;
    float v[16];
    for (int l = 1; l <= NH; ++l) {
#pragma unroll
      for (int ci = 0; ci < F / 16; ++ci) {
        tmem_ld16(lane_addr + 2 * F + (l - 1) * F + ci * 16, v);
        tmem_ld_wait();
        if (row_ok && o < f) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int k = ci * 16 + i;
            if (k < f) part[dl_W(n, l) + o * F4 + k] = v[i] * unscale;
            else if (k == f) part[dl_b(n, l) + o] = v[i] * unscale;
          }
        }
      }
    }
    tmem_ld16(lane_addr + 2 * F + NH * F, v);  // dW0 block
    tmem_ld_wait();
    if (row_ok && o < f) {
      part[dl_W0(n) + 4 * o + 0] = (v[0] + v[4]) * unscale;
      part[dl_W0(n) + 4 * o + 1] = (v[1] + v[5]) * unscale;
      if (n.in_dim == 3) part[dl_W0(n) + 4 * o + 2] = (v[2] + v[6]) * unscale;
      part[dl_b0(n) + o] = v[3] * unscale;
    }
    tmem_ld16(lane_addr + 2 * F + NH * F + 16, v);  // dWlast block: row = feature of a_NH, column 0
    tmem_ld_wait();
    if (row_ok) {
      if (o < f) part[dl_Wlast(n) + o] = v[0] * unscale;
      else if (o == f) part[dl_blast(n)] = v[0] * unscale;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}

// ==================================================================================================================
// host side
// ==================================================================================================================
int tc_fpad(int f) { return ((f + 1 + 15) / 16) * 16; }
size_t tc_wpack_bytes(int F, int L) { return img_bytes(F, L - 2); }
size_t tc_eval_smem(int F, int L) { return img_bytes_padded(F, L - 2) + (size_t)kTile * F * 2; }
size_t tc_fit_smem(int F, int L) {
  return (size_t)(2 + (L - 1)) * kTile * F * 2 + 2 * (size_t)kTile * 16 * 2 + img_bytes_padded(F, L - 2);
}

bool tc_supported(int f, int L, int in_dim, int out_dim) {
  const int F = tc_fpad(f);
  if (out_dim != 1 || (in_dim != 2 && in_dim != 3)) return false;
  if (L < 3 || F > 64) return false;
  if ((L - 2) * F + 2 * F + 32 > 512) return false;  // TMEM: Z, X, dW_l accumulators, dW0 / dWlast blocks
  if (tc_fit_smem(F, L) > 225 * 1024) return false;
  return true;
}

template <int F>
static cudaError_t launch_eval_f(const EvalArgs& a, int L_max, int n_blocks, cudaStream_t st) {
  const size_t smem = tc_eval_smem(F, L_max);
  cudaError_t e = cudaFuncSetAttribute(tc_eval_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  tc_eval_kernel<F><<<n_blocks, kTcThreads, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_tc_eval(const EvalArgs& a, int F_PAD, int L_max, int n_blocks, cudaStream_t st) {
  switch (F_PAD) {
    case 16: return launch_eval_f<16>(a, L_max, n_blocks, st);
    case 32: return launch_eval_f<32>(a, L_max, n_blocks, st);
    case 48: return launch_eval_f<48>(a, L_max, n_blocks, st);
    case 64: return launch_eval_f<64>(a, L_max, n_blocks, st);
    default: return cudaErrorInvalidValue;
  }
}

template <int F>
static cudaError_t launch_fit_f(const FitArgs& a, int L_max, int n_blocks, cudaStream_t st) {
  // at least 48 KB so that the 16 KB the M = 64 contractions read from a dz buffer always exist
  const size_t smem = tc_fit_smem(F, L_max) < 49152 ? 49152 : tc_fit_smem(F, L_max);
  cudaError_t e = cudaFuncSetAttribute(tc_fit_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  tc_fit_kernel<F><<<n_blocks, kTcThreads, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_tc_fit(const FitArgs& a, int F_PAD, int L_max, int n_blocks, cudaStream_t st) {
  switch (F_PAD) {
    case 16: return launch_fit_f<16>(a, L_max, n_blocks, st);
    case 32: return launch_fit_f<32>(a, L_max, n_blocks, st);
    case 48: return launch_fit_f<48>(a, L_max, n_blocks, st);
    case 64: return launch_fit_f<64>(a, L_max, n_blocks, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace brief
