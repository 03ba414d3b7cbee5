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
// thread mapping shared by both kernels
// ==================================================================================================================
// A CTA has 128 * (F/16) threads.  Warp w serves TMEM lane quadrant q = w & 3 (rows 32q .. 32q+31 of the tile, the
// only lanes tcgen05.ld lets it touch) and column group cg = w >> 2 (columns 16cg .. 16cg+15): every thread owns
// ONE 16-column chunk of its sample row per layer, so a hidden-layer epilogue is one tcgen05.ld.x16 + 16 sines +
// two 16-byte operand stores, and an SM holds 4 warps per scheduler for F = 64 instead of 1.
template <int F>
struct TcCfg {
  static constexpr int CW = F / 16;
  static constexpr int THREADS = 128 * CW;
  static constexpr int EVAL_MIN_BLOCKS = F >= 48 ? 2 : F == 32 ? 4 : 8;  // ~1024 threads per SM
  static constexpr int FIT_MIN_BLOCKS = F >= 48 ? 1 : 2;                  // must match fit_ctas_per_sm()
};

__device__ __forceinline__ void store_chunk16(unsigned char* buf, int r, int cg, const float* v) {
  *reinterpret_cast<uint4*>(buf + chunk_off(r, 2 * cg, kTile)) =
      make_uint4(pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]), pack_f16x2(v[4], v[5]), pack_f16x2(v[6], v[7]));
  *reinterpret_cast<uint4*>(buf + chunk_off(r, 2 * cg + 1, kTile)) =
      make_uint4(pack_f16x2(v[8], v[9]), pack_f16x2(v[10], v[11]), pack_f16x2(v[12], v[13]), pack_f16x2(v[14], v[15]));
}
__device__ __forceinline__ void store_chunk16_sat(unsigned char* buf, int r, int cg, const float* v) {
  *reinterpret_cast<uint4*>(buf + chunk_off(r, 2 * cg, kTile)) =
      make_uint4(pack_f16x2_sat(v[0], v[1]), pack_f16x2_sat(v[2], v[3]), pack_f16x2_sat(v[4], v[5]),
                 pack_f16x2_sat(v[6], v[7]));
  *reinterpret_cast<uint4*>(buf + chunk_off(r, 2 * cg + 1, kTile)) =
      make_uint4(pack_f16x2_sat(v[8], v[9]), pack_f16x2_sat(v[10], v[11]), pack_f16x2_sat(v[12], v[13]),
                 pack_f16x2_sat(v[14], v[15]));
}

// theta_i = w * z_i + (w * b)_i for the thread's 16 columns
__device__ __forceinline__ void theta16(const float* z, const float* __restrict__ wb, float w, float* th) {
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    const float4 b4 = *reinterpret_cast<const float4*>(wb + i);
    th[i] = fmaf(z[i], w, b4.x); th[i + 1] = fmaf(z[i + 1], w, b4.y);
    th[i + 2] = fmaf(z[i + 2], w, b4.z); th[i + 3] = fmaf(z[i + 3], w, b4.w);
  }
}

// ==================================================================================================================
// forward / decompress
// ==================================================================================================================
template <int F, bool DUMP>
__global__ void __launch_bounds__(TcCfg<F>::THREADS, TcCfg<F>::EVAL_MIN_BLOCKS) tc_eval_kernel(EvalArgs a) {
  constexpr int CW = TcCfg<F>::CW;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ NetDev sn;
  __shared__ __align__(8) uint64_t bar_w, bar_mma;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float4 s_row[kTile];
  __shared__ float s_y[CW][kTile];
  __shared__ __align__(16) unsigned short s_out[kTile];

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int q = warp & 3, cg = warp >> 2, r = 32 * q + lane;
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
  if (warp == 0) tmem_alloc(&tmem_base_s, TCOLS);
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
  const uint32_t tm = tmem_base_s;
  const uint32_t my_tmem = tm + ((uint32_t)(32 * q) << 16) + 16 * cg;
  const uint32_t aAct = smem_u32(sAct), aW = smem_u32(sW);
  const long long total = a.coords ? a.n_coords : n.n_vox;
  const long long n_tiles = (total + kTile - 1) / kTile;
  const long long tile_begin = chunk * kEvalTilesPerBlock;
  const long long tile_end = min(n_tiles, tile_begin + kEvalTilesPerBlock);
  uint32_t phase = 0;
  const float wh = n.wh;

  auto load_row = [&](long long tile) {  // column group 0 fetches the tile's coordinates for everyone
    const long long s = tile * kTile + r;
    float x0 = 0.f, x1 = 0.f, x2 = 0.f;
    if (tile < tile_end && s < total) {
      if (a.coords) {
        x0 = a.coords[s * n.in_dim];
        x1 = a.coords[s * n.in_dim + 1];
        x2 = n.in_dim == 3 ? a.coords[s * n.in_dim + 2] : 0.f;
      } else {
        brief_coords(n, a.axes, s, x0, x1, x2);
      }
    }
    s_row[r] = make_float4(x0, x1, x2, 0.f);
  };
  if (cg == 0) load_row(tile_begin);
  mbar_wait(&bar_w, 0);
  __syncthreads();

  for (long long tile = tile_begin; tile < tile_end; ++tile) {
    const long long s = tile * kTile + r;
    const bool valid = s < total;
    const float4 xr = s_row[r];
    // ---- layer 0 on CUDA cores -> fp16 operand rows (this thread: features 16cg .. 16cg+15)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float z[8];
      const uint4 pk = first_layer8<DUMP>(s_w0b, 16 * cg + 8 * h, xr.x, xr.y, xr.z, n.w0, z);
      if (DUMP && valid)
        for (int i = 0; i < 8; ++i)
          if (16 * cg + 8 * h + i < n.f) a.layers_out[s * n.f + 16 * cg + 8 * h + i] = z[i];
      *reinterpret_cast<uint4*>(sAct + chunk_off(r, 2 * cg + h, kTile)) = pk;
    }
    float ypart = 0.f;
    // ---- hidden layers on the tensor core
    for (int l = 1; l <= NH; ++l) {
      tc_fence_before();
      fence_async_smem();
      __syncthreads();
      if (warp == 0 && elect_one()) {
        tc_fence_after();
        issue_forward<F>(tm, aAct, aW + (uint32_t)(l - 1) * F * F * 2);
        commit(&bar_mma);
      }
      mbar_wait(&bar_mma, phase);
      phase ^= 1;
      tc_fence_after();
      float v[16], th[16];
      tmem_ld16(my_tmem, v);
      tmem_ld_wait();
      theta16(v, s_wb + (l - 1) * F + 16 * cg, wh, th);
      if (DUMP && valid) {
        float* zdump = a.layers_out + (long long)l * total * n.f + s * n.f;
        for (int i = 0; i < 16; ++i)
          if (16 * cg + i < n.f) zdump[16 * cg + i] = th[i] / wh;
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fast_sin(th[i]);
      if (l < NH) {
        store_chunk16(sAct, r, cg, v);
      } else {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(s_wl + 16 * cg + i);
          ypart = fmaf(w4.x, v[i], ypart); ypart = fmaf(w4.y, v[i + 1], ypart);
          ypart = fmaf(w4.z, v[i + 2], ypart); ypart = fmaf(w4.w, v[i + 3], ypart);
        }
      }
    }
    // ---- last layer: fixed-order sum of the column groups' partial dot products, then the output epilogue
    s_y[cg][r] = ypart;
    __syncthreads();
    const bool full = tile * kTile + kTile <= total;
    if (cg == 0) {
      float y = s_bl[0];
#pragma unroll
      for (int c = 0; c < CW; ++c) y += s_y[c][r];
      if (a.out_f32) {
        if (valid) a.out_f32[s] = y;
      } else {
        void* dst = a.out_ptrs[net_id];
        if (a.out_dtype == 2) {
          if (valid) reinterpret_cast<float*>(dst)[s] = y;
        } else {  // inverse normalisation + truncating cast
          const float vden = brief_denorm(n, y);
          if (a.out_dtype == 1) {
            if (full) s_out[r] = (unsigned short)(int)vden;
            else if (valid) reinterpret_cast<unsigned short*>(dst)[s] = (unsigned short)(int)vden;
          } else {
            if (full) reinterpret_cast<unsigned char*>(s_out)[r] = (unsigned char)(int)vden;
            else if (valid) reinterpret_cast<unsigned char*>(dst)[s] = (unsigned char)(int)vden;
          }
        }
      }
      load_row(tile + 1);
    }
    __syncthreads();
    // staged tile -> 16-byte vector stores (256 B contiguous for uint16)
    if (!a.out_f32 && a.out_dtype != 2 && full) {
      void* dst = a.out_ptrs[net_id];
      if (a.out_dtype == 1) {
        if (t < 16)
          reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(dst) + tile * kTile)[t] =
              reinterpret_cast<const uint4*>(s_out)[t];
      } else if (t < 8) {
        reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(dst) + tile * kTile)[t] =
            reinterpret_cast<const uint4*>(s_out)[t];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, TCOLS);
}

// ==================================================================================================================
// fit: gather + forward + weighted L2 + backward -> per-slice gradient partials
// ==================================================================================================================
// shared memory (dynamic):  sDz[2] | sAct[NS] | sX | sDY | image        (NS = L-1 sine layers)
//   sDz first: the M = 64 dW contractions read 8 feature groups (16 KB) from the start of a dz buffer whatever F is.
// TMEM columns: Z [0,F) | X [F,2F) | dW_l [2F + (l-1)F, +F) l=1..NH | dW0 [.., +16) | dWlast [.., +16)
__host__ __device__ constexpr int fit_tmem_cols(int F, int NH) { return tmem_cols_pow2((NH + 2) * F + 32); }

#ifdef BRIEF_TC_TIMING
__device__ unsigned long long g_tc_timing[16];
#define TT(var) const long long var = clock64()
#define TACC(slot, expr) if (blockIdx.x == 0 && (t == 0 || t == TcCfg<F>::THREADS - 1)) atomicAdd(&g_tc_timing[(slot) + (t == 0 ? 0 : 8)], (unsigned long long)(expr))
#else
#define TT(var)
#define TACC(slot, expr)
#endif

template <int F>
__global__ void __launch_bounds__(TcCfg<F>::THREADS, TcCfg<F>::FIT_MIN_BLOCKS) tc_fit_kernel(FitArgs a) {
  constexpr int CW = TcCfg<F>::CW;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ NetDev sn;
  __shared__ __align__(8) uint64_t bar_w, bar_mma;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float4 s_row[kTile];
  __shared__ float s_y[CW][kTile];
  __shared__ float s_dy[kTile];
  __shared__ float s_red[4];

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int q = warp & 3, cg = warp >> 2, r = 32 * q + lane;
  const int wi = tc_find_work(a.work_prefix, a.n_work, blockIdx.x);
  const int net_id = a.work_net[wi];
  const int slice = blockIdx.x - a.work_prefix[wi];
  tc_load_net(sn, a.nets[net_id]);
  if (t == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const NetDev& n = sn;
  const int NH = n.L - 2, NS = n.L - 1, f = n.f;
  const int tcols = fit_tmem_cols(F, NH);
  if (warp == 0) tmem_alloc(&tmem_base_s, tcols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
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
  if (cg == 0) {  // second column group of the two 16-column blocks is constant zero
    *reinterpret_cast<uint4*>(sX + chunk_off(r, 1, kTile)) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(sDY + chunk_off(r, 1, kTile)) = make_uint4(0, 0, 0, 0);
  }

  const uint32_t tm = tmem_base_s;
  const uint32_t my_tmem = tm + ((uint32_t)(32 * q) << 16) + 16 * cg;
  const uint32_t TZ = tm, TX = tm + F, TDW = tm + 2 * F, TDW0 = TDW + NH * F, TDWL = TDW0 + 16;
  const uint32_t aDz = smem_u32(sDz), aAct = smem_u32(sAct), aX = smem_u32(sX), aDY = smem_u32(sDY), aW = smem_u32(sW);
  uint32_t phase = 0;
  const float wh = n.wh, w0 = n.w0;
  const long long s_begin = (long long)slice * n.slice_len;
  const long long s_end = min((long long)n.batch, s_begin + n.slice_len);
  float loss_acc = 0.f;
  int it = 0;

  // the sampler gather for one tile (column group 0): index -> coordinates, normalised target, weight
  float gx0 = 0.f, gx1 = 0.f, gx2 = 0.f, gy = 0.f, gw = 0.f;
  auto gather = [&](long long tile0) {
    const long long s = tile0 + r;
    gx0 = gx1 = gx2 = gy = gw = 0.f;
    if (tile0 < s_end && s < s_end) {
      long long idx;
      if (n.mode == 0) idx = s;
      else if (a.idx) idx = a.idx[n.idx_off + s];
      else idx = brief_sample_index(a.seed, a.step, (uint32_t)net_id, (uint64_t)s, (uint64_t)n.n_vox);
      brief_coords(n, a.axes, idx, gx0, gx1, gx2);
      const float raw = brief_raw_value(n, idx);
      gy = brief_normalize(n, raw);
      gw = brief_weight(n, idx, raw);
    }
  };
  if (cg == 0) gather(s_begin);
  mbar_wait(&bar_w, 0);

  for (long long tile0 = s_begin; tile0 < s_end; tile0 += kTile, ++it) {
    TT(ctile);
    const bool valid = tile0 + r < s_end;
    const float yv = gy, wv = gw;
    if (cg == 0) {
      s_row[r] = make_float4(gx0, gx1, gx2, 0.f);
      // B operand of the dW0 contraction: [x_hi(3), 1, x_lo(3), 0]  (hi/lo split keeps fp32-grade coordinates)
      const float h0 = __half2float(__float2half_rn(gx0)), h1 = __half2float(__float2half_rn(gx1)),
                  h2 = __half2float(__float2half_rn(gx2));
      *reinterpret_cast<uint4*>(sX + chunk_off(r, 0, kTile)) =
          make_uint4(pack_f16x2(h0, h1), pack_f16x2(h2, 1.0f), pack_f16x2(gx0 - h0, gx1 - h1), pack_f16x2(gx2 - h2, 0.f));
    }
    __syncthreads();
    const float4 xr = s_row[r];
    // ---- forward, layer 0 (CUDA cores): features 16cg .. 16cg+15
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float z[8];
      *reinterpret_cast<uint4*>(sAct + chunk_off(r, 2 * cg + h, kTile)) =
          first_layer8<false>(s_w0b, 16 * cg + 8 * h, xr.x, xr.y, xr.z, w0, z);
    }
    // ---- forward, hidden layers (tensor core); a_l -> sAct[l]
    float th[16];
    float ypart = 0.f;
    for (int l = 1; l <= NH; ++l) {
      TT(c0);
      tc_fence_before();
      fence_async_smem();
      TT(c1);
      __syncthreads();
      TT(c2);
      if (warp == 0 && elect_one()) {
        tc_fence_after();
        issue_forward<F>(TZ, aAct + (uint32_t)(l - 1) * BUF, aW + (uint32_t)(l - 1) * F * F * 2);
        commit(&bar_mma);
      }
      TT(c3);
      mbar_wait(&bar_mma, phase);
      TT(c4);
      phase ^= 1;
      tc_fence_after();
      float v[16];
      tmem_ld16(my_tmem, v);
      tmem_ld_wait();
      TT(c5);
      TACC(0, c1 - c0); TACC(1, c2 - c1); TACC(2, c3 - c2); TACC(3, c4 - c3); TACC(4, c5 - c4);
      theta16(v, s_wb + (l - 1) * F + 16 * cg, wh, th);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fast_sin(th[i]);
      store_chunk16(sAct + (size_t)l * BUF, r, cg, v);
      { TT(c6); TACC(5, c6 - c5); }
      if (l == NH) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(s_wl + 16 * cg + i);
          ypart = fmaf(w4.x, v[i], ypart); ypart = fmaf(w4.y, v[i + 1], ypart);
          ypart = fmaf(w4.z, v[i + 2], ypart); ypart = fmaf(w4.w, v[i + 3], ypart);
        }
      }
    }
    // ---- loss (datal2, main.py:176-182) and the scaled output gradient
    s_y[cg][r] = ypart;
    __syncthreads();
    if (cg == 0) {
      float y = s_bl[0];
#pragma unroll
      for (int c = 0; c < CW; ++c) y += s_y[c][r];
      float dys = 0.f;
      if (valid) {
        const float e = y - yv;
        const float wt = (n.tau != 0.f && y <= n.tau) ? 1.0f : wv;
        loss_acc = fmaf(wt * e, e, loss_acc);
        dys = kGradScale * wt * e;
      }
      s_dy[r] = dys;
      *reinterpret_cast<uint4*>(sDY + chunk_off(r, 0, kTile)) = make_uint4(pack_f16x2_sat(dys, 0.f), 0, 0, 0);
      gather(tile0 + kTile);  // prefetch the next tile's samples; consumed at the top of the next iteration
    }
    __syncthreads();
    {  // dz_NH = dy * Wlast * w * cos(w z_NH) from the sine arguments still in registers
      const float dys = s_dy[r];
      float dz[16];
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(s_wl + 16 * cg + i);
        dz[i] = dys * w4.x * wh * fast_cos(th[i]); dz[i + 1] = dys * w4.y * wh * fast_cos(th[i + 1]);
        dz[i + 2] = dys * w4.z * wh * fast_cos(th[i + 2]); dz[i + 3] = dys * w4.w * wh * fast_cos(th[i + 3]);
      }
      store_chunk16_sat(sDz, r, cg, dz);
    }
    // ---- backward through the hidden layers
    int cur = 0;
    for (int l = NH; l >= 1; --l) {
      tc_fence_before();
      fence_async_smem();
      __syncthreads();
      if (warp == 0 && elect_one()) {
        tc_fence_after();
        const uint32_t dzb = aDz + (uint32_t)cur * BUF;
        // what the epilogue waits for: z_{l-1} (recomputed) and dX_{l-1}
        if (l >= 2) issue_forward<F>(TZ, aAct + (uint32_t)(l - 2) * BUF, aW + (uint32_t)(l - 2) * F * F * 2);
        issue_dx<F>(TX, dzb, aW + (uint32_t)(l - 1) * F * F * 2);
        commit(&bar_mma);
        // off the critical path (tracked by the next commit): dW_l, db_l (+ dWlast, dblast)
        issue_dw<F>(TDW + (uint32_t)(l - 1) * F, dzb, aAct + (uint32_t)(l - 1) * BUF, it > 0);
        if (l == NH) issue_dw<16>(TDWL, aAct + (uint32_t)NH * BUF, aDY, it > 0);
      }
      mbar_wait(&bar_mma, phase);
      phase ^= 1;
      tc_fence_after();
      float vx[16], dz[16];
      if (l >= 2) {
        float vz[16];
        tmem_ld16(my_tmem, vz);
        tmem_ld16(my_tmem + F, vx);
        tmem_ld_wait();
        theta16(vz, s_wb + (l - 2) * F + 16 * cg, wh, dz);
#pragma unroll
        for (int i = 0; i < 16; ++i) dz[i] = vx[i] * wh * fast_cos(dz[i]);
      } else {  // l == 1: dz_0 = dX_0 * w0 * cos(w0 z_0), z_0 recomputed on CUDA cores
        tmem_ld16(my_tmem + F, vx);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 w = s_w0b[16 * cg + i];
          float z = w.w;
          z = fmaf(w.x, xr.x, z); z = fmaf(w.y, xr.y, z); z = fmaf(w.z, xr.z, z);
          dz[i] = vx[i] * w0 * fast_cos(w0 * z);
        }
      }
      store_chunk16_sat(sDz + (size_t)(cur ^ 1) * BUF, r, cg, dz);
      cur ^= 1;
    }
    // ---- dW0 += dz_0^T [x_hi, 1, x_lo]; wait so that the next tile may overwrite sX / sDz / sAct
    tc_fence_before();
    fence_async_smem();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      issue_dw<16>(TDW0, aDz + (uint32_t)cur * BUF, aX, it > 0);
      commit(&bar_mma);
    }
    mbar_wait(&bar_mma, phase);
    phase ^= 1;
    tc_fence_after();
    { TT(cend); TACC(6, cend - ctile); TACC(7, 1); }
  }

  // ---- slice epilogue: loss partial + gradient partials (TMEM -> global), scale removed in fp32
  const float inv_count = 1.0f / ((float)n.batch * (float)n.out_dim);
  if (cg == 0) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) loss_acc += __shfl_down_sync(0xffffffffu, loss_acc, off);
    if (lane == 0) s_red[q] = loss_acc;
  }
  float* part = a.partials + n.part_off + (long long)slice * n.P_dev;
  for (int i = t; i < n.P_dev; i += TcCfg<F>::THREADS) part[i] = 0.f;
  __syncthreads();
  if (t == 0) a.loss_partials[n.slice_off + slice] = (((s_red[0] + s_red[1]) + s_red[2]) + s_red[3]) * inv_count;
  if (it > 0) {
    const float unscale = 2.0f * inv_count / kGradScale;
    const int o = q * 16 + lane;  // accumulator row held by this thread (M = 64 layout: lanes 0..15 of each quadrant)
    const bool row_ok = lane < 16 && o < F;
    const int F4 = n.F4;
    float v[16];
    for (int l = 1; l <= NH; ++l) {
      tmem_ld16(my_tmem + 2 * F + (l - 1) * F, v);
      tmem_ld_wait();
      if (row_ok && o < f) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int k = 16 * cg + i;
          if (k < f) part[dl_W(n, l) + o * F4 + k] = v[i] * unscale;
          else if (k == f) part[dl_b(n, l) + o] = v[i] * unscale;
        }
      }
    }
    if (cg == 0) {
      tmem_ld16(my_tmem + 2 * F + NH * F, v);  // dW0 block
      tmem_ld_wait();
      if (row_ok && o < f) {
        part[dl_W0(n) + 4 * o + 0] = (v[0] + v[4]) * unscale;
        part[dl_W0(n) + 4 * o + 1] = (v[1] + v[5]) * unscale;
        if (n.in_dim == 3) part[dl_W0(n) + 4 * o + 2] = (v[2] + v[6]) * unscale;
        part[dl_b0(n) + o] = v[3] * unscale;
      }
      tmem_ld16(my_tmem + 2 * F + NH * F + 16, v);  // dWlast block: row = feature of a_NH, column 0
      tmem_ld_wait();
      if (row_ok) {
        if (o < f) part[dl_Wlast(n) + o] = v[0] * unscale;
        else if (o == f) part[dl_blast(n)] = v[0] * unscale;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, tcols);
}

// ==================================================================================================================
// host side
// ==================================================================================================================
int tc_fpad(int f) { return ((f + 1 + 15) / 16) * 16; }
int tc_fit_ctas_per_sm(int F) { return F >= 48 ? 1 : 2; }  // == TcCfg<F>::FIT_MIN_BLOCKS
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
  cudaError_t e;
  if (a.layers_out) {
    e = cudaFuncSetAttribute(tc_eval_kernel<F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    tc_eval_kernel<F, true><<<n_blocks, TcCfg<F>::THREADS, smem, st>>>(a);
  } else {
    e = cudaFuncSetAttribute(tc_eval_kernel<F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    tc_eval_kernel<F, false><<<n_blocks, TcCfg<F>::THREADS, smem, st>>>(a);
  }
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
  tc_fit_kernel<F><<<n_blocks, TcCfg<F>::THREADS, smem, st>>>(a);
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

#ifdef BRIEF_TC_TIMING
extern "C" int brief_debug_read_timing(unsigned long long* out, int reset) {
  cudaMemcpyFromSymbol(out, brief::g_tc_timing, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(brief::g_tc_timing, z, sizeof z); }
  return 0;
}
#endif
