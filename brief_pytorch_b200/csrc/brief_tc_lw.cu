// brief_tc_lw.cu — the tensor-core path for WIDE networks (128 < F_PAD <= 256: neuron.yaml as shipped, f = 228; the upper
// widths of hipct.yaml's by_var allocation), layer by layer.
//
// Reference work replaced (file:line relative to the reference root): SIREN.forward (utils/Networks.py:269-271), the
// samplers' gather (main.py:126-163), datal2 (main.py:176-182) and the autograd backward (main.py:396) for one training
// step, and reconstruct_flattened (utils/misc.py:59-92) for the decode.
//
// At these widths ONE layer's weights are 41 .. 128 KB of fp16 and one 128-sample tile of activations is 36 .. 64 KB per
// layer, so neither the all-layers-resident kernel of brief_tc.cu nor the two-buffer kernel of brief_tc_wide.cu fits in
// shared memory, and one layer's dW accumulator (F x F fp32) fills tensor memory by itself.  These widths are ordinary
// GEMMs, and they are run as such: one launch per layer and direction over ALL tiles of a pass, activations travelling
// between launches through global memory as fp16 tiles that are ALREADY in the tcgen05 operand layout (brief_umma.cuh),
// so that every operand is fetched by plain bulk (TMA) copies — a K chunk of 64 columns of a [128 x F] tile is 16 KB of
// contiguous bytes, and the same tile is a valid K-major operand (forward, dX) and MN-major operand (dW):
//
//   sample_l0   index -> coordinates / target / weight; layer 0 on CUDA cores (K = 3)      -> ACT_0, COS_0, X block
//   gemm FWD    theta_j = ACT_{j-1} W'_j^T (N = F, weights resident in shared memory), sin / cos -> ACT_j, COS_j
//   loss        y = a_NH . Wlast + b (fp32), datal2, dy', dz_NH = dy' w Wlast cos                -> DZ, DY block, losses
//   dw          dW_l = DZ_l^T ACT_{l-1} summed over a slice of tiles in TENSOR MEMORY (M = 128 output rows per CTA,
//               N = F), one drain per slice into the per-slice partial slot the optimiser kernel reduces in order;
//               the same kernel with a 16-column B block gives dW0 / db0 (B = X block) and dWlast / dblast (B = DY block)
//   gemm BWD    dX = DZ_l W'_l (B read MN-major), dz_{l-1} = dX * COS_{l-1}                       -> DZ (ping-pong)
//   gemm EVAL   forward without the cosine (decode), last layer + inverse normalisation by a CUDA-core kernel
//
// Same numerics as the fused kernels: fp16 operands, fp32 accumulation in TMEM, omega and bias inside W' (two constant-one
// columns), dz carried with the static kGradScale, layer 0 and the last layer in fp32.  The cosine of every layer is kept
// (fp16) instead of recomputed: the forward GEMM has theta in registers anyway.
// Bound: HBM — a step moves ~4.5 KB per sample and hidden layer at F = 256 against 0.79 MFLOP.
#include "brief_tc_common.cuh"

namespace brief {

using namespace umma;

constexpr int kLwKC = 64;                      // columns of the A operand per pipeline stage (16 KB)
constexpr int kLwStages = 4;
constexpr uint32_t kLwStageBytes = kTile * kLwKC * 2;
constexpr int kLwEpiWarps = 8;                 // two per lane quadrant (column halves)
constexpr int kLwThreads = (kLwEpiWarps + 2) * 32;  // + producer warp + MMA-issue warp

__host__ __device__ inline size_t lw_tile_bytes(int F) { return (size_t)kTile * F * 2; }

// scratch of one pass of T tiles: ACT[NH+1] | COS[NH+1] | DZ[2] | XB | DYB | per-sample arrays
__host__ __device__ inline size_t lw_off_act(int F, long long T, int j) { return (size_t)j * T * lw_tile_bytes(F); }
__host__ __device__ inline size_t lw_off_cos(int F, long long T, int NH, int j) { return (size_t)(NH + 1 + j) * T * lw_tile_bytes(F); }
__host__ __device__ inline size_t lw_off_dz(int F, long long T, int NH, int b) { return (size_t)(2 * NH + 2 + b) * T * lw_tile_bytes(F); }
__host__ __device__ inline size_t lw_off_xb(int F, long long T, int NH) { return (size_t)(2 * NH + 4) * T * lw_tile_bytes(F); }
__host__ __device__ inline size_t lw_off_dyb(int F, long long T, int NH) { return lw_off_xb(F, T, NH) + (size_t)T * kTile * 16 * 2; }
__host__ __device__ inline size_t lw_off_xs(int F, long long T, int NH) { return lw_off_dyb(F, T, NH) + (size_t)T * kTile * 16 * 2; }
__host__ __device__ inline size_t lw_off_ws(int F, long long T, int NH) { return lw_off_xs(F, T, NH) + (size_t)T * kTile * 16; }
__host__ __device__ inline size_t lw_off_ls(int F, long long T, int NH) { return lw_off_ws(F, T, NH) + (size_t)T * kTile * 4; }
__host__ __device__ inline size_t lw_off_tl(int F, long long T, int NH) { return lw_off_ls(F, T, NH) + (size_t)T * kTile * 4; }
size_t lw_scratch_bytes(int F, int L, long long T) { return lw_off_tl(F, T, L - 2) + (size_t)T * 4 + 256; }
// decode: two activation buffers
size_t lw_eval_scratch_bytes(int F, long long T) { return 2 * (size_t)T * lw_tile_bytes(F) + 256; }

size_t lw_tile_bytes_host(int F) { return lw_tile_bytes(F); }
size_t lw_fit_offset(int F, int L, long long T, int what, int j) {
  const int NH = L - 2;
  switch (what) {
    case 0: return lw_off_act(F, T, j);
    case 1: return lw_off_cos(F, T, NH, j);
    case 2: return lw_off_dz(F, T, NH, j);
    case 3: return lw_off_xb(F, T, NH);
    default: return lw_off_dyb(F, T, NH);
  }
}

// ---- sampler + layer 0 -------------------------------------------------------------------------------------------------
// One CTA per tile.  Fit: index -> voxel -> (coords, normalised target, weight), kept for the loss; decode / explicit
// coordinates: coords only.  Layer 0 (K = 3) runs on CUDA cores in fp32; the X block [x_hi 1 x_lo ...] is the B operand
// of the dW0 contraction (same split as the fused kernels).
__global__ void __launch_bounds__(256) lw_sample_l0_kernel(LwArgs a) {
  __shared__ float4 s_x[kTile];
  const int wi = tc_find_work(a.work_prefix, a.n_work, blockIdx.x);
  const NetDev& n = a.nets[a.work_net[wi]];
  const long long tile = (long long)(blockIdx.x - a.work_prefix[wi]) + a.tile_first[wi];  // tile of the network
  const long long st = a.tile_base[wi] + (blockIdx.x - a.work_prefix[wi]);                 // tile of the pass (scratch)
  const int F = a.F, NH = n.L - 2, f = n.f, t = threadIdx.x;
  const long long total = a.eval ? (a.coords ? a.n_coords : n.n_vox) : (long long)n.batch;
  if (t < kTile) {
    const long long s = tile * kTile + t;
    const bool ok = s < total;
    float x0 = 0.f, x1 = 0.f, x2 = 0.f, yv = 0.f, wv = 0.f;
    if (ok) {
      if (a.eval) {
        if (a.coords) {
          x0 = a.coords[s * n.in_dim];
          x1 = a.coords[s * n.in_dim + 1];
          x2 = n.in_dim == 3 ? a.coords[s * n.in_dim + 2] : 0.f;
        } else {
          brief_coords(n, a.axes, s, x0, x1, x2);
        }
      } else {
        long long v;
        if (n.mode == 0) v = s;
        else if (a.idx) v = a.idx[n.idx_off + s];
        else v = brief_sample_index(a.seed, a.state ? a.state->step : a.step, n.stream_id, (uint64_t)s, (uint64_t)n.n_vox);
        const float raw = brief_raw_value(n, v);
        brief_coords(n, a.axes, v, x0, x1, x2);
        yv = brief_normalize(n, raw);
        wv = brief_weight(n, v, raw);
      }
    }
    s_x[t] = make_float4(x0, x1, x2, yv);
    if (!a.eval) {
      unsigned char* S = a.scratch;
      reinterpret_cast<float4*>(S + lw_off_xs(F, a.T, NH))[st * kTile + t] = make_float4(x0, x1, x2, yv);
      reinterpret_cast<float*>(S + lw_off_ws(F, a.T, NH))[st * kTile + t] = wv;
      const float h0 = __half2float(__float2half_rn(x0)), h1 = __half2float(__float2half_rn(x1)),
                  h2 = __half2float(__float2half_rn(x2));
      const uint32_t p01 = pack_f16x2(h0, h1), p21 = pack_f16x2(h2, 1.0f);
      unsigned char* xb = S + lw_off_xb(F, a.T, NH) + (size_t)st * kTile * 16 * 2;
      *reinterpret_cast<uint4*>(xb + chunk_off(t, 0, kTile)) =
          make_uint4(p01, p21, pack_f16x2(x0 - h0, x1 - h1), pack_f16x2(x2 - h2, 1.0f));
      *reinterpret_cast<uint4*>(xb + chunk_off(t, 1, kTile)) = make_uint4(p01, pack_f16x2(h2, 0.f), 0u, 0u);
    }
  }
  __syncthreads();
  const float4* w0b = reinterpret_cast<const float4*>(a.wpack + n.wpack_off + img_side_off(F, NH));
  unsigned char* act = a.scratch + (a.eval ? 0 : lw_off_act(F, a.T, 0)) + (size_t)st * lw_tile_bytes(F);
  unsigned char* cs = a.scratch + lw_off_cos(F, a.T, NH, 0) + (size_t)st * lw_tile_bytes(F);
  float* zdump = a.layers_out;
  for (int i = t; i < kTile * (F / 8); i += 256) {
    const int r = i & (kTile - 1), g = i >> 7;  // consecutive threads = consecutive rows: 512 contiguous bytes per warp
    const float4 x = s_x[r];
    float sv[8], cv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = 8 * g + e;
      if (col < f) {
        const float4 w = __ldg(w0b + col);
        float z = w.w;
        z = fmaf(w.x, x.x, z); z = fmaf(w.y, x.y, z); z = fmaf(w.z, x.z, z);
        if (zdump && tile * kTile + r < total) zdump[(tile * kTile + r) * f + col] = z;
        sv[e] = fast_sin(n.w0 * z);
        cv[e] = fast_cos(n.w0 * z);
      } else {
        sv[e] = col < f + 2 ? 1.0f : 0.0f;
        cv[e] = 0.0f;
      }
    }
    *reinterpret_cast<uint4*>(act + chunk_off(r, g, kTile)) =
        make_uint4(pack_f16x2(sv[0], sv[1]), pack_f16x2(sv[2], sv[3]), pack_f16x2(sv[4], sv[5]), pack_f16x2(sv[6], sv[7]));
    if (!a.eval)
      *reinterpret_cast<uint4*>(cs + chunk_off(r, g, kTile)) =
          make_uint4(pack_f16x2(cv[0], cv[1]), pack_f16x2(cv[2], cv[3]), pack_f16x2(cv[4], cv[5]), pack_f16x2(cv[6], cv[7]));
  }
}

// ---- the layer GEMM: D[128 x F] = A[128 x F] * B, B = one layer's weights resident in shared memory ------------------------
// MODE 0 (FWD):  B K-major (W'_j),  epilogue sin / cos -> ACT_out, COS_out
// MODE 1 (EVAL): B K-major,         epilogue sin -> ACT_out (optionally the pre-activations for brief_forward's layer dump)
// MODE 2 (BWD):  B MN-major (W'_l), epilogue dz = dX * scale * COS_in -> DZ_out
// Roles: producer warp (bulk copies of 64-column K chunks of the A tile into a 4-stage ring), MMA-issue warp (N = F
// contractions into one of two TMEM accumulators), 8 epilogue warps (tile t's epilogue under tile t+1's contractions).
// A CTA serves tiles c, c + C, c + 2C, ... of ONE network (C = the CTAs the host gave that network), so the weights are
// staged once per CTA.
template <int MODE>
__global__ void __launch_bounds__(kLwThreads, 1) lw_gemm_kernel(LwArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar_w, bar_full[kLwStages], bar_empty[kLwStages], bar_accf[2], bar_acce[2];
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int wi = tc_find_work(a.work_prefix, a.n_work, blockIdx.x);
  const NetDev& n = a.nets[a.work_net[wi]];
  const int c0 = blockIdx.x - a.work_prefix[wi], C = a.work_prefix[wi + 1] - a.work_prefix[wi];
  const int F = a.F, NH = n.L - 2, f = n.f;
  const int n_tiles = a.tile_count[wi];
  const long long tbase = a.tile_base[wi];
  const int n_kc = (F + kLwKC - 1) / kLwKC;
  if (t == 0) {
    mbar_init(&bar_w, 1);
    for (int i = 0; i < kLwStages; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_accf[i], 1); mbar_init(&bar_acce[i], kLwEpiWarps); }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  unsigned char* sW = smem;
  unsigned char* sA = smem + (((size_t)F * F * 2 + 127) & ~(size_t)127);
  const unsigned char* S = a.scratch;
  const unsigned char* src = S + a.in_off;
  const size_t TB = lw_tile_bytes(F);

  if (warp == kLwEpiWarps) {
    // ================================================= producer =====================================================
    if (c0 < n_tiles && elect_one()) {
      mbar_expect_tx(&bar_w, (uint32_t)F * F * 2);
      bulk_g2s(sW, a.wpack + n.wpack_off + (size_t)a.layer * F * F * 2, (uint32_t)F * F * 2, &bar_w);
      int it = 0;
      for (int tile = c0; tile < n_tiles; tile += C)
        for (int kc = 0; kc < n_kc; ++kc, ++it) {
          const int s = it % kLwStages;
          if (it >= kLwStages) mbar_wait(&bar_empty[s], (uint32_t)((it / kLwStages) - 1) & 1);
          const int cols = min(kLwKC, F - kc * kLwKC);
          const uint32_t bytes = (uint32_t)(cols / 8) * 2048u;
          mbar_expect_tx(&bar_full[s], bytes);
          bulk_g2s(sA + (size_t)s * kLwStageBytes, src + (size_t)(tbase + tile) * TB + (size_t)kc * kLwStageBytes, bytes, &bar_full[s]);
        }
    }
    __syncwarp();
  } else if (warp == kLwEpiWarps + 1) {
    // ================================================= MMA issue ====================================================
    const uint32_t idesc = make_idesc(128, F, false, MODE == 2);
    const uint32_t aW = smem_u32(sW), aA = smem_u32(sA);
    if (c0 < n_tiles) mbar_wait(&bar_w, 0);
    int it = 0, tl = 0;
    for (int tile = c0; tile < n_tiles; tile += C, ++tl) {
      const int acc = tl & 1;
      if (tl >= 2) mbar_wait(&bar_acce[acc], (uint32_t)((tl >> 1) - 1) & 1);
      tc_fence_after();
      for (int kc = 0; kc < n_kc; ++kc, ++it) {
        const int s = it % kLwStages;
        mbar_wait(&bar_full[s], (uint32_t)(it / kLwStages) & 1);
        tc_fence_after();
        if (elect_one()) {
          const int cols = min(kLwKC, F - kc * kLwKC);
          const uint32_t ab = aA + (uint32_t)s * kLwStageBytes;
          for (int kk = 0; kk < cols / 16; ++kk) {
            const int kg = kc * (kLwKC / 16) + kk;  // 16-column K step of the layer
            const uint64_t bd = MODE == 2 ? make_desc(aW + kg * 2 * 128, 128, (F / 8) * 128)
                                          : make_desc(aW + kg * 2 * (F / 8) * 128, (F / 8) * 128, 128);
            mma_f16(tm + (uint32_t)acc * 256, make_desc(ab + kk * 2 * kActLBO, kActLBO, 128), bd, idesc, (kc > 0 || kk > 0) ? 1u : 0u);
          }
          commit(&bar_empty[s]);
          if (kc == n_kc - 1) commit(&bar_accf[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // ================================================= epilogue =====================================================
    const int q = warp & 3, h = warp >> 2, r = 32 * q + lane;
    const int NC = F / 16, c_lo = h ? (NC + 1) / 2 : 0, c_hi = h ? NC : (NC + 1) / 2;
    const uint32_t lane_base = (uint32_t)(32 * q) << 16;
    unsigned char* out0 = a.scratch + a.out_off;
    unsigned char* out1 = a.scratch + a.out2_off;        // FWD: cosine tiles
    const unsigned char* cin = a.scratch + a.in2_off;    // BWD: cosine tiles of the layer below
    const float scale = a.layer == 0 ? n.w0 / n.wh : 1.0f;  // dX carries w_hidden (omega-scaled weights); layer 0 wants w_0
    const float inv_w = 1.0f / n.wh;
    int tl = 0;
    for (int tile = c0; tile < n_tiles; tile += C, ++tl) {
      const int acc = tl & 1;
      const size_t toff = (size_t)(tbase + tile) * TB;
      // BWD: this row's cosines of the layer below (up to 8 chunks of 32 bytes per thread) are fetched BEFORE the wait for
      // the accumulator, all loads in flight together: their DRAM latency sits under the tile's contractions
      uint4 kc[MODE == 2 ? 16 : 1];
      if (MODE == 2) {
#pragma unroll
        for (int ci = 0; ci < 8; ++ci)
          if (c_lo + ci < c_hi) {
            kc[2 * ci] = __ldcs(reinterpret_cast<const uint4*>(cin + toff + chunk_off(r, 2 * (c_lo + ci), kTile)));
            kc[2 * ci + 1] = __ldcs(reinterpret_cast<const uint4*>(cin + toff + chunk_off(r, 2 * (c_lo + ci) + 1, kTile)));
          }
      }
      mbar_wait(&bar_accf[acc], (uint32_t)(tl >> 1) & 1);
      tc_fence_after();
      auto chunk = [&](int c, uint4 k0, uint4 k1) {
        float v[16];
        tmem_ld16(tm + lane_base + (uint32_t)acc * 256 + 16 * c, v);
        tmem_ld_wait();
        if (MODE == 2) {
          const __half2* hc = reinterpret_cast<const __half2*>(&k0);
          const __half2* hd = reinterpret_cast<const __half2*>(&k1);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 c2 = __half22float2(hc[i]), d2 = __half22float2(hd[i]);
            v[2 * i] *= scale * c2.x; v[2 * i + 1] *= scale * c2.y;
            v[8 + 2 * i] *= scale * d2.x; v[8 + 2 * i + 1] *= scale * d2.y;
          }
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) w[i] = pack_f16x2_sat(v[2 * i], v[2 * i + 1]);
          *reinterpret_cast<uint4*>(out0 + toff + chunk_off(r, 2 * c, kTile)) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(out0 + toff + chunk_off(r, 2 * c + 1, kTile)) = make_uint4(w[4], w[5], w[6], w[7]);
        } else {
          if (MODE == 1 && a.layers_out) {
            const long long s = ((long long)a.tile_first[wi] + tile) * kTile + r;
            if (s < a.n_coords) {
              float* zd = a.layers_out + (long long)(a.layer + 1) * a.n_coords * f + s * f;
              for (int i = 0; i < 16; ++i)
                if (16 * c + i < f) zd[16 * c + i] = v[i] * inv_w;
            }
          }
          float cv[16];
          if (MODE == 0) {
            if (16 * c + 16 <= f) {
#pragma unroll
              for (int i = 0; i < 16; ++i) cv[i] = fast_cos(v[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) cv[i] = 16 * c + i < f ? fast_cos(v[i]) : 0.0f;
            }
          }
          sin_chunk16(v, 16 * c, f);
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) w[i] = pack_f16x2(v[2 * i], v[2 * i + 1]);
          *reinterpret_cast<uint4*>(out0 + toff + chunk_off(r, 2 * c, kTile)) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(out0 + toff + chunk_off(r, 2 * c + 1, kTile)) = make_uint4(w[4], w[5], w[6], w[7]);
          if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = pack_f16x2(cv[2 * i], cv[2 * i + 1]);
            *reinterpret_cast<uint4*>(out1 + toff + chunk_off(r, 2 * c, kTile)) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(out1 + toff + chunk_off(r, 2 * c + 1, kTile)) = make_uint4(w[4], w[5], w[6], w[7]);
          }
        }
      };
      if (MODE == 2) {
#pragma unroll
        for (int ci = 0; ci < 8; ++ci)
          if (c_lo + ci < c_hi) chunk(c_lo + ci, kc[2 * ci], kc[2 * ci + 1]);
      } else {
        for (int c = c_lo; c < c_hi; ++c) chunk(c, make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_acce[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// ---- last layer, loss, dz_NH (fit) / output epilogue (decode) -------------------------------------------------------------
// One CTA per tile, two threads per sample row (column halves of the F columns).
__global__ void __launch_bounds__(256) lw_last_kernel(LwArgs a) {
  __shared__ float s_y[2][kTile];
  __shared__ float s_red[8];
  const int wi = tc_find_work(a.work_prefix, a.n_work, blockIdx.x);
  const int net_id = a.work_net[wi];
  const NetDev& n = a.nets[net_id];
  const long long tile = (long long)(blockIdx.x - a.work_prefix[wi]) + a.tile_first[wi];
  const long long st = a.tile_base[wi] + (blockIdx.x - a.work_prefix[wi]);
  const int F = a.F, NH = n.L - 2, f = n.f, t = threadIdx.x, r = t & (kTile - 1), h = t >> 7;
  const size_t TB = lw_tile_bytes(F);
  const float* side = reinterpret_cast<const float*>(a.wpack + n.wpack_off + img_side_off(F, NH));
  const float* wl = side + 4 * F + NH * F;
  const float blast = wl[F];
  const unsigned char* act = a.scratch + a.in_off + (size_t)st * TB;
  const int G = F / 8, g_lo = h ? (G + 1) / 2 : 0, g_hi = h ? G : (G + 1) / 2;
  float part = 0.f;
  for (int g = g_lo; g < g_hi; ++g) {
    const uint4 k = *reinterpret_cast<const uint4*>(act + chunk_off(r, g, kTile));
    const __half2* hp = reinterpret_cast<const __half2*>(&k);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 v = __half22float2(hp[i]);
      part = fmaf(__ldg(wl + 8 * g + 2 * i), v.x, part);
      part = fmaf(__ldg(wl + 8 * g + 2 * i + 1), v.y, part);
    }
  }
  s_y[h][r] = part;
  __syncthreads();
  const float y = blast + s_y[0][r] + s_y[1][r];
  const long long total = a.eval ? (a.coords ? a.n_coords : n.n_vox) : (long long)n.batch;
  const long long s = tile * kTile + r;
  const bool ok = s < total;
  if (a.eval) {
    if (h == 0 && ok) {
      if (a.out_f32) {
        a.out_f32[s] = y;
      } else {
        void* dst = a.out_ptrs[net_id];
        if (a.out_dtype == 2) reinterpret_cast<float*>(dst)[s] = y;
        else if (a.out_dtype == 1) reinterpret_cast<unsigned short*>(dst)[s] = (unsigned short)(int)brief_denorm(n, y);
        else reinterpret_cast<unsigned char*>(dst)[s] = (unsigned char)(int)brief_denorm(n, y);
      }
    }
    return;
  }
  unsigned char* S = a.scratch;
  const float4 xf = reinterpret_cast<const float4*>(S + lw_off_xs(F, a.T, NH))[st * kTile + r];
  float dys = 0.f, ls = 0.f;
  if (ok) {
    const float e = y - xf.w;
    const float wt = (n.tau != 0.f && y <= n.tau) ? 1.0f : reinterpret_cast<const float*>(S + lw_off_ws(F, a.T, NH))[st * kTile + r];
    ls = wt * e * e;
    dys = kGradScale * wt * e;
  }
  if (h == 0) {
    unsigned char* dyb = S + lw_off_dyb(F, a.T, NH) + (size_t)st * kTile * 16 * 2;
    *reinterpret_cast<uint4*>(dyb + chunk_off(r, 0, kTile)) = make_uint4(pack_f16x2_sat(dys, 0.f), 0, 0, 0);
    *reinterpret_cast<uint4*>(dyb + chunk_off(r, 1, kTile)) = make_uint4(0, 0, 0, 0);
    // tile loss: fixed-order reduction (warp shuffles, then the four warps of the half in order)
    float acc = ls;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
    if ((t & 31) == 0) s_red[t >> 5] = acc;
  }
  // dz_NH = (w_h dy') Wlast cos(theta_NH)
  const unsigned char* cs = S + lw_off_cos(F, a.T, NH, NH) + (size_t)st * TB;
  unsigned char* dz = S + a.out_off + (size_t)st * TB;
  const float dw = dys * n.wh;
  for (int g = g_lo; g < g_hi; ++g) {
    const uint4 k = *reinterpret_cast<const uint4*>(cs + chunk_off(r, g, kTile));
    const __half2* hp = reinterpret_cast<const __half2*>(&k);
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 v = __half22float2(hp[i]);
      w[i] = pack_f16x2_sat(dw * __ldg(wl + 8 * g + 2 * i) * v.x, dw * __ldg(wl + 8 * g + 2 * i + 1) * v.y);
    }
    *reinterpret_cast<uint4*>(dz + chunk_off(r, g, kTile)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  __syncthreads();
  if (t == 0) reinterpret_cast<float*>(S + lw_off_tl(F, a.T, NH))[st] = ((s_red[0] + s_red[1]) + s_red[2]) + s_red[3];
}

// ---- dW: D[128 x NB] = A^T B summed over a slice of tiles in tensor memory ---------------------------------------------------
// CTA = (network, slice, half): output rows [128 half, 128 half + 128) of the layer.  A = column half of the A tiles seen
// MN-major (M = feature column of the tile, K = the 128 samples), B = the B tiles seen MN-major (N = NB columns).
//   kind 0: dW_l  (A = DZ_l, B = ACT_{l-1}, NB = F; column f of the result is db_l)
//   kind 1: dW0   (A = DZ_0, B = X block, NB = 16)
//   kind 2: dWlast(A = ACT_NH, B = DY block, NB = 16; row f of the result is dblast)
// Two-stage ring of (A half, B tile) pairs filled by a producer warp; one drain per slice into the partial slot.
constexpr int kDwThreads = 192;  // 4 drain warps + producer + MMA
__global__ void __launch_bounds__(kDwThreads, 1) lw_dw_kernel(LwArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar_full[2], bar_empty[2], bar_acc;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int per = a.max_slices * 2;
  const int wi = blockIdx.x / per, sl = (blockIdx.x % per) >> 1, half = blockIdx.x & 1;
  const NetDev& n = a.nets[a.work_net[wi]];
  const int F = a.F, NH = n.L - 2, f = n.f, F4 = n.F4, NB = a.nb;
  if (sl >= n.n_slices || 128 * half >= f + 2) return;  // CTA-uniform, before any barrier
  const int tps = n.slice_len / kTile;
  const int t_lo = sl * tps, t_hi = min(a.tile_count[wi], t_lo + tps);
  const long long tbase = a.tile_base[wi];
  const size_t TB = lw_tile_bytes(F);
  const int a_cols = min(128, F - 128 * half);                 // columns of the A half that exist in the tile
  const uint32_t a_bytes = (uint32_t)(a_cols / 8) * 2048u;
  const uint32_t b_bytes = (uint32_t)(NB / 8) * 2048u;
  const uint32_t stage = 128 * 128 * 2 + ((b_bytes + 127) & ~127u);
  if (t == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
    mbar_init(&bar_acc, 1);
    fence_mbar_init();
  }
  // rows of the A half beyond the tile's width are never loaded: keep them finite
  for (int i = t; i < (int)(2 * stage / 16); i += kDwThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  const unsigned char* srcA = a.scratch + a.in_off + (size_t)half * 16 * 2048;
  const unsigned char* srcB = a.scratch + a.in2_off;
  const size_t strideB = a.kind == 0 ? TB : (size_t)kTile * 16 * 2;
  if (warp == 4) {
    if (elect_one()) {
      int it = 0;
      for (int tile = t_lo; tile < t_hi; ++tile, ++it) {
        const int s = it & 1;
        if (it >= 2) mbar_wait(&bar_empty[s], (uint32_t)((it >> 1) - 1) & 1);
        mbar_expect_tx(&bar_full[s], a_bytes + b_bytes);
        bulk_g2s(smem + (size_t)s * stage, srcA + (size_t)(tbase + tile) * TB, a_bytes, &bar_full[s]);
        bulk_g2s(smem + (size_t)s * stage + 128 * 128 * 2, srcB + (size_t)(tbase + tile) * strideB, b_bytes, &bar_full[s]);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    const uint32_t idesc = make_idesc(128, NB, true, true);
    int it = 0;
    for (int tile = t_lo; tile < t_hi; ++tile, ++it) {
      const int s = it & 1;
      mbar_wait(&bar_full[s], (uint32_t)(it >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t ab = smem_u32(smem + (size_t)s * stage), bb = ab + 128 * 128 * 2;
        for (int k = 0; k < kTile / 16; ++k)
          mma_f16(tm, make_desc(ab + k * 2 * 128, 128, kActLBO), make_desc(bb + k * 2 * 128, 128, kActLBO), idesc,
                  (it > 0 || k > 0) ? 1u : 0u);
        commit(&bar_empty[s]);
        if (tile == t_hi - 1) commit(&bar_acc);
      }
      __syncwarp();
    }
  } else if (t_hi > t_lo) {
    // drain: thread = output row o of the half
    const int o = 128 * half + 32 * warp + lane;
    const float inv_count = 1.0f / ((float)n.batch * (float)n.out_dim);
    const float unscale = 2.0f * inv_count / kGradScale;
    float* part = a.partials + n.part_off + (long long)sl * n.P_dev;
    mbar_wait(&bar_acc, 0);
    tc_fence_after();
    const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
    for (int c = 0; c < NB / 16; ++c) {
      float v[16];
      tmem_ld16(tm + lane_base + 16 * c, v);
      tmem_ld_wait();
      if (a.kind == 0) {
        if (o < f) {
          float* wrow = part + dl_W(n, a.layer) + (long long)o * F4;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int k = 16 * c + i;
            if (k < f) wrow[k] = v[i] * unscale;
            else if (k == f) part[dl_b(n, a.layer) + o] = v[i] * unscale;
          }
        }
      } else if (a.kind == 1) {
        if (c == 0 && o < f) {
          part[dl_W0(n) + 4 * o + 0] = (v[0] + v[4]) * unscale;
          part[dl_W0(n) + 4 * o + 1] = (v[1] + v[5]) * unscale;
          if (n.in_dim == 3) part[dl_W0(n) + 4 * o + 2] = (v[2] + v[6]) * unscale;
          part[dl_b0(n) + o] = v[3] * unscale;
        }
      } else {
        if (c == 0) {
          if (o < f) part[dl_Wlast(n) + o] = v[0] * unscale;
          else if (o == f) part[dl_blast(n)] = v[0] * unscale;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}

// per (network, slice): loss = ordered sum of the slice's tile losses
__global__ void lw_loss_reduce_kernel(LwArgs a) {
  const int wi = blockIdx.x / a.max_slices, sl = blockIdx.x % a.max_slices;
  const NetDev& n = a.nets[a.work_net[wi]];
  if (sl >= n.n_slices || threadIdx.x != 0) return;
  const int NH = n.L - 2, tps = n.slice_len / kTile;
  const int t_lo = sl * tps, t_hi = min(a.tile_count[wi], t_lo + tps);
  const float* tl = reinterpret_cast<const float*>(a.scratch + lw_off_tl(a.F, a.T, NH)) + a.tile_base[wi];
  float acc = 0.f;
  for (int i = t_lo; i < t_hi; ++i) acc += tl[i];
  a.loss_partials[n.slice_off + sl] = acc / ((float)n.batch * (float)n.out_dim);
}

// ==================================================================================================================
// host side
// ==================================================================================================================
bool tc_lw_supported(int f, int L, int in_dim, int out_dim) {
  const int F = tc_fpad(f);
  return out_dim == 1 && (in_dim == 2 || in_dim == 3) && L >= 3 && F > 128 && F <= 256;
}
static size_t lw_gemm_smem(int F) { return (((size_t)F * F * 2 + 127) & ~(size_t)127) + (size_t)kLwStages * kLwStageBytes; }
static size_t lw_dw_smem(int F, int NB) { return 2 * ((size_t)128 * 128 * 2 + (((size_t)(NB / 8) * 2048 + 127) & ~(size_t)127)); }

template <int MODE>
static cudaError_t launch_gemm(const LwArgs& a, int n_ctas, cudaStream_t st) {
  const size_t smem = lw_gemm_smem(a.F);
  cudaError_t e = cudaFuncSetAttribute(lw_gemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  lw_gemm_kernel<MODE><<<n_ctas, kLwThreads, smem, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_lw_sample_l0(const LwArgs& a, int n_tiles, cudaStream_t st) {
  lw_sample_l0_kernel<<<n_tiles, 256, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_lw_gemm(const LwArgs& a, int mode, int n_ctas, cudaStream_t st) {
  return mode == 0 ? launch_gemm<0>(a, n_ctas, st) : mode == 1 ? launch_gemm<1>(a, n_ctas, st) : launch_gemm<2>(a, n_ctas, st);
}
cudaError_t launch_lw_last(const LwArgs& a, int n_tiles, cudaStream_t st) {
  lw_last_kernel<<<n_tiles, 256, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_lw_dw(const LwArgs& a, int n_nets, cudaStream_t st) {
  const size_t smem = lw_dw_smem(a.F, a.nb);
  cudaError_t e = cudaFuncSetAttribute(lw_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  lw_dw_kernel<<<n_nets * a.max_slices * 2, kDwThreads, smem, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_lw_loss_reduce(const LwArgs& a, int n_nets, cudaStream_t st) {
  lw_loss_reduce_kernel<<<n_nets * a.max_slices, 32, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace brief
