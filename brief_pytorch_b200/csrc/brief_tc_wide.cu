// brief_tc_wide.cu — the fit step (sampler gather + forward + datal2 + backward, main.py:126-163, 176-182, 385-396) on
// tcgen05 for networks too wide for the fully fused kernel of brief_tc.cu: 64 < F_PAD <= 128 (hipct 256^3 blocks, f = 113).
#include "brief_tc_common.cuh"

namespace brief {

using namespace umma;

// ==================================================================================================================
// fit, wide networks (64 < F_PAD <= 128: hipct 256^3 blocks, f = 113)
// ==================================================================================================================
// At these widths nothing of the narrow kernel's residency survives: one layer's weights are 32 KB (the image of an
// L = 7 network 168 KB), one tile's activations 32 KB per layer, one layer's dW accumulator 128 TMEM columns.  So the
// wide kernel keeps only WORKING SETS on chip and streams the rest through L2, one 128-sample tile at a time:
//   * weights:      two [F x F] buffers; layer l lives in buffer (l-1) & 1, fetched by bulk (TMA) copies two stages
//                   ahead in the forward pass and one stage ahead in the backward pass (the backward pass ends with
//                   W_1, W_2 resident — exactly what the next tile's forward pass starts with);
//   * activations:  two [128 x F] buffers (a_j in buffer j & 1).  The forward epilogue also writes a_j (j <= NH-2) in
//                   operand layout to a per-SM STASH in global memory (L2-resident: 4 x 32 KB per SM; one wide CTA
//                   is resident per SM, so the slot is indexed by %smid), from where the backward pass brings it back
//                   with one bulk copy per stage;
//   * dW:           per-tile accumulators of F columns, M = 128.  The five hidden layers of an L = 7 network would need
//                   640 TMEM columns beside theta and dX, so the sums over the slice live in two places: dW_3 in its
//                   own TMEM accumulator (lane = input feature), and the others (lane = output feature) are added per
//                   tile into a per-SM scratch in L2 by a plain 16-byte read-add-write whose reads are issued a whole
//                   stage early; the scratch is laid out so that a warp's access is 512 contiguous bytes.  Every
//                   element belongs to one thread and the tile order is fixed, so the fp32 sums are deterministic.
//                   The slice's partial slot is written once, at the end.
// Two threads per sample row (column halves), stages run back to back with CTA barriers: the MMA of a stage, its
// epilogue and the dW update are not overlapped with each other (only the bulk copies and the scratch reads run ahead).  Same numerics as the
// narrow kernel (fp16 operands, fp32 accumulation, hi/lo layer 0, kGradScale).
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
// D[128 x N] = A^T B over the tile's 128 samples (both operands MN-major, K = samples); rows past the operand's width
// read whatever follows the buffer and are ignored by the drain
template <int N>
__device__ __forceinline__ void issue_dw128(uint32_t d, uint32_t a_buf, uint32_t b_buf, bool accumulate = false) {
  constexpr uint32_t idesc = make_idesc(128, N, true, true);
#pragma unroll
  for (int k = 0; k < kTile / 16; ++k)
    mma_f16(d, make_desc(a_buf + k * 2 * 128, 128, kActLBO), make_desc(b_buf + k * 2 * 128, 128, kActLBO), idesc,
            (accumulate || k > 0) ? 1u : 0u);
}

constexpr int kWideThreads = 256;
__host__ __device__ constexpr size_t wide_resident_bytes(int F, int NH) { return img_bytes(F, NH) - img_l0_off(F, NH); }

template <int F>
__global__ void __launch_bounds__(kWideThreads, 1) tc_fit_wide_kernel(FitArgs a) {
  constexpr int NC = F / 16, NCH = (NC + 1) / 2;  // 16-column chunks per row / per thread (two threads per row)
  constexpr uint32_t BUF = kTile * F * 2, BLK = kTile * 16 * 2;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ NetDev sn;
  __shared__ __align__(8) uint64_t bar_res, bar_mma, bar_dw, bar_w[2], bar_act[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float4 s_g[kTile];
  __shared__ float s_gw[kTile];
  __shared__ float s_y[2][kTile];
  __shared__ float s_red[4];

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int q = warp & 3, cg = warp >> 2, r = 32 * q + lane;
  const int c_lo = cg ? (NC + 1) / 2 : 0, c_hi = cg ? NC : (NC + 1) / 2;  // this thread's 16-column chunks
#ifdef BRIEF_TC_TIMING
  const int tslot = warp == 0 ? 0 : -1;
#endif
  TT(k_start);
  const int wi = tc_find_work(a.work_prefix, a.n_work, blockIdx.x);
  const int net_id = a.work_net[wi];
  const int slice = blockIdx.x - a.work_prefix[wi];
  tc_load_net(sn, a.nets[net_id]);
  if (t == 0) {
    mbar_init(&bar_res, 1);
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_dw, 1);
    mbar_init(&bar_w[0], 1);
    mbar_init(&bar_w[1], 1);
    mbar_init(&bar_act[0], 1);
    mbar_init(&bar_act[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const NetDev& n = sn;
  const int NH = n.L - 2, f = n.f, F4 = n.F4;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  unsigned char* sA = smem;             // a_j in sA + (j & 1) * BUF
  unsigned char* sDz = sA + 2 * BUF;    // dz_l in sDz + ((NH - l) & 1) * BUF
  unsigned char* sWt = sDz + 2 * BUF;   // W_l in sWt + ((l - 1) & 1) * BUF
  unsigned char* sX = sWt + 2 * BUF;
  unsigned char* sDY = sX + BLK;
  unsigned char* sRes = sDY + BLK;      // layer-0 block | last block | fp32 side block of the image
  const float* side = reinterpret_cast<const float*>(sRes + (img_side_off(F, NH) - img_l0_off(F, NH)));
  const float* s_wl = side + 4 * F + NH * F;
  const float* s_bl = s_wl + F;
  const unsigned char* img = a.wpack + n.wpack_off;
  if (t == 0) {
    const uint32_t bytes = (uint32_t)wide_resident_bytes(F, NH);
    mbar_expect_tx(&bar_res, bytes);
    bulk_g2s(sRes, img + img_l0_off(F, NH), bytes, &bar_res);
  }
  if (cg == 0) *reinterpret_cast<uint4*>(sDY + chunk_off(r, 1, kTile)) = make_uint4(0, 0, 0, 0);
  float* part = a.partials + n.part_off + (long long)slice * n.P_dev;
  for (int i = t; i < (n.P_dev >> 2); i += kWideThreads) reinterpret_cast<float4*>(part)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  // per-SM scratch (one wide CTA is resident per SM: the slot of this SM is ours until the CTA exits), so its size does
  // not grow with the grid and later CTAs of a multi-wave launch reuse lines that are already in L2
  unsigned int smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  // smid < %nsmid == a.stash_slots: the host sizes the scratch from the device's own %nsmid (brief_capi.cu, finalize)
  unsigned char* stash = a.stash + (size_t)smid * a.stash_stride;  // a_j (j <= NH-2) at stash + j * BUF
  float4* scr = reinterpret_cast<float4*>(stash + (size_t)(NH >= 2 ? NH - 1 : 0) * BUF);  // running dW sums, see below

  const uint32_t tm = tmem_base_s;
  const uint32_t TZ = tm, TXB = tm + F, TDW = tm + 2 * F, TDW3 = tm + 3 * F;  // TDW3: dW_3, summed over the slice
  const uint32_t lane_base = (uint32_t)(32 * q) << 16;
  const uint32_t aA = smem_u32(sA), aDz = smem_u32(sDz), aWt = smem_u32(sWt), aX = smem_u32(sX), aDY = smem_u32(sDY),
                 aL0 = smem_u32(sRes);
  const float wh = n.wh, w0 = n.w0;
  const long long s_begin = (long long)slice * n.slice_len;
  const long long s_end = min((long long)n.batch, s_begin + n.slice_len);
  const int n_tiles = (int)((s_end - s_begin + kTile - 1) / kTile);
  const float inv_count = 1.0f / ((float)n.batch * (float)n.out_dim);
  const float unscale = 2.0f * inv_count / kGradScale;
  float loss_acc = 0.f;
  uint32_t ph_mma = 0, ph_dw = 0;
  // warp 0 only (kept warp-uniform): phase and "copy in flight" of the weight / activation buffers
  uint32_t ph_w[2] = {0, 0}, ph_act[2] = {0, 0};
  bool pend_w[2] = {false, false}, pend_act[2] = {false, false};

  auto cta_sync = [&]() {  // operand rows written / TMEM read -> the next stage's MMAs may touch them
    tc_fence_before();
    fence_async_smem();
    __syncthreads();
    tc_fence_after();
  };
  auto load_w = [&](int l) {  // warp 0: W_l -> its buffer
    const int b = (l - 1) & 1;
    if (elect_one()) {
      mbar_expect_tx(&bar_w[b], (uint32_t)F * F * 2);
      bulk_g2s(sWt + (size_t)b * BUF, img + (size_t)(l - 1) * F * F * 2, (uint32_t)F * F * 2, &bar_w[b]);
    }
    __syncwarp();
    pend_w[b] = true;
  };
  auto need_w = [&](int l) {
    const int b = (l - 1) & 1;
    if (pend_w[b]) { mbar_wait(&bar_w[b], ph_w[b]); ph_w[b] ^= 1; pend_w[b] = false; }
  };
  auto load_act = [&](int j) {  // warp 0: a_j from the stash -> its buffer
    const int b = j & 1;
    if (elect_one()) {
      mbar_expect_tx(&bar_act[b], BUF);
      bulk_g2s(sA + (size_t)b * BUF, stash + (size_t)j * BUF, BUF, &bar_act[b]);
    }
    __syncwarp();
    pend_act[b] = true;
  };
  auto need_act = [&](int j) {
    const int b = j & 1;
    if (pend_act[b]) { mbar_wait(&bar_act[b], ph_act[b]); ph_act[b] ^= 1; pend_act[b] = false; }
  };
  auto wait_mma = [&]() { mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1; tc_fence_after(); };
  auto wait_dw = [&]() { mbar_wait(&bar_dw, ph_dw); ph_dw ^= 1; tc_fence_after(); };
  constexpr uint32_t idesc_f = make_idesc(128, F, false, false);
  auto layer0 = [&]() {  // theta_0 = A0 [128 x 16] * B0^T
    mma_f16(TZ, make_desc(aX, kActLBO, 128), make_desc(aL0, (F / 8) * 128, 128), idesc_f, 0);
  };

  if (warp == 0 && n_tiles > 0) {
    if (NH >= 1) load_w(1);
    if (NH >= 2) load_w(2);
  }
  mbar_wait(&bar_res, 0);
  __syncthreads();

  // sampler (main.py:126-163 / whole-block cube): one row per thread of the first column group, fetched one tile ahead
  // (the index -> voxel / axis-table loads of tile k+1 are issued under the backward pass of tile k)
  float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
  float nxt_w = 0.f;
  auto fetch_sample = [&](int k) {
    const long long s = s_begin + (long long)k * kTile + r;
    const bool ok = s < s_end;
    long long v = 0;
    if (ok) {
      if (n.mode == 0) v = s;
      else if (a.idx) v = a.idx[n.idx_off + s];
      else v = brief_sample_index(a.seed, a.state ? a.state->step : a.step, n.stream_id, (uint64_t)s, (uint64_t)n.n_vox);
    }
    const float raw = brief_raw_value(n, v);
    float x0, x1, x2;
    brief_coords(n, a.axes, v, x0, x1, x2);
    nxt = ok ? make_float4(x0, x1, x2, brief_normalize(n, raw)) : make_float4(0.f, 0.f, 0.f, 0.f);
    nxt_w = ok ? brief_weight(n, v, raw) : 0.f;
  };
  if (cg == 0 && n_tiles > 0) fetch_sample(0);

  for (int k = 0; k < n_tiles; ++k) {
    TT(t0);
    if (cg == 0) {
      const float x0 = nxt.x, x1 = nxt.y, x2 = nxt.z;
      s_gw[r] = nxt_w;
      s_g[r] = nxt;
      // layer-0 operand row [x_hi(3) 1 x_lo(3) 1 | x_hi(3) 0 0 0 0 0]; also the B operand of dW0
      const float h0 = __half2float(__float2half_rn(x0)), h1 = __half2float(__float2half_rn(x1)),
                  h2 = __half2float(__float2half_rn(x2));
      const uint32_t p01 = pack_f16x2(h0, h1), p21 = pack_f16x2(h2, 1.0f);
      *reinterpret_cast<uint4*>(sX + chunk_off(r, 0, kTile)) =
          make_uint4(p01, p21, pack_f16x2(x0 - h0, x1 - h1), pack_f16x2(x2 - h2, 1.0f));
      *reinterpret_cast<uint4*>(sX + chunk_off(r, 1, kTile)) = make_uint4(p01, pack_f16x2(h2, 0.f), 0u, 0u);
    }
    cta_sync();
    { TT(t1); TACC(0, t1 - t0); }
    const float4 xf = s_g[r];

    // ---- forward: stage j computes theta_j (TZ) and a_j
    float ypart = 0.f;
    for (int j = 0; j <= NH; ++j) {
      TT(f0);
      if (warp == 0) {
        if (j >= 1) need_w(j);
        if (elect_one()) {
          if (j == 0) layer0();
          else issue_forward<F>(TZ, aA + (uint32_t)((j - 1) & 1) * BUF, aWt + (uint32_t)((j - 1) & 1) * BUF);
          commit(&bar_mma);
        }
        __syncwarp();
      }
      wait_mma();
      TT(f1);
      if (warp == 0 && j >= 1 && j + 2 <= NH) load_w(j + 2);  // the buffer of W_j is free again
      unsigned char* dst = sA + (size_t)(j & 1) * BUF;
      unsigned char* gst = stash + (size_t)j * BUF;
      const bool to_stash = j <= NH - 2;
      float vb[2][16];
      tmem_ld16(TZ + lane_base + 16 * c_lo, vb[0]);
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci) {
        const int c = c_lo + ci;
        if (c >= c_hi) break;
        tmem_ld_wait();
        if (c + 1 < c_hi) tmem_ld16(TZ + lane_base + 16 * (c + 1), vb[(ci + 1) & 1]);  // under this chunk's sines
        float* v = vb[ci & 1];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fast_sin(v[i]);
        const uint4 lo4 = make_uint4(pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]), pack_f16x2(v[4], v[5]), pack_f16x2(v[6], v[7]));
        const uint4 hi4 = make_uint4(pack_f16x2(v[8], v[9]), pack_f16x2(v[10], v[11]), pack_f16x2(v[12], v[13]), pack_f16x2(v[14], v[15]));
        const uint32_t o0 = chunk_off(r, 2 * c, kTile), o1 = chunk_off(r, 2 * c + 1, kTile);
        *reinterpret_cast<uint4*>(dst + o0) = lo4;
        *reinterpret_cast<uint4*>(dst + o1) = hi4;
        if (to_stash) {
          *reinterpret_cast<uint4*>(gst + o0) = lo4;
          *reinterpret_cast<uint4*>(gst + o1) = hi4;
        }
        if (j == NH) {
#pragma unroll
          for (int i = 0; i < 16; ++i) ypart = fmaf(s_wl[16 * c + i], v[i], ypart);
        }
      }
      if (j == NH) {
        s_y[cg][r] = ypart;
        fence_async_all();  // the stash rows written above (generic proxy, global) -> the backward pass's bulk copies
      }
      TT(f2);
      cta_sync();
      { TT(f3); TACC(1, f1 - f0); TACC(2, f2 - f1); TACC(3, f3 - f2); }
    }

    // ---- loss (datal2, main.py:176-182), scaled output gradient, dWlast, dz_NH (theta_NH is still in TZ)
    TT(l0);
    const float y = s_bl[0] + s_y[0][r] + s_y[1][r];
    float dys = 0.f;
    if (s_begin + (long long)k * kTile + r < s_end) {
      const float e = y - xf.w;
      const float wt = (n.tau != 0.f && y <= n.tau) ? 1.0f : s_gw[r];
      if (cg == 0) loss_acc = fmaf(wt * e, e, loss_acc);
      dys = kGradScale * wt * e;
    }
    if (cg == 0) *reinterpret_cast<uint4*>(sDY + chunk_off(r, 0, kTile)) = make_uint4(pack_f16x2_sat(dys, 0.f), 0, 0, 0);
    cta_sync();
    TT(l1);
    // dWlast (+ dblast in row f) = a_NH^T dY first: as soon as it has consumed a_NH, the bulk copy of a_{NH-2} into that
    // buffer starts and runs under the dz_NH pass below
    if (warp == 0) {
      if (elect_one()) {
        issue_dw128<16>(TDW, aA + (uint32_t)(NH & 1) * BUF, aDY);
        commit(&bar_dw);
      }
      __syncwarp();
    }
    if (cg == 0 && k + 1 < n_tiles) fetch_sample(k + 1);
    wait_dw();
    if (warp == 0 && NH >= 2) load_act(NH - 2);
    dys *= wh;
    for (int c = c_lo; c < c_hi; ++c) {
      float v[16];
      tmem_ld16(TZ + lane_base + 16 * c, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = dys * s_wl[16 * c + i] * fast_cos(v[i]);
      store_chunk16_both<true>(sDz, r, c, v, false, 0);
    }
    if (cg == 0) {
      float v[16];
      tmem_ld16(TDW + lane_base, v);
      tmem_ld_wait();
      if (r < f) red_add_f32(part + dl_Wlast(n) + r, v[0] * unscale);
      else if (r == f) red_add_f32(part + dl_blast(n), v[0] * unscale);
    }
    cta_sync();
    { TT(l2); TACC(4, l1 - l0); TACC(5, l2 - l1); }

    // ---- backward: stage l turns dz_l into dz_{l-1} and drains dW_l
    for (int l = NH; l >= 1; --l) {
      const uint32_t dz_l = aDz + (uint32_t)((NH - l) & 1) * BUF;
      TT(b0);
      // layers summed in the CTA's scratch (all but 3): this thread's 16 float4 of the running sum are read NOW,
      // so that the L2 latency sits under the MMAs and the cosine epilogue of this stage.  Scratch layout
      // [chunk][float4 i][row r]: a warp's access is 512 contiguous bytes (weight rows in the slot have a 464-byte
      // pitch: the same float4 access there touches 32 lines per instruction and ran 8x slower).
      const bool drained = l != 3;
      float4* const scr_l = scr + (size_t)(l < 3 ? l - 1 : l - 2) * (NC * 4 * kTile) + r;  // + (16-col chunk * 4 + i) * 128
      if (warp == 0) {
        if (l >= 2) { need_w(l - 1); need_act(l - 2); }
        need_w(l);
        if (elect_one()) {
          if (l >= 2) issue_forward<F>(TZ, aA + (uint32_t)(l & 1) * BUF, aWt + (uint32_t)(l & 1) * BUF);  // theta_{l-1}
          else layer0();
          issue_dx<F>(TXB, dz_l, aWt + (uint32_t)((l - 1) & 1) * BUF);                                     // dX_{l-1}
          commit(&bar_mma);
          if (l == 3) issue_dw128<F>(TDW3, aA + (uint32_t)((l - 1) & 1) * BUF, dz_l, k > 0);  // dW_3^T, resident
          else issue_dw128<F>(TDW, dz_l, aA + (uint32_t)((l - 1) & 1) * BUF);                  // dW_l [out][in]
          commit(&bar_dw);
        }
        __syncwarp();
      }
      float4 pf[NCH][4];
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          pf[ci][i] = (drained && k > 0 && c_lo + ci < c_hi) ? scr_l[((c_lo + ci) * 4 + i) * kTile] : make_float4(0.f, 0.f, 0.f, 0.f);
      wait_mma();
      TT(b1);
      if (warp == 0 && l >= 3) load_w(l - 2);  // into the buffer of W_l (dX_{l-1} is done)
      {
        unsigned char* dzb = sDz + (size_t)((NH - l + 1) & 1) * BUF;  // dz_{l-1}
        const float scale = l >= 2 ? 1.0f : w0 / wh;  // dX carries w_hidden (omega-scaled weights); layer 0 wants w_0
        float zb[2][16], xb[2][16];
        tmem_ld16(TZ + lane_base + 16 * c_lo, zb[0]);
        tmem_ld16(TXB + lane_base + 16 * c_lo, xb[0]);
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          const int c = c_lo + ci;
          if (c >= c_hi) break;
          tmem_ld_wait();
          if (c + 1 < c_hi) {
            tmem_ld16(TZ + lane_base + 16 * (c + 1), zb[(ci + 1) & 1]);
            tmem_ld16(TXB + lane_base + 16 * (c + 1), xb[(ci + 1) & 1]);
          }
          float* z = zb[ci & 1];
          const float* x = xb[ci & 1];
#pragma unroll
          for (int i = 0; i < 16; ++i) z[i] = x[i] * scale * fast_cos(z[i]);
          store_chunk16_both<true>(dzb, r, c, z, false, 0);
        }
      }
      TT(b2);
      wait_dw();
      TT(b3);
      if (warp == 0 && l >= 3) load_act(l - 3);  // into the buffer of a_{l-1} (dW_l is done)
      // dW_l of this tile: layer 3 stays in its own TMEM accumulator across the slice's tiles, the others are added into
      // the CTA's scratch in L2: every element belongs to one thread, so a plain 16-byte read (above) - add - write
      // replaces atomics (red.global.add ran at ~1 per clock per SM here)
      if (drained) {
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          if (c_lo + ci < c_hi) {
            float v[16];
            tmem_ld16(TDW + lane_base + 16 * (c_lo + ci), v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 4; ++i)
              scr_l[((c_lo + ci) * 4 + i) * kTile] = make_float4(pf[ci][i].x + v[4 * i], pf[ci][i].y + v[4 * i + 1],
                                                                 pf[ci][i].z + v[4 * i + 2], pf[ci][i].w + v[4 * i + 3]);
          }
        }
      }
      TT(b4);
      cta_sync();
      { TT(b5); TACC(6, b1 - b0); TACC(7, b2 - b1); TACC(8, b3 - b2); TACC(9, b4 - b3); TACC(10, b5 - b4); }
    }
    TT(d0);
    // ---- dW0 (+ db0): dz_0^T [x_hi, 1, x_lo, ...]; lane = output feature
    if (warp == 0) {
      if (elect_one()) {
        issue_dw128<16>(TDW, aDz + (uint32_t)(NH & 1) * BUF, aX);
        commit(&bar_dw);
      }
      __syncwarp();
    }
    wait_dw();
    if (cg == 0) {
      float v[16];
      tmem_ld16(TDW + lane_base, v);
      tmem_ld_wait();
      if (r < f) {
        red_add_f32(part + dl_W0(n) + 4 * r + 0, (v[0] + v[4]) * unscale);
        red_add_f32(part + dl_W0(n) + 4 * r + 1, (v[1] + v[5]) * unscale);
        if (n.in_dim == 3) red_add_f32(part + dl_W0(n) + 4 * r + 2, (v[2] + v[6]) * unscale);
        red_add_f32(part + dl_b0(n) + r, v[3] * unscale);
      }
    }
    cta_sync();
    { TT(d1); TACC(11, d1 - d0); TACC(12, 1); TACC(13, d1 - t0); }
  }

  // ---- slice epilogue: the dW sums -> the slot (zeroed above; every element has exactly one writer)
  if (n_tiles > 0) {
    for (int l = 1; l <= NH; ++l) {  // scratch layers: lane r = output feature, chunks of 16 input features
      if (l == 3) continue;
      const float4* scr_l = scr + (size_t)(l < 3 ? l - 1 : l - 2) * (NC * 4 * kTile) + r;
      float* const wrow = part + dl_W(n, l) + r * F4;
      if (r < f) {
        for (int c = c_lo; c < c_hi; ++c)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int k0 = 16 * c + 4 * i;
            if (k0 > f) continue;
            const float4 v = scr_l[(c * 4 + i) * kTile];
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (k0 + j < f) wrow[k0 + j] = e[j] * unscale;
              else if (k0 + j == f) part[dl_b(n, l) + r] = e[j] * unscale;
            }
          }
      }
    }
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) {
      if (c_lo + ci < c_hi) {
        float v3[16];
        if (NH >= 3) {
          tmem_ld16(TDW3 + lane_base + 16 * (c_lo + ci), v3);
          tmem_ld_wait();
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int o = 16 * (c_lo + ci) + i;
          if (o < f && r <= f) {
            if (NH >= 3) __stcg(r < f ? part + dl_W(n, 3) + r + o * F4 : part + dl_b(n, 3) + o, v3[i] * unscale);
          }
        }
      }
    }
  }
  if (cg == 0) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) loss_acc += __shfl_down_sync(0xffffffffu, loss_acc, off);
    if (lane == 0) s_red[q] = loss_acc;
  }
  tc_fence_before();
  __syncthreads();
  if (t == 0) a.loss_partials[n.slice_off + slice] = (((s_red[0] + s_red[1]) + s_red[2]) + s_red[3]) * inv_count;
  if (warp == 0) tmem_dealloc(tm, 512);
}

// ==================================================================================================================
// host side
// ==================================================================================================================
// wide fit kernel (64 < F_PAD <= 128): 2 activation + 2 dz + 2 weight buffers, X, dY, the resident tail of the image
size_t tc_wide_smem(int F, int L) {
  return (size_t)6 * kTile * F * 2 + 2 * (size_t)kTile * 16 * 2 + ((wide_resident_bytes(F, L - 2) + 127) & ~(size_t)127);
}
bool tc_wide_supported(int f, int L, int in_dim, int out_dim) {
  const int F = tc_fpad(f);
  if (out_dim != 1 || (in_dim != 2 && in_dim != 3)) return false;
  if (L < 3 || F <= 64 || F > 128) return false;
  if (tc_wide_smem(F, L) > 221 * 1024) return false;
  return tc_eval_groups(F, L) >= 1;
}
// bytes of the wide kernel's per-CTA scratch: the activation stash (a_0 .. a_{NH-2}) followed by the running dW sums of
// the layers that are not kept in TMEM (all but 3); 0 for the narrow kernel
size_t tc_fit_stash_bytes(int F, int L) {
  if (F <= 64) return 0;
  const int NH = L - 2;
  const int scratch_layers = NH - (NH >= 3 ? 1 : 0);
  return (size_t)(NH >= 2 ? NH - 1 : 0) * kTile * F * 2 + (size_t)scratch_layers * (F / 16) * 4 * kTile * 16;
}

template <int F>
static cudaError_t launch_fit_wide_f(const FitArgs& a, int L_max, int n_blocks, cudaStream_t st) {
  const size_t smem = tc_wide_smem(F, L_max);
  if (tc_fit_stash_bytes(F, L_max) > 0 && (a.stash == nullptr || a.stash_stride < tc_fit_stash_bytes(F, L_max)))
    return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(tc_fit_wide_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  tc_fit_wide_kernel<F><<<n_blocks, kWideThreads, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_tc_fit_wide(const FitArgs& a, int F_PAD, int L_max, int n_blocks, cudaStream_t st) {
  switch (F_PAD) {
    case 80: return launch_fit_wide_f<80>(a, L_max, n_blocks, st);
    case 96: return launch_fit_wide_f<96>(a, L_max, n_blocks, st);
    case 112: return launch_fit_wide_f<112>(a, L_max, n_blocks, st);
    case 128: return launch_fit_wide_f<128>(a, L_max, n_blocks, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace brief

#ifdef BRIEF_TC_TIMING
extern "C" int brief_debug_read_timing_wide(unsigned long long* out, int reset) {
  cudaMemcpyFromSymbol(out, brief::g_tc_timing, sizeof(unsigned long long) * 64);
  if (reset) { unsigned long long z[64] = {0}; cudaMemcpyToSymbol(brief::g_tc_timing, z, sizeof z); }
  return 0;
}
#endif
