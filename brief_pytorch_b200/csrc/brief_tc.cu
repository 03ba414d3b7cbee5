// brief_tc.cu — the tensor-core path (BRIEF_PREC_F16): grouped, fully fused SIREN kernels on tcgen05.
//
// Reference work replaced (file:line relative to the reference root):
//   SIREN.forward / reconstruct_flattened / invnormalize_data   utils/Networks.py:269-271, utils/misc.py:59-92,
//                                                               utils/io.py:136-147            -> tc_eval_kernel
//   sampler gather + forward + datal2 + backward               main.py:126-163, 176-182, 385-396 -> tc_fit_kernel
//
// Shape of the computation.  A tile is 128 samples (UMMA M = 128): one thread per sample row == TMEM lane, a warp may
// only touch the lane quadrant 32 * (warp % 4).  EVERY sine layer's argument comes out of a tcgen05.mma (fp16 operands,
// fp32 accumulator in TMEM): the packed operand image (brief_image.cuh) carries omega and the bias inside the weights
// (two constant-one activation columns f, f+1 hold the bias as an fp16 hi/lo pair), layer 0 is one K = 16 MMA against the
// hi/lo-split coordinate row, and in the decompress kernel the last layer is an N = 16 MMA against hi/lo-split Wlast.
// An epilogue therefore is: tcgen05.ld the accumulator row, FMUL.RZ + MUFU.SIN per element, pack to fp16, write the row
// into the next contraction's operand (shared memory; in the fit kernel also tensor memory).  A network's image is staged
// ONCE per CTA by one bulk (TMA) copy and stays resident for all of the CTA's tiles.  MMAs are issued by a dedicated warp
// (a tcgen05.mma blocks its issuing thread for about as long as it executes), told through mbarriers when a tile's
// operand rows are written; the epilogue warps only ever wait for the commit of the MMAs they consume.
//
// Why fp16 and not bf16: activations are sines (|a| <= 1) and weights are << 1, so fp16's 11-bit significand is
// usable without range problems and gives 8x smaller rounding error than bf16 at the same tensor rate.  A and B of
// one tcgen05.mma must share a format (mixing is an illegal instruction), so the backward operands are fp16 too:
// dz is carried as dz * kGradScale * (B*C)/2 — i.e. w*(yhat-y)/256 at the output — which keeps it inside fp16's
// normal range with orders of magnitude to spare on both sides; conversions saturate instead of overflowing and
// the scale is removed in fp32 when the accumulated dW leaves TMEM.
//
// Operand layout: brief_umma.cuh ("interleaved" 8x8 cores).  Width f is padded to F = F_PAD (multiple of 16,
// F >= f + 2); pad weights are zero.  Because columns f, f+1 of every activation buffer are the constant 1, the bias
// gradients fall out of the dW contractions as column f.
//
// Decompress kernel: up to 4 groups x 2 alternating tile slots per CTA share one image (8 tiles in flight per SM), one
// MMA-issue warp per group at F = 48 / 64; at F >= 112 the hidden weights are streamed layer by layer through two buffers
// so that four tiles fit instead of one; bound by the special-function unit.  Fit kernel: two tiles in flight on separate
// warp groups (forward / backward), ONE MMA-ISSUE WARP PER CHAIN and two sampler warps (F >= 32), dz_NH formed by the
// backward group, dW accumulated across the slice's tiles in TMEM; bound by the sine / cosine epilogues of two chains on
// one SFU.  Networks with 64 < F_PAD <= 128 fit on the wide kernel of brief_tc_wide.cu (streamed
// weights, stashed activations), 128 < F_PAD <= 256 on the layer-wise kernels of brief_tc_lw.cu.
#include "brief_tc_common.cuh"

namespace brief {

using namespace umma;

// ==================================================================================================================
// forward / decompress
// ==================================================================================================================
// One CTA runs G GROUPS of 4 warps (128 threads: thread = one sample row = one TMEM lane) plus one MMA-issue warp.
// Every group owns TWO tile slots and alternates between them, so 2G tiles (up to 8) are in flight per CTA and share
// ONE staged copy of the weights:
//
//     coords -> A0 row -> [MMA layer 0] -> sin -> [MMA hidden 1] -> sin -> ... -> [MMA last (N = 16)] -> y -> store
//
// While a group computes the sines of one slot, the MMA of its other slot executes; MMA latency, TMEM loads, fences and
// stores of one tile always sit under another tile's sines, which leaves the special-function unit (16 sin/clk/SM) as
// the binding pipe.  ALL layers run on the tensor core: layer 0 as one K = 16 MMA against the hi/lo-split coordinate
// row (fp32-grade), the last layer as N = 16 MMAs against hi/lo-split Wlast (brief_image.cuh); the bias of every layer
// rides inside the MMA, so an epilogue element is FMUL.RZ + MUFU.SIN + half a pack.  There is no CTA-wide or
// group-wide barrier in the tile loop: a warp only ever touches its own 32 rows, and the MMA warp serves the
// (slot, group) units in a fixed order, blocking on each unit's "operand rows written" mbarrier (4 warp arrivals).
constexpr int kEvalMaxGroups = 4;  // 4 groups x 2 slots x F columns <= 512 TMEM columns at F = 64
constexpr int kEvalSlots = 2;
constexpr int kEvalMaxUnits = kEvalMaxGroups * kEvalSlots;
__host__ __device__ constexpr size_t eval_unit_bytes(int F) { return (size_t)kTile * F * 2 + (size_t)kTile * 16 * 2; }

#ifndef BRIEF_EVAL_ISSUERS
#define BRIEF_EVAL_ISSUERS 4  // MMA-issue warps of the decode kernel: issuer i serves the units of groups i, i + 4, ... in a fixed order
#endif
// F <= 32: two CTAs per SM, register-bound; F > 64: one or two units per CTA, nothing to interleave
__host__ __device__ constexpr int eval_issuers(int F) { return (F <= 32 || F > 64) ? 1 : BRIEF_EVAL_ISSUERS; }
// Wide decode (F >= 112): the hidden weights do not stay resident (5 x 32 KB at F = 128 would leave room for ONE tile, i.e.
// no overlap at all).  The units of a CTA walk the layers in lockstep anyway (fixed service order), so ONE layer's weights at
// a time are enough: a producer warp streams W_1 .. W_NH, round after round, through two [F x F] buffers (bulk copies,
// freed by tcgen05.commit), and the shared memory saved holds four tiles in flight instead of one.
__host__ __device__ constexpr bool eval_streams(int F) { return F >= 112; }
__host__ __device__ constexpr size_t eval_tail_bytes(int F, int NH) { return (img_bytes(F, NH) - img_l0_off(F, NH) + 127) & ~(size_t)127; }
__host__ __device__ constexpr size_t eval_weight_bytes(int F, int NH) {  // shared memory in front of the tile units
  return eval_streams(F) ? eval_tail_bytes(F, NH) + 2 * (size_t)F * F * 2 : img_bytes_padded(F, NH);
}
template <int F, bool DUMP>
__global__ void __launch_bounds__(kEvalMaxGroups * 128 + 32 * eval_issuers(F) + (eval_streams(F) ? 32 : 0), F <= 32 ? 2 : 1)
    tc_eval_kernel(EvalArgs a) {
  constexpr int kEvalIssuers = eval_issuers(F);
  constexpr bool STREAM = eval_streams(F);
  constexpr int NC = F / 16;  // 16-column chunks per row
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ NetDev sn;
  __shared__ __align__(8) uint64_t bar_w, bar_mma[kEvalMaxUnits], bar_r[kEvalMaxUnits], bar_wfull[2], bar_wfree[2];
  __shared__ uint32_t tmem_base_s;

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int G = (blockDim.x - 32 * kEvalIssuers - (STREAM ? 32 : 0)) >> 7, S = a.eval_slots, U = G * S;  // S tile slots per group
  const bool mma_warp = warp >= 4 * G && warp < 4 * G + kEvalIssuers;
  const bool producer_warp = STREAM && warp == 4 * G + kEvalIssuers;
  const int issuer = warp - 4 * G;  // which of the issue warps (mma_warp only)
  const int g = warp >> 2, q = warp & 3, r = 32 * q + lane;
  int net_id;
  long long chunk;
  if (a.single_net >= 0) {
    net_id = a.single_net;
    chunk = blockIdx.x;
  } else {
    const int wi = tc_find_work(a.work_prefix, a.n_work, blockIdx.x);
    net_id = a.work_net[wi];
    chunk = blockIdx.x - a.work_prefix[wi];
    if (a.out_ptrs[net_id] == nullptr) return;  // this network is not part of the call (CTA-uniform, before any barrier)
  }
  tc_load_net(sn, a.nets[net_id]);
  if (t == 0) {
    mbar_init(&bar_w, 1);
    for (int i = 0; i < U; ++i) {
      mbar_init(&bar_mma[i], 1);
      mbar_init(&bar_r[i], 4);  // "this unit's operand rows are written": one arrival per warp of the group
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_wfull[i], 1); mbar_init(&bar_wfree[i], 1); }
    fence_mbar_init();
  }
  const uint32_t tcols = (uint32_t)tmem_cols_pow2(U * F);
  if (warp == 0) tmem_alloc(&tmem_base_s, tcols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const NetDev& n = sn;
  const int NH = n.L - 2;
  // resident: the whole packed image, or (STREAM) its tail [layer-0 block | last block | side] followed by two weight buffers
  unsigned char* sW = smem;
  unsigned char* sWbuf = smem + eval_tail_bytes(F, NH);
  unsigned char* sUnits = sW + eval_weight_bytes(F, NH);  // per unit: [128 x F] activations | [128 x 16] layer-0 rows
  if (t == 0) {
    const uint32_t bytes = (uint32_t)(STREAM ? img_bytes(F, NH) - img_l0_off(F, NH) : img_bytes(F, NH));
    mbar_expect_tx(&bar_w, bytes);
    bulk_g2s(sW, a.wpack + n.wpack_off + (STREAM ? img_l0_off(F, NH) : 0), bytes, &bar_w);
  }
  const uint32_t tm = tmem_base_s;
  const long long total = a.coords ? a.n_coords : n.n_vox;
  const long long n_tiles = (total + kTile - 1) / kTile;
  const long long tile_begin = chunk * a.tiles_per_block;
  const int count = (int)(min(n_tiles, tile_begin + a.tiles_per_block) - tile_begin);  // tiles of this CTA
  const int rounds = (count + U - 1) / U;  // unit u serves tiles tile_begin + u, + U, ...
  mbar_wait(&bar_w, 0);
  __syncthreads();

  if (producer_warp) {
    // hidden weights in stream order: round after round W_1 .. W_NH into buffer (h & 1)
    if (elect_one()) {
      const int total_h = rounds * NH;
      for (int h = 0; h < total_h; ++h) {
        const int b = h & 1;
        if (h >= 2) mbar_wait(&bar_wfree[b], (uint32_t)((h >> 1) - 1) & 1);
        mbar_expect_tx(&bar_wfull[b], (uint32_t)F * F * 2);
        bulk_g2s(sWbuf + (size_t)b * F * F * 2, a.wpack + n.wpack_off + (size_t)(h % NH) * F * F * 2, (uint32_t)F * F * 2, &bar_wfull[b]);
      }
    }
    __syncwarp();
  } else if (mma_warp) {
    const uint32_t aW = smem_u32(sW), aU0 = smem_u32(sUnits);
    const uint32_t aL0 = aW + (STREAM ? 0u : (uint32_t)img_l0_off(F, NH));
    const uint32_t aLast = aL0 + (uint32_t)img_l0_bytes(F);
    const uint32_t aWb = smem_u32(sWbuf);
    const int steps = NH + 2;  // MMA batches per tile
    uint32_t ph = 0;           // every active unit's barrier flips once per step: one shared parity
    constexpr uint32_t idesc_last = make_idesc(128, 16, false, false);
    constexpr uint32_t idesc_f = make_idesc(128, F, false, false);
    for (int it = 0; it < rounds; ++it)
      for (int st = 0; st < steps; ++st) {
        const int hb = (it * NH + st - 1) & 1;  // STREAM: buffer of this hidden step's weights
        if (STREAM && st >= 1 && st <= NH) {
          mbar_wait(&bar_wfull[hb], (uint32_t)((it * NH + st - 1) >> 1) & 1);
          tc_fence_after();
        }
        for (int slot = 0; slot < S; ++slot)
          for (int gi = 0; gi < G; ++gi) {
            const int u = gi * S + slot;
            if (kEvalIssuers > 1 && (gi % kEvalIssuers) != issuer) continue;  // the other issue warp's unit
            if (it * U + u >= count) continue;  // this unit has no tile in the last round
            mbar_wait(&bar_r[u], ph);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t act = aU0 + (uint32_t)u * (uint32_t)eval_unit_bytes(F);
              const uint32_t d = tm + (uint32_t)u * F;
              if (st == 0) {  // theta_0 = A0 [128 x 16] * B0^T
                mma_f16(d, make_desc(act + kTile * F * 2, kActLBO, 128), make_desc(aL0, (F / 8) * 128, 128), idesc_f, 0);
              } else if (st <= NH) {
                issue_forward<F>(d, act, STREAM ? aWb + (uint32_t)hb * F * F * 2 : aW + (uint32_t)(st - 1) * F * F * 2);
              } else {  // y = a_NH * [hi(Wlast); lo(Wlast)]^T  (N = 16)
#pragma unroll
                for (int k = 0; k < F / 16; ++k)
                  mma_f16(d, make_desc(act + k * 2 * kActLBO, kActLBO, 128), make_desc(aLast + k * 2 * 256, 256, 128), idesc_last, k > 0);
              }
              commit(&bar_mma[u]);
            }
            __syncwarp();
          }
        if (STREAM && st >= 1 && st <= NH) {  // every contraction of this layer has been issued: its buffer is free once they complete
          if (elect_one()) commit(&bar_wfree[hb]);
          __syncwarp();
        }
        ph ^= 1;
      }
  } else {
    const float inv_w0 = 1.0f / n.w0, inv_wh = 1.0f / n.wh;
    uint32_t phase = 0;  // both slots wait the same number of times per round
    for (int it = 0; it < rounds; ++it) {
      const int u0 = g * S;
      bool act[kEvalSlots], valid[kEvalSlots];
      long long sidx[kEvalSlots];
#pragma unroll
      for (int slot = 0; slot < kEvalSlots; ++slot) {
        const int u = u0 + slot;
        act[slot] = slot < S && it * U + u < count;
        sidx[slot] = (tile_begin + (long long)it * U + u) * kTile + r;
        valid[slot] = act[slot] && sidx[slot] < total;
      }
      auto signal = [&](int u) {  // this warp's operand rows are written and its TMEM reads are done
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_r[u]);
      };
      // ---- coordinates -> layer-0 operand row [x_hi(3) 1 x_lo(3) 1 | x_hi(3) 0 0 0 0 0]
#pragma unroll
      for (int slot = 0; slot < kEvalSlots; ++slot) {
        if (!act[slot]) continue;
        const int u = u0 + slot;
        unsigned char* sX0 = sUnits + (size_t)u * eval_unit_bytes(F) + (size_t)kTile * F * 2;
        const long long s = sidx[slot];
        float x0 = 0.f, x1 = 0.f, x2 = 0.f;
        if (valid[slot]) {
          if (a.coords) {
            x0 = a.coords[s * n.in_dim];
            x1 = a.coords[s * n.in_dim + 1];
            x2 = n.in_dim == 3 ? a.coords[s * n.in_dim + 2] : 0.f;
          } else {
            brief_coords(n, a.axes, s, x0, x1, x2);
          }
        }
        const float h0 = __half2float(__float2half_rn(x0)), h1 = __half2float(__float2half_rn(x1)),
                    h2 = __half2float(__float2half_rn(x2));
        const uint32_t p01 = pack_f16x2(h0, h1), p21 = pack_f16x2(h2, 1.0f);
        *reinterpret_cast<uint4*>(sX0 + chunk_off(r, 0, kTile)) =
            make_uint4(p01, p21, pack_f16x2(x0 - h0, x1 - h1), pack_f16x2(x2 - h2, 1.0f));
        *reinterpret_cast<uint4*>(sX0 + chunk_off(r, 1, kTile)) = make_uint4(p01, pack_f16x2(h2, 0.f), 0u, 0u);
        signal(u);
      }
      // ---- sine layers: stage st reads theta_st from TMEM and writes a_st as the next MMA's operand
      for (int st = 0; st <= NH; ++st) {
#pragma unroll
        for (int slot = 0; slot < kEvalSlots; ++slot) {
          if (!act[slot]) continue;
          const int u = u0 + slot;
          unsigned char* sAct = sUnits + (size_t)u * eval_unit_bytes(F);
          const uint32_t my_tmem = tm + ((uint32_t)(32 * q) << 16) + (uint32_t)u * F;
          mbar_wait(&bar_mma[u], phase);
          tc_fence_after();
          float v[2][16];
          tmem_ld16(my_tmem, v[0]);
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            tmem_ld_wait();
            if (c + 1 < NC) tmem_ld16(my_tmem + 16 * (c + 1), v[(c + 1) & 1]);  // next chunk in flight under these sines
            float* vc = v[c & 1];
            if (DUMP && valid[slot]) {
              float* zdump = a.layers_out + (long long)st * total * n.f + sidx[slot] * n.f;
              const float inv = st == 0 ? inv_w0 : inv_wh;
              for (int i = 0; i < 16; ++i)
                if (16 * c + i < n.f) zdump[16 * c + i] = vc[i] * inv;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) vc[i] = fast_sin(vc[i]);
            store_chunk16(sAct, r, c, vc);
          }
          signal(u);
        }
        phase ^= 1;
      }
      // ---- last layer: y = acc[0] + acc[1] (hi + lo rows of Wlast, bias inside), output epilogue
#pragma unroll
      for (int slot = 0; slot < kEvalSlots; ++slot) {
        if (!act[slot]) continue;
        const int u = u0 + slot;
        const uint32_t my_tmem = tm + ((uint32_t)(32 * q) << 16) + (uint32_t)u * F;
        mbar_wait(&bar_mma[u], phase);
        tc_fence_after();
        uint32_t y0, y1;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(y0), "=r"(y1) : "r"(my_tmem) : "memory");
        tmem_ld_wait();
        const float y = __uint_as_float(y0) + __uint_as_float(y1);
        const long long s = sidx[slot];
        const bool full = (s - r) + kTile <= total;
        if (a.out_f32) {
          if (valid[slot]) a.out_f32[s] = y;
        } else {
          void* dst = a.out_ptrs[net_id];
          if (a.out_dtype == 2) {
            if (valid[slot]) reinterpret_cast<float*>(dst)[s] = y;
          } else {  // inverse normalisation + truncating cast
            const float vden = brief_denorm(n, y);
            if (a.out_dtype == 1) {
              const uint32_t uv = (uint32_t)(unsigned short)(int)vden;
              if (full) {  // 8 voxels per 16-byte store: lanes 0, 8, 16, 24 write the warp's 64 contiguous bytes
                const uint32_t w = uv | (__shfl_down_sync(0xffffffffu, uv, 1) << 16);
                const uint32_t w1 = __shfl_down_sync(0xffffffffu, w, 2), w2 = __shfl_down_sync(0xffffffffu, w, 4),
                               w3 = __shfl_down_sync(0xffffffffu, w, 6);
                if ((lane & 7) == 0)
                  *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(dst) + s) = make_uint4(w, w1, w2, w3);
              } else if (valid[slot]) {
                reinterpret_cast<unsigned short*>(dst)[s] = (unsigned short)uv;
              }
            } else if (valid[slot]) {
              reinterpret_cast<unsigned char*>(dst)[s] = (unsigned char)(int)vden;
            }
          }
        }
      }
      phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, tcols);
}

// ==================================================================================================================
// fit: gather + forward + weighted L2 + backward -> per-slice gradient partials
// ==================================================================================================================
// TWO tiles are in flight per CTA, in opposite phases, and each has its OWN warps: group B walks tile k+1 FORWARD
// (layer-0 operand row, sines, last layer, loss, dz_NH) while group A walks tile k BACKWARD (dz_{l-1} = dX * cos).
// The two chains only meet in the tensor pipe: a dedicated MMA-issue warp polls both groups' "operands written"
// mbarriers and issues whichever batch is ready (backward first: it is the longer chain), so one tile's MMA latency,
// TMEM loads, fences and stores sit under the other tile's special-function work instead of being serialised in one
// thread (a tcgen05.mma blocks its issuing thread for about as long as it executes: ~62 cycles per 128x64x16,
// tests/cuda/umma_timing_probe.cu).  A fourth role, the sampler warp, stages the next tile's samples.
//
// Every sine layer's argument comes out of an MMA (bias and omega inside, brief_image.cuh): layer 0 is one K = 16 MMA
// against the hi/lo-split coordinate row, which doubles as the B operand of the dW0 contraction.
//
// Both tiles' activations fit because their lifetimes are complementary: backward stage l of tile k still needs
// a_0..a_{l-1} while the forward pass of tile k+1 has produced a_0..a_j.  Even tiles keep layer j in ring slot j, odd
// tiles in slot R-1-j (R = NS + 2 slots); the MMA warp issues forward batch b of tile k+1 only after backward batch
// b-2 of tile k (whose trailing dW is the last reader of the slot that a_b overwrites) — the pipe executes in order.
// Hand-overs between the groups go through tcgen05.commit mbarriers: bar_f1 (tile k's dW_1 done: the dz buffer that
// tile k+1's dz_NH goes to is free) and bar_f2 (tile k complete: its ring slots and coordinate rows are free).
//
// shared memory (dynamic):  sDz[2] | ring[R] | sX[2] | sDY | image          (NS = L-1 sine layers, R = NS + 2)
//   sDz first: the M = 64 dW contractions read 8 feature groups (16 KB) from the start of a buffer whatever F is.
// TMEM columns: Zf [0,F) forward theta | Zb [F,2F) recomputed theta | Xb [2F,3F) dX | accumulators from 3F:
//   accumulator i (M = 64: 16 lanes per quadrant) lives in column block i/2, lane half i%2 — two per block.
//   i = 0..NH-1: dW_{i+1} (F columns; column f = db), i = NH: dW0 block (16 columns), i = NH+1: dWlast block.
// Optional: Zb double-buffered (when the columns allow it) so that the recomputed theta of the NEXT backward stage is
// issued behind the current stage's dW instead of in front of its dX.  Measured neutral on B200 (178 vs 174 us at
// F = 64): the tensor pipe is operand-bandwidth-bound here (~100 B/clk of shared-memory operand fetch), so moving four
// MMAs off the critical path does not shorten it.  Kept switchable for wider tiles.
constexpr bool kFitPrefetchTheta = false;
__host__ __device__ constexpr int fit_acc_blocks(int NH) { return (NH + 2 + 1) / 2; }
__host__ __device__ constexpr bool fit_zb_double(int F, int NH) {
  return kFitPrefetchTheta && 4 * F + fit_acc_blocks(NH) * F <= 512;
}
// A operands of the forward and dX contractions in TMEM (F/2 columns each) when the columns allow it
__host__ __device__ constexpr bool fit_ts_mode(int F, int NH) {
  return (fit_zb_double(F, NH) ? 5 : 4) * F + fit_acc_blocks(NH) * F <= 512;
}
__host__ __device__ constexpr int fit_fixed_cols(int F, int NH) {
  return (3 + (fit_zb_double(F, NH) ? 1 : 0) + (fit_ts_mode(F, NH) ? 1 : 0)) * F;
}
__host__ __device__ constexpr int fit_tmem_cols(int F, int NH) {
  return tmem_cols_pow2(fit_fixed_cols(F, NH) + fit_acc_blocks(NH) * F);
}

#ifndef BRIEF_FIT_CPT_B64
#define BRIEF_FIT_CPT_B64 2
#endif
#ifndef BRIEF_FIT_CPT_48
#define BRIEF_FIT_CPT_48 1
#endif
#ifndef BRIEF_FIT_CPT_A64
#define BRIEF_FIT_CPT_A64 2
#endif
#ifndef BRIEF_FIT_TWO_ISSUERS
#define BRIEF_FIT_TWO_ISSUERS 1
#endif
#ifndef BRIEF_FIT_DZ_IN_A
#define BRIEF_FIT_DZ_IN_A 1
#endif
#ifndef BRIEF_FIT_TWO_SAMPLERS
#define BRIEF_FIT_TWO_SAMPLERS 1
#endif
#ifndef BRIEF_FIT_TWO_ISSUERS_MIN_F
#define BRIEF_FIT_TWO_ISSUERS_MIN_F 32
#endif
template <int F>
struct FitCfg {
  static constexpr int NC = F / 16;                       // 16-column chunks per row
  // chunks per thread of the backward group (A) and of the forward group (B).  Measured at F = 64 (back to back):
  // A2/B2 (8 + 8 warps) 165 us, A2/B1 (8 + 16 warps) 173 us, A1/B2 (16 + 8) 175 us — more warps on one role lengthen
  // the other role's epilogues (shared XU / issue slots) by more than they shorten its own
  static constexpr int CPT_A = F == 64 ? BRIEF_FIT_CPT_A64 : F == 48 ? BRIEF_FIT_CPT_48 : F == 32 ? 2 : 1;
  static constexpr int CPT_B = F == 64 ? BRIEF_FIT_CPT_B64 : CPT_A;
  static constexpr int CG_A = NC / CPT_A, CG_B = NC / CPT_B;  // column groups per role
  static constexpr int GW_A = 4 * CG_A, GW_B = 4 * CG_B;      // warps per role
  // Two MMA-issue warps (one per chain) where one CTA owns the SM: a forward batch then shares the tensor pipe with a
  // backward batch in flight instead of queueing behind all of its contractions (the forward chain is the longer one)
  static constexpr bool TWO_ISSUERS = BRIEF_FIT_TWO_ISSUERS && F >= BRIEF_FIT_TWO_ISSUERS_MIN_F;
  // ... and two sampler warps (two rows per lane each instead of four): the index -> voxel chain of a tile is half as long
  static constexpr int SAMPLERS = (BRIEF_FIT_TWO_SAMPLERS && TWO_ISSUERS) ? 2 : 1;
  // dz_NH = (w_h dy') Wlast cos(theta_NH) formed by the BACKWARD group (whose chain is the shorter one) instead of the
  // forward group's loss phase; only with two issuers (the single-issuer schedule keeps the loss hand-over it was tuned with)
  static constexpr bool DZ_IN_A = BRIEF_FIT_DZ_IN_A && TWO_ISSUERS;
  static constexpr int THREADS = (GW_A + GW_B) * 32 + 64 + (TWO_ISSUERS ? 32 : 0) + (SAMPLERS - 1) * 32;  // + issue + sampler warps
  static constexpr int MIN_BLOCKS = F >= 48 ? 1 : F == 32 ? 2 : 3;  // must match tc_fit_ctas_per_sm()
};

template <int F>
__global__ void __launch_bounds__(FitCfg<F>::THREADS, FitCfg<F>::MIN_BLOCKS) tc_fit_kernel(FitArgs a) {
  using C = FitCfg<F>;
  constexpr int GW_A = C::GW_A, GW_B = C::GW_B, CG_B = C::CG_B, NC = C::NC;
  constexpr int NW = GW_A + GW_B;  // epilogue warps: [0, GW_B) group B (forward), [GW_B, NW) group A (backward)
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ NetDev sn;
  __shared__ __align__(8) uint64_t bar_w, bar_a, bar_b, bar_ra, bar_rb, bar_lb, bar_zf, bar_f1, bar_f2[2], bar_gfull[2], bar_gfree[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float4 s_g[2][kTile];  // sampler staging: (x0, x1, x2, normalised target) per row
  __shared__ float s_gw[2][kTile];                //                  loss weight per row
  __shared__ float s_y[CG_B][kTile];
  __shared__ float s_dys[C::DZ_IN_A ? 2 : 1][C::DZ_IN_A ? kTile : 1];  // w_h dy' per row: loss phase (group B) -> dz_NH stage (group A)
  __shared__ float s_red[4];

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const bool mma_warp = warp == NW, fwd_issue_warp = C::TWO_ISSUERS && warp == NW + 2;
  const bool sampler_warp = warp == NW + 1 || (C::SAMPLERS == 2 && warp == NW + 3);
  __shared__ volatile int s_bwd_progress;  // two issuers: (tile, batches issued) of the backward chain, for the ring rule
  const bool group_a = warp >= GW_B && warp < NW;
  const int gw = group_a ? warp - GW_B : warp;  // warp index inside the role
  const int CPT = group_a ? C::CPT_A : C::CPT_B, CG = group_a ? C::CG_A : C::CG_B;
  const int q = warp & 3, cg = (gw >> 2) % CG, r = 32 * q + lane;
  const int c_base = cg * CPT;                // first 16-column chunk of this thread
#ifdef BRIEF_TC_TIMING
  const int tslot = warp == 0 ? 0 : warp == GW_B ? 16 : warp == NW ? 32 : warp == NW + 1 ? 48 : -1;
#endif
  TT(k_start);
  const int wi = tc_find_work(a.work_prefix, a.n_work, blockIdx.x);
  const int net_id = a.work_net[wi];
  const int slice = blockIdx.x - a.work_prefix[wi];
  tc_load_net(sn, a.nets[net_id]);
  if (t == 0) {
    s_bwd_progress = 0;
    mbar_init(&bar_w, 1);
    mbar_init(&bar_a, 1);
    mbar_init(&bar_b, 1);
    mbar_init(&bar_f1, 1);
    mbar_init(&bar_f2[0], 1);  // "tile complete", one barrier per tile parity: the waiter (group B, two tiles later) can
    mbar_init(&bar_f2[1], 1);  // never see two completions of the same barrier before it has consumed the first
    mbar_init(&bar_ra, GW_A);  // "operands of the backward tile are in place": one arrival per warp of group A
    mbar_init(&bar_rb, GW_B);  // same for the forward tile / group B
    // "loss phase done" has its own barrier: group B raises it and the NEXT tile's first bar_rb signal back to back,
    // and a parity wait cannot tell a phase from the one two completions later
    mbar_init(&bar_lb, GW_B);
    mbar_init(&bar_zf, GW_A);  // DZ_IN_A: group A has read theta_NH out of Zf (the next tile's layer-0 contraction may overwrite it)
    for (int k = 0; k < 2; ++k) {
      mbar_init(&bar_gfull[k], C::SAMPLERS);  // sampler warp(s): staging slot k holds a tile's samples
      mbar_init(&bar_gfree[k], GW_B);   // group B: staging slot k has been consumed
    }
    fence_mbar_init();
  }
  __syncthreads();
  const NetDev& n = sn;
  const int NH = n.L - 2, NS = n.L - 1, f = n.f, R = NS + 2;
  const int tcols = fit_tmem_cols(F, NH);
  if (warp == 0) tmem_alloc(&tmem_base_s, tcols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  constexpr uint32_t BUF = kTile * F * 2;    // one [128 x F] fp16 buffer
  constexpr uint32_t BLK = kTile * 16 * 2;   // one [128 x 16] fp16 block
  unsigned char* sDz = smem;
  unsigned char* sRing = sDz + 2 * BUF;
  unsigned char* sX = sRing + (size_t)R * BUF;
  unsigned char* sDY = sX + 2 * BLK;
  unsigned char* sW = sDY + BLK;
  const float* side = reinterpret_cast<const float*>(sW + img_side_off(F, NH));
  const float* s_wl = side + 4 * F + NH * F;
  const float* s_bl = s_wl + F;
  if (t == 0) {
    const uint32_t bytes = (uint32_t)img_bytes(F, NH);
    mbar_expect_tx(&bar_w, bytes);
    bulk_g2s(sW, a.wpack + n.wpack_off, bytes, &bar_w);
  }
  if (!group_a && warp < GW_B && cg == 0)  // second column group of the dY block is constant zero
    *reinterpret_cast<uint4*>(sDY + chunk_off(r, 1, kTile)) = make_uint4(0, 0, 0, 0);

  const uint32_t tm = tmem_base_s;
  const uint32_t my_tmem = tm + ((uint32_t)(32 * q) << 16) + 16 * c_base;
  const bool zb2 = fit_zb_double(F, NH);           // theta of backward stage l lives in Zb[l & 1] (Zb[0] if single)
  const uint32_t zb_stride = zb2 ? F : 0;
  const uint32_t acc0 = (uint32_t)fit_fixed_cols(F, NH);
  const bool ts = fit_ts_mode(F, NH);              // A operands of forward / dX contractions live in TMEM
  const uint32_t TZF = tm, TZB = tm + F, TXB = tm + 2 * F + zb_stride;
  const uint32_t TAF = TXB + F, TAD = TAF + F / 2;  // fp16 a_{j-1} of the forward tile | dz_l of the backward tile
  const uint32_t my_af = TAF + ((uint32_t)(32 * q) << 16) + 8 * c_base, my_ad = TAD + ((uint32_t)(32 * q) << 16) + 8 * c_base;
  auto acc_addr = [&](int i) { return tm + acc0 + (uint32_t)(i >> 1) * F + ((uint32_t)(i & 1) << 20); };  // lane 16 = 16 << 16
  const uint32_t aDz = smem_u32(sDz), aRing = smem_u32(sRing), aX = smem_u32(sX), aDY = smem_u32(sDY), aW = smem_u32(sW);
  auto slot = [&](int parity, int j) { return (uint32_t)(parity ? R - 1 - j : j) * BUF; };  // byte offset into the ring
  const float wh = n.wh, w0 = n.w0;
  const long long s_begin = (long long)slice * n.slice_len;
  const long long s_end = min((long long)n.batch, s_begin + n.slice_len);
  const int n_tiles = (int)((s_end - s_begin + kTile - 1) / kTile);
  const int sb_flip = (NH - 1) & 1;  // dz_NH of tile k+1 goes where dz_1 of tile k lived: start buffer sb(k+1) = sb(k) ^ sb_flip
  float loss_acc = 0.f;

  // epilogue warp -> MMA warp: this warp's operand rows are written (and its TMEM reads are done)
  auto signal = [&](uint64_t* bar) {
    if (ts) tmem_st_wait();
    tc_fence_before();
    fence_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
  };

  mbar_wait(&bar_w, 0);
  { TT(k_roles); TACC(9, k_roles - k_start); }
  if (mma_warp) {
    // ============================================ MMA-issue warp ===============================================
    // B events per tile, in order: NH+1 forward batches (0 = layer 0), then the "loss done" signal that arms the
    // backward chain.  A batches per tile: NH gated ones (l = NH..1: recomputed theta_{l-1}, dX_{l-1} | commit |
    // dW_l) and the final dW0.
    constexpr uint32_t idesc_f = make_idesc(128, F, false, false);
    uint32_t ph_ra = 0, ph_rb = 0, ph_lb = 0;
    int kA = 0, ia = 0, kB = 0, ib = 0, loss_done = 0, sbA = 0;
    auto layer0 = [&](uint32_t d, int parity) {  // theta_0 = A0 [128 x 16] * B0^T
      mma_f16(d, make_desc(aX + parity * BLK, kActLBO, 128), make_desc(aW + (uint32_t)img_l0_off(F, NH), (F / 8) * 128, 128),
              idesc_f, 0);
    };
    // backward batch ia of tile kA (the caller has seen its trigger)
    auto issue_backward = [&]() {
      tc_fence_after();
      const int pa = kA & 1;
      const bool accum = kA > 0;
      if (elect_one()) {
        if (ia < NH) {
          const int l = NH - ia;
          const uint32_t dzb = aDz + (uint32_t)(sbA ^ (ia & 1)) * BUF;  // dz_l
          auto recompute = [&](int m) {  // theta_m -> Zb[(m + 1) & 1], read by backward stage m + 1
            const uint32_t d = TZB + (uint32_t)((m + 1) & 1) * zb_stride;
            if (m >= 1) issue_forward<F>(d, aRing + slot(pa, m - 1), aW + (uint32_t)(m - 1) * F * F * 2);
            else layer0(d, pa);
          };
          if (!zb2 || l == NH) recompute(l - 1);
          if (ts) issue_dx_ts<F>(TXB, TAD, aW + (uint32_t)(l - 1) * F * F * 2);
          else issue_dx<F>(TXB, dzb, aW + (uint32_t)(l - 1) * F * F * 2);
          commit(&bar_a);
          issue_dw<F>(acc_addr(l - 1), dzb, aRing + slot(pa, l - 1), accum);
          if (l == NH) issue_dw<16>(acc_addr(NH + 1), aRing + slot(pa, NH), aDY, accum);
          if (l == 1) commit(&bar_f1);  // dW_1 done: the buffer of dz_1 may take the next tile's dz_NH
          if (zb2 && l >= 2) recompute(l - 2);  // next stage's theta, behind this stage's dW
        } else {  // dW0 += dz_0^T [x_hi, 1, x_lo, ...]; completes the tile
          issue_dw<16>(acc_addr(NH), aDz + (uint32_t)(sbA ^ (NH & 1)) * BUF, aX + pa * BLK, accum);
          commit(&bar_f2[pa]);
        }
      }
      __syncwarp();
      if (++ia > NH) { ia = 0; ++kA; sbA ^= sb_flip; }
    };
    TT(m_start);
    if (C::TWO_ISSUERS) {
      // this warp issues the backward chain only; the forward chain has its own issue warp (below)
      while (kA < n_tiles) {
        if (ia == 0 && !C::DZ_IN_A) { mbar_wait(&bar_lb, ph_lb); ph_lb ^= 1; }  // loss done: dz_NH in place
        else { mbar_wait(&bar_ra, ph_ra); ph_ra ^= 1; }                         // group A wrote dz_l (DZ_IN_A: also dz_NH)
        TT(m0);
        issue_backward();
        if (lane == 0) s_bwd_progress = kA * 32 + ia;  // published AFTER the batch has been issued
        { TT(m1); TACC(0, m1 - m0); TACC(1, 1); }
      }
    } else
    while (kA < n_tiles) {
      bool progressed = false;
      TT(m0);
      // ---- backward chain of tile kA
      bool a_ready = false;
      if (ia == 0) {
        a_ready = loss_done > kA;
      } else if (mbar_test(&bar_ra, ph_ra)) {
        ph_ra ^= 1;
        a_ready = true;
      }
      if (a_ready) {
        issue_backward();
        progressed = true;
        { TT(m1); TACC(0, m1 - m0); TACC(1, 1); }
      }
      TT(m2);
      // ---- forward chain of tile kB (ring safety: batch b only after backward batch b-2 of tile kB-1 was issued)
      if (kB < n_tiles) {
        const bool allowed = ib > NH || ib < 2 || kA >= kB || (kA == kB - 1 && ia >= ib - 1);
        if (allowed && (ib <= NH ? mbar_test(&bar_rb, ph_rb) : mbar_test(&bar_lb, ph_lb))) {
          if (ib <= NH) ph_rb ^= 1; else ph_lb ^= 1;
          tc_fence_after();
          if (ib <= NH) {
            if (elect_one()) {
              const int pb = kB & 1;
              if (ib == 0) layer0(TZF, pb);
              else if (ts) issue_forward_ts<F>(TZF, TAF, aW + (uint32_t)(ib - 1) * F * F * 2);
              else issue_forward<F>(TZF, aRing + slot(pb, ib - 1), aW + (uint32_t)(ib - 1) * F * F * 2);
              commit(&bar_b);
            }
            __syncwarp();
            ++ib;
          } else {  // loss phase done: dz_NH, sDY in place
            loss_done = kB + 1;
            ib = 0;
            ++kB;
          }
          progressed = true;
          { TT(m3); TACC(2, m3 - m2); TACC(3, 1); }
        }
      }
      if (!progressed) __nanosleep(20);
      if (!progressed) { TT(m4); TACC(4, m4 - m0); }
    }
    { TT(m_end); TACC(5, m_end - m_start); TACC(6, n_tiles); }
  } else if (fwd_issue_warp) {
    // ============================================ forward-chain issue warp (two issuers) =========================
    constexpr uint32_t idesc_f = make_idesc(128, F, false, false);
    uint32_t ph_rb = 0;
    for (int kB = 0; kB < n_tiles; ++kB) {
      const int pb = kB & 1;
      for (int ib = 0; ib <= NH; ++ib) {
        // ring safety: batch b only after backward batch b-2 of tile kB-1 has been issued (its trailing dW is the last
        // reader of the slot that a_b overwrites; the rule keeps one whole backward batch of slack behind that reader)
        if (ib >= 2 && kB >= 1) {
          for (;;) {
            const int pr = s_bwd_progress, pk = pr >> 5, pi = pr & 31;
            if (pk >= kB || (pk == kB - 1 && pi >= ib - 1)) break;
            __nanosleep(20);
          }
        }
        mbar_wait(&bar_rb, ph_rb);
        ph_rb ^= 1;
        if (C::DZ_IN_A && ib == 0 && kB >= 1) mbar_wait(&bar_zf, (uint32_t)(kB - 1) & 1);  // theta_NH of tile kB-1 consumed
        tc_fence_after();
        if (elect_one()) {
          if (ib == 0)
            mma_f16(TZF, make_desc(aX + pb * BLK, kActLBO, 128), make_desc(aW + (uint32_t)img_l0_off(F, NH), (F / 8) * 128, 128),
                    idesc_f, 0);
          else if (ts) issue_forward_ts<F>(TZF, TAF, aW + (uint32_t)(ib - 1) * F * F * 2);
          else issue_forward<F>(TZF, aRing + slot(pb, ib - 1), aW + (uint32_t)(ib - 1) * F * F * 2);
          commit(&bar_b);
        }
        __syncwarp();
      }
    }
  } else if (sampler_warp) {
    // ============================================ sampler warp ==================================================
    // main.py:126-163 / whole-block cube for one tile, up to two tiles ahead of its use: index -> coordinates (axis
    // tables), raw voxel -> normalised target, loss weight; 4 rows per lane
    constexpr int SR = 4 / C::SAMPLERS;                 // rows per lane of this warp
    const int h0 = (warp == NW + 1) ? 0 : SR;             // first 32-row group of this warp
    for (int k = 0; k < n_tiles; ++k) {
      const int sl = k & 1;
      if (k >= 2) mbar_wait(&bar_gfree[sl], (uint32_t)((k >> 1) - 1) & 1);
      // three passes so that the rows' dependent loads (index -> voxel, index -> axis tables) are all in flight
      // together instead of one row's chain after the other (rows past the slice end read voxel 0 and are zeroed)
      long long idx[SR];
      bool ok[SR];
#pragma unroll
      for (int h = 0; h < SR; ++h) {
        const long long s = s_begin + (long long)k * kTile + lane + 32 * (h0 + h);
        ok[h] = s < s_end;
        long long v = 0;
        if (ok[h]) {
          if (n.mode == 0) v = s;
          else if (a.idx) v = a.idx[n.idx_off + s];
          else v = brief_sample_index(a.seed, a.state ? a.state->step : a.step, n.stream_id, (uint64_t)s, (uint64_t)n.n_vox);
        }
        idx[h] = v;
      }
      float raw[SR], cx[SR][3];
#pragma unroll
      for (int h = 0; h < SR; ++h) {
        raw[h] = brief_raw_value(n, idx[h]);
        brief_coords(n, a.axes, idx[h], cx[h][0], cx[h][1], cx[h][2]);
      }
#pragma unroll
      for (int h = 0; h < SR; ++h) {
        const int row = lane + 32 * (h0 + h);
        const float yv = ok[h] ? brief_normalize(n, raw[h]) : 0.f;
        const float wv = ok[h] ? brief_weight(n, idx[h], raw[h]) : 0.f;
        s_g[sl][row] = ok[h] ? make_float4(cx[h][0], cx[h][1], cx[h][2], yv) : make_float4(0.f, 0.f, 0.f, 0.f);
        s_gw[sl][row] = wv;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_gfull[sl]);
    }
  } else if (!group_a) {
    // ============================================ group B: forward ==============================================
    uint32_t ph_b = 0;
    int sb = 0;
    TT(b_start);
    for (int k = 0; k < n_tiles; ++k) {
      const int p = k & 1;
      TT(b0);
      if (k >= 2) mbar_wait(&bar_f2[p], (uint32_t)((k >> 1) - 1) & 1);  // tile k-2 complete: ring slots of this parity and sX[p] are free
      TT(b1);
      mbar_wait(&bar_gfull[p], (uint32_t)(k >> 1) & 1);
      TT(b2);
      TACC(0, b1 - b0); TACC(1, b2 - b1);
      const float4 xf = s_g[p][r];
      if (cg == 0) {  // layer-0 operand row [x_hi(3) 1 x_lo(3) 1 | x_hi(3) 0 0 0 0 0]; also the B operand of dW0
        const float h0 = __half2float(__float2half_rn(xf.x)), h1 = __half2float(__float2half_rn(xf.y)),
                    h2 = __half2float(__float2half_rn(xf.z));
        const uint32_t p01 = pack_f16x2(h0, h1), p21 = pack_f16x2(h2, 1.0f);
        *reinterpret_cast<uint4*>(sX + p * BLK + chunk_off(r, 0, kTile)) =
            make_uint4(p01, p21, pack_f16x2(xf.x - h0, xf.y - h1), pack_f16x2(xf.z - h2, 1.0f));
        *reinterpret_cast<uint4*>(sX + p * BLK + chunk_off(r, 1, kTile)) = make_uint4(p01, pack_f16x2(h2, 0.f), 0u, 0u);
      }
      signal(&bar_rb);
      float ypart = 0.f;
      for (int st = 0; st <= NH; ++st) {  // stage st: theta_st (Zf) -> a_st
        TT(b3);
        mbar_wait(&bar_b, ph_b);
        ph_b ^= 1;
        tc_fence_after();
        TT(b4);
        TACC(2, b4 - b3);
        unsigned char* dst = sRing + slot(p, st);
        float v[2][16];
        tmem_ld16(my_tmem, v[0]);
#pragma unroll
        for (int c = 0; c < C::CPT_B; ++c) {
          tmem_ld_wait();
          if (c + 1 < C::CPT_B) tmem_ld16(my_tmem + 16 * (c + 1), v[(c + 1) & 1]);
          float* vc = v[c & 1];
#pragma unroll
          for (int i = 0; i < 16; ++i) vc[i] = fast_sin(vc[i]);
          store_chunk16_both<false>(dst, r, c_base + c, vc, ts && st < NH, my_af + 8 * c);
          if (st == NH) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 w4 = *reinterpret_cast<const float4*>(s_wl + 16 * (c_base + c) + i);
              ypart = fmaf(w4.x, vc[i], ypart); ypart = fmaf(w4.y, vc[i + 1], ypart);
              ypart = fmaf(w4.z, vc[i + 2], ypart); ypart = fmaf(w4.w, vc[i + 3], ypart);
            }
          }
        }
        if (st < NH) signal(&bar_rb);
        { TT(b5); TACC(3, b5 - b4); }
      }
      TT(b6);
      // ---- loss (datal2, main.py:176-182), scaled output gradient and dz_NH
      if (CG_B > 1) {
        s_y[cg][r] = ypart;
        named_bar_sync(1, GW_B * 32);
      }
      // every thread of the row forms y, the error and dy itself (same operands, same order: identical values);
      // column group 0 additionally accumulates the loss and writes the dWlast operand block
      float y = s_bl[0];
      if (CG_B > 1) {
#pragma unroll
        for (int c = 0; c < CG_B; ++c) y += s_y[c][r];
      } else {
        y += ypart;
      }
      float dys = 0.f;
      if (s_begin + (long long)k * kTile + r < s_end) {
        const float e = y - xf.w;  // xf.w: the row's normalised target
        const float wt = (n.tau != 0.f && y <= n.tau) ? 1.0f : s_gw[p][r];
        if (cg == 0) loss_acc = fmaf(wt * e, e, loss_acc);
        dys = kGradScale * wt * e;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_gfree[p]);  // staging slot consumed (target and weight are in registers)
      TT(b7);
      if (k >= 1) mbar_wait(&bar_f1, (uint32_t)(k - 1) & 1);  // dW_1 of tile k-1 done: dz buffer sb is free
      TT(b8);
      TACC(4, b7 - b6); TACC(5, b8 - b7);
      if (cg == 0) *reinterpret_cast<uint4*>(sDY + chunk_off(r, 0, kTile)) = make_uint4(pack_f16x2_sat(dys, 0.f), 0, 0, 0);
      dys *= wh;
      if (C::DZ_IN_A) {
        if (cg == 0) s_dys[p][r] = dys;  // group A forms dz_NH from this and theta_NH (still in Zf)
      } else {
        unsigned char* dzb = sDz + (size_t)sb * BUF;
        float v[2][16];
        tmem_ld16(my_tmem, v[0]);  // theta_NH is still in Zf: the next forward MMA is issued after this signal
#pragma unroll
        for (int c = 0; c < C::CPT_B; ++c) {
          tmem_ld_wait();
          if (c + 1 < C::CPT_B) tmem_ld16(my_tmem + 16 * (c + 1), v[(c + 1) & 1]);
          float* vc = v[c & 1];
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(s_wl + 16 * (c_base + c) + i);
            vc[i] = dys * w4.x * fast_cos(vc[i]); vc[i + 1] = dys * w4.y * fast_cos(vc[i + 1]);
            vc[i + 2] = dys * w4.z * fast_cos(vc[i + 2]); vc[i + 3] = dys * w4.w * fast_cos(vc[i + 3]);
          }
          store_chunk16_both<true>(dzb, r, c_base + c, vc, ts, my_ad + 8 * c);
        }
      }
      signal(&bar_lb);  // the "loss done" event
      sb ^= sb_flip;
      { TT(b9); TACC(6, b9 - b8); }
    }
    { TT(b_end); TACC(7, b_end - b_start); TACC(8, n_tiles); }
  } else {
    // ============================================ group A: backward =============================================
    uint32_t ph_a = 0;
    int sb = 0;
    const float w0_over_wh = w0 / wh;
    TT(a_start);
    for (int k = 0; k < n_tiles; ++k) {
      if (C::DZ_IN_A) {  // dz_NH = (w_h dy') Wlast cos(theta_NH): theta_NH is still in Zf, dy' comes from group B's loss phase
        TT(a0);
        mbar_wait(&bar_lb, (uint32_t)k & 1);
        if (k >= 1) mbar_wait(&bar_f1, (uint32_t)(k - 1) & 1);  // dW_1 of tile k-1 done: dz buffer sb is free
        tc_fence_after();
        TT(a1);
        TACC(0, a1 - a0);
        unsigned char* dzb = sDz + (size_t)sb * BUF;
        const float dys = s_dys[k & 1][r];
        float v[C::CPT_A][16];
#pragma unroll
        for (int c = 0; c < C::CPT_A; ++c) tmem_ld16(my_tmem + 16 * c, v[c]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_zf);  // Zf may be overwritten by the next tile's layer-0 contraction
#pragma unroll
        for (int c = 0; c < C::CPT_A; ++c) {
          float* vc = v[c];
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(s_wl + 16 * (c_base + c) + i);
            vc[i] = dys * w4.x * fast_cos(vc[i]); vc[i + 1] = dys * w4.y * fast_cos(vc[i + 1]);
            vc[i + 2] = dys * w4.z * fast_cos(vc[i + 2]); vc[i + 3] = dys * w4.w * fast_cos(vc[i + 3]);
          }
          store_chunk16_both<true>(dzb, r, c_base + c, vc, ts, my_ad + 8 * c);
        }
        signal(&bar_ra);
        { TT(a2); TACC(4, a2 - a1); }
      }
      for (int ia = 0; ia < NH; ++ia) {  // stage l = NH - ia: dz_{l-1} = dX_{l-1} * w * cos(theta_{l-1})
        const int l = NH - ia;
        TT(a0);
        mbar_wait(&bar_a, ph_a);
        ph_a ^= 1;
        tc_fence_after();
        TT(a1);
        if (ia == 0 && !C::DZ_IN_A) TACC(0, a1 - a0); else TACC(1, a1 - a0);
        unsigned char* dzb = sDz + (size_t)(sb ^ ((ia + 1) & 1)) * BUF;  // dz_{l-1}
        const float scale = l >= 2 ? 1.0f : w0_over_wh;  // dX carries w_hidden (omega-scaled weights); layer 0 wants w_0
        float vz[2][16], vx[2][16];
        const uint32_t tz = my_tmem + F + (uint32_t)(l & 1) * zb_stride, tx = my_tmem + 2 * F + zb_stride;
        tmem_ld16(tz, vz[0]);
        tmem_ld16(tx, vx[0]);
#pragma unroll
        for (int c = 0; c < C::CPT_A; ++c) {
          tmem_ld_wait();
          if (c + 1 < C::CPT_A) {
            tmem_ld16(tz + 16 * (c + 1), vz[(c + 1) & 1]);
            tmem_ld16(tx + 16 * (c + 1), vx[(c + 1) & 1]);
          }
          float* z = vz[c & 1];
          const float* x = vx[c & 1];
#pragma unroll
          for (int i = 0; i < 16; ++i) z[i] = x[i] * scale * fast_cos(z[i]);
          store_chunk16_both<true>(dzb, r, c_base + c, z, ts && l >= 2, my_ad + 8 * c);
        }
        TT(a2);
        signal(&bar_ra);
        TT(a3);
        TACC(2, a2 - a1); TACC(3, a3 - a2);
      }
      sb ^= sb_flip;
    }
    { TT(a_end); TACC(7, a_end - a_start); TACC(8, n_tiles); }
  }
  TT(k_sync0);
  __syncthreads();
  TT(k_sync1);
  TACC(10, k_sync1 - k_sync0);

  // ---- slice epilogue: loss partial + gradient partials (TMEM -> global), scale removed in fp32
  const float inv_count = 1.0f / ((float)n.batch * (float)n.out_dim);
  if (warp < GW_B && cg == 0) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) loss_acc += __shfl_down_sync(0xffffffffu, loss_acc, off);
    if (lane == 0) s_red[q] = loss_acc;
  }
  // The slot image (padded device layout, pads zero) is assembled in shared memory — the operand buffers are dead
  // by now — and leaves with coalesced 16-byte stores; scattered 4-byte stores from the accumulator rows cost ~10 us.
  float* img = reinterpret_cast<float*>(smem);
  for (int i = t; i < n.P_dev; i += C::THREADS) img[i] = 0.f;
  __syncthreads();
  if (t == 0) a.loss_partials[n.slice_off + slice] = (((s_red[0] + s_red[1]) + s_red[2]) + s_red[3]) * inv_count;
  if (n_tiles > 0 && warp < 4 * NC) {  // warp w: lane quadrant w & 3, 16-column chunk w >> 2 of every accumulator block
    const float unscale = 2.0f * inv_count / kGradScale;
    // lanes 0..15 of a quadrant hold the even accumulator of a block, lanes 16..31 the odd one; row = 16q + lane%16
    const int ch = warp >> 2;
    const int o = q * 16 + (lane & 15), half = lane >> 4;
    const int F4 = n.F4;
    const int n_blocks = fit_acc_blocks(NH);
    const uint32_t base = tm + ((uint32_t)(32 * q) << 16) + 16 * ch + acc0;
    float v[16];
    for (int b = 0; b < n_blocks; ++b) {
      const int i = 2 * b + half;  // accumulator index served by this thread in block b
      tmem_ld16(base + b * F, v);
      tmem_ld_wait();
      if (i < NH) {  // dW_{i+1}, db_{i+1}
        if (o < f) {
          const int l = i + 1;
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const int k = 16 * ch + c;
            if (k < f) img[dl_W(n, l) + o * F4 + k] = v[c] * unscale;
            else if (k == f) img[dl_b(n, l) + o] = v[c] * unscale;
          }
        }
      } else if (i == NH) {  // dW0 block
        if (ch == 0 && o < f) {
          img[dl_W0(n) + 4 * o + 0] = (v[0] + v[4]) * unscale;
          img[dl_W0(n) + 4 * o + 1] = (v[1] + v[5]) * unscale;
          if (n.in_dim == 3) img[dl_W0(n) + 4 * o + 2] = (v[2] + v[6]) * unscale;
          img[dl_b0(n) + o] = v[3] * unscale;
        }
      } else if (i == NH + 1) {  // dWlast block: row = feature of a_NH, column 0; row f is dblast
        if (ch == 0) {
          if (o < f) img[dl_Wlast(n) + o] = v[0] * unscale;
          else if (o == f) img[dl_blast(n)] = v[0] * unscale;
        }
      }
    }
  }
  __syncthreads();
  {
    float4* dst = reinterpret_cast<float4*>(a.partials + n.part_off + (long long)slice * n.P_dev);
    const float4* src = reinterpret_cast<const float4*>(img);
    for (int i = t; i < (n.P_dev >> 2); i += C::THREADS) dst[i] = src[i];
  }
  tc_fence_before();
  __syncthreads();
  { TT(k_end); TACC(11, k_end - k_sync1); TACC(12, k_end - k_start); }
  if (warp == 0) tmem_dealloc(tm, tcols);
}

// ==================================================================================================================
// host side
// ==================================================================================================================
int tc_fpad(int f) { return ((f + 2 + 15) / 16) * 16; }  // two constant-one columns (bias hi / lo)

size_t tc_wpack_bytes(int F, int L) { return img_bytes(F, L - 2); }
// groups per CTA and tile slots per group: TMEM columns (G S F <= 512) and shared memory (image + G S unit buffers)
static void eval_shape(int F, int L, int* groups, int* slots) {
  const size_t img = eval_weight_bytes(F, L - 2);
  const size_t budget = (F <= 32 ? (size_t)110 : (size_t)224) * 1024;  // F <= 32: two CTAs per SM
  *groups = 0;
  *slots = 0;
  if (img + eval_unit_bytes(F) > budget) return;
  int units = (int)((budget - img) / eval_unit_bytes(F));
  units = units < 512 / F ? units : 512 / F;
  if (units >= 2) {
    *slots = kEvalSlots;
    *groups = units / kEvalSlots < kEvalMaxGroups ? units / kEvalSlots : kEvalMaxGroups;
  } else {
    *slots = 1;
    *groups = 1;
  }
}
int tc_eval_groups(int F, int L) { int g, s; eval_shape(F, L, &g, &s); return g; }
int tc_eval_slots(int F, int L) { int g, s; eval_shape(F, L, &g, &s); return s; }
size_t tc_eval_smem(int F, int L) {
  return eval_weight_bytes(F, L - 2) + (size_t)tc_eval_groups(F, L) * tc_eval_slots(F, L) * eval_unit_bytes(F);
}
// forward / decompress on the tensor core: any width whose padded image + one tile fit (f <= 126 at L = 7)
bool tc_eval_supported(int f, int L, int in_dim, int out_dim) {
  const int F = tc_fpad(f);
  if (out_dim != 1 || (in_dim != 2 && in_dim != 3) || L < 3 || F > 128) return false;
  return tc_eval_groups(F, L) >= 1;
}
size_t tc_fit_smem(int F, int L) {  // sDz[2] + ring[NS + 2] + sX[2] + sDY + image
  return (size_t)(2 + (L - 1) + 2) * kTile * F * 2 + 3 * (size_t)kTile * 16 * 2 + img_bytes_padded(F, L - 2);
}

// resident fit CTAs per SM: the launch bound (registers), shared memory (dynamic + ~8 KB static) and TMEM columns
int tc_fit_ctas_per_sm(int F, int L) {
  if (F > 64) return 1;  // wide kernel: the whole shared memory and all 512 TMEM columns
  const int by_bound = F >= 48 ? 1 : F == 32 ? 2 : 3;  // == FitCfg<F>::MIN_BLOCKS
  const size_t dyn = tc_fit_smem(F, L) < 49152 ? 49152 : tc_fit_smem(F, L);
  const int by_smem = (int)((size_t)227 * 1024 / (dyn + 8192));
  const int by_tmem = 512 / fit_tmem_cols(F, L - 2);
  int r = by_bound < by_smem ? by_bound : by_smem;
  r = r < by_tmem ? r : by_tmem;
  return r < 1 ? 1 : r;
}

bool tc_supported(int f, int L, int in_dim, int out_dim) {
  const int F = tc_fpad(f);
  if (out_dim != 1 || (in_dim != 2 && in_dim != 3)) return false;
  if (F > 128) return tc_lw_supported(f, L, in_dim, out_dim);  // layer-wise path (brief_tc_lw.cu)
  if (F > 64) return tc_wide_supported(f, L, in_dim, out_dim);
  if (L < 3 || F > 64) return false;
  if (3 * F + fit_acc_blocks(L - 2) * F > 512) return false;  // TMEM: Zf, Zb (x2 if it fits), Xb + packed dW accumulators
  if (tc_fit_smem(F, L) > 221 * 1024) return false;           // + ~5 KB static shared memory <= 227 KB
  if (tc_eval_groups(F, L) < 1) return false;
  return true;
}

template <int F>
static cudaError_t launch_eval_f(const EvalArgs& a_in, int L_max, int n_blocks, cudaStream_t st) {
  const int G = tc_eval_groups(F, L_max);
  if (G < 1) return cudaErrorInvalidValue;
  const size_t smem = tc_eval_smem(F, L_max);
  EvalArgs a = a_in;
  a.eval_slots = tc_eval_slots(F, L_max);
  cudaError_t e;
  if (a.layers_out) {
    e = cudaFuncSetAttribute(tc_eval_kernel<F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    tc_eval_kernel<F, true><<<n_blocks, G * 128 + 32 * eval_issuers(F) + (eval_streams(F) ? 32 : 0), smem, st>>>(a);
  } else {
    e = cudaFuncSetAttribute(tc_eval_kernel<F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    tc_eval_kernel<F, false><<<n_blocks, G * 128 + 32 * eval_issuers(F) + (eval_streams(F) ? 32 : 0), smem, st>>>(a);
  }
  return cudaGetLastError();
}

cudaError_t launch_tc_eval(const EvalArgs& a, int F_PAD, int L_max, int n_blocks, cudaStream_t st) {
  switch (F_PAD) {
    case 16: return launch_eval_f<16>(a, L_max, n_blocks, st);
    case 32: return launch_eval_f<32>(a, L_max, n_blocks, st);
    case 48: return launch_eval_f<48>(a, L_max, n_blocks, st);
    case 64: return launch_eval_f<64>(a, L_max, n_blocks, st);
    case 80: return launch_eval_f<80>(a, L_max, n_blocks, st);
    case 96: return launch_eval_f<96>(a, L_max, n_blocks, st);
    case 112: return launch_eval_f<112>(a, L_max, n_blocks, st);
    case 128: return launch_eval_f<128>(a, L_max, n_blocks, st);
    default: return cudaErrorInvalidValue;
  }
}

template <int F>
static cudaError_t launch_fit_f(const FitArgs& a, int L_max, int n_blocks, cudaStream_t st) {
  // at least 48 KB so that the 16 KB the M = 64 contractions read from a dz buffer always exist
  const size_t smem = tc_fit_smem(F, L_max) < 49152 ? 49152 : tc_fit_smem(F, L_max);
  cudaError_t e = cudaFuncSetAttribute(tc_fit_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  tc_fit_kernel<F><<<n_blocks, FitCfg<F>::THREADS, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_tc_fit(const FitArgs& a, int F_PAD, int L_max, int n_blocks, cudaStream_t st) {
  switch (F_PAD) {
    case 80: case 96: case 112: case 128: return launch_tc_fit_wide(a, F_PAD, L_max, n_blocks, st);
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
  cudaMemcpyFromSymbol(out, brief::g_tc_timing, sizeof(unsigned long long) * 64);
  if (reset) { unsigned long long z[64] = {0}; cudaMemcpyToSymbol(brief::g_tc_timing, z, sizeof z); }
  return 0;
}
#endif

