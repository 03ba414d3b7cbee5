// brief_common.cuh — device-side data model shared by every kernel of libbrief_b200.
//
// Layout in HBM (one BriefGroup = many independent per-block SIRENs, main.py:484-532):
//   params / grads / opt-state : one fp32 arena per group; network n owns P_dev(n) floats at
//                                 param_off(n), in the PADDED device layout below (feature dims
//                                 rounded up to 4 so every row is 16-byte aligned):
//        W0 [F4][4]   b0 [F4]   { Wl [F4][F4]  bl [F4] } x (L-2)   Wlast [F4]  blast [4]
//     (Wlast is the single output row; data_channel == 1.)  Pad entries are zero and stay zero
//     under every optimiser (zero gradient).  The reference's packed order (utils/ModelSave.py)
//     exists only at the ABI boundary (set/get_params).
//   axis tables                : per network dims[0]+dims[1]+dims[2] floats (create_coords,
//                                 utils/dataset.py:11-35) — coordinates are looked up per voxel,
//                                 never materialised as an [N,3] tensor.
//   raw volume                 : caller-owned, original dtype (u8/u16/f32), normalised on chip.
//   partials                   : per (network, slice) one P_dev-float gradient slot + one loss
//                                 slot; the optimiser kernel reduces slots in fixed order, so a
//                                 network's result does not depend on what else shares the GPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define BRIEF_MAX_RULES 4
#define BRIEF_MAX_LAYERS 16

struct NetDev {
  // architecture
  int in_dim, out_dim, f, L;   // L = total linear layers
  int F4;                      // f rounded up to 4
  float w0, wh;
  int d, h, w;                 // block dims (2-D: d = 1)
  long long n_vox;
  // arenas
  long long param_off;         // floats, into params/grads/m/v arenas
  int P_dev;                   // padded parameter count
  int P_ref;                   // reference parameter count
  int axis_off;                // floats, into axis arena: [d | h | w]
  // volume binding
  const void* raw;
  const float* weight;         // optional explicit per-voxel weights
  int dtype;                   // BriefDType
  int bound;
  float vmin, vmax, lo, hi;    // normalisation: ((x-vmin)/(vmax-vmin))*(hi-lo)+lo
  float dn_vmin, dn_vmax, dn_lo, dn_hi;  // inverse normalisation (decompress)
  float dn_range;              // fp32(double(dn_vmax) - double(dn_vmin)), torch's scalar rounding
  int n_rules;
  float rule_lo[BRIEF_MAX_RULES], rule_hi[BRIEF_MAX_RULES], rule_s[BRIEF_MAX_RULES];
  float tau;
  // sampler
  int mode;                    // BriefSamplerMode
  int batch;                   // samples per step (== n_vox for FULL_BLOCK)
  long long idx_off;           // offset into the replay index array
  unsigned int stream_id;      // key of this network's on-device sampler stream (default: its index in the group)
  // work decomposition for fit
  int slice_len;               // samples per slice
  int n_slices;
  long long slice_off;         // first loss-partial slot of this network
  long long part_off;          // floats, into the gradient-partials arena (n_slices slots of P_dev)
  // tensor-core path
  int prec;                    // resolved BriefPrecision of the FIT path
  int eval_tc;                 // 1: forward / decompress run on the tensor core (an operand image exists)
  int F_PAD;                   // padded width of the fp16 operand image (multiple of 16, >= f + 2)
  long long wpack_off;         // bytes, into the fp16 packed-weight arena
};

// ---- padded device layout offsets (floats, relative to param_off) --------------------------
__host__ __device__ inline int dl_W0(const NetDev& n) { return 0; }
__host__ __device__ inline int dl_b0(const NetDev& n) { return 4 * n.F4; }
__host__ __device__ inline int dl_W(const NetDev& n, int l) {  // hidden layer l in [1, L-2]
  return 5 * n.F4 + (l - 1) * (n.F4 * n.F4 + n.F4);
}
__host__ __device__ inline int dl_b(const NetDev& n, int l) { return dl_W(n, l) + n.F4 * n.F4; }
__host__ __device__ inline int dl_Wlast(const NetDev& n) { return 5 * n.F4 + (n.L - 2) * (n.F4 * n.F4 + n.F4); }
__host__ __device__ inline int dl_blast(const NetDev& n) { return dl_Wlast(n) + n.F4; }
__host__ __device__ inline int dl_total(int F4, int L) { return 5 * F4 + (L - 2) * (F4 * F4 + F4) + F4 + 4; }

// ---- Philox4x32-10 (on-device sampler; restated in oracle/brief_oracle.py) --------------------
struct Philox4 { uint32_t x, y, z, w; };
__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
  return o;
}

// Index of sample s of network `net` at step `step` (with replacement, like torch.randint).
__host__ __device__ inline long long brief_sample_index(uint64_t seed, uint64_t step, uint32_t net,
                                                        uint64_t s, uint64_t pop) {
  Philox4 r = philox4x32_10((uint32_t)(s >> 2), (uint32_t)step, (uint32_t)(step >> 32), net,
                            (uint32_t)seed, (uint32_t)(seed >> 32));
  uint32_t word = (s & 3) == 0 ? r.x : (s & 3) == 1 ? r.y : (s & 3) == 2 ? r.z : r.w;
  return (long long)(((uint64_t)word * pop) >> 32);
}

// General RandomCubeSampler (main.py:61-69, 112-116): the population is every position of a ch x cw (x cd) window slid
// over the block with stride 1, listed '(dc hc wc)'; sample o of cube `cube` in 'ds hs ws' order is this voxel.
__host__ __device__ inline long long brief_cube_voxel(int h, int w, int ch, int cw, long long cube, long long o) {
  const long long HC = h - ch + 1, WC = w - cw + 1;
  const long long wc = cube % WC, r = cube / WC, hc = r % HC, dc = r / HC;
  const long long ws = o % cw, q = o / cw, hs = q % ch, ds = q / ch;
  return ((dc + ds) * h + (hc + hs)) * w + (wc + ws);
}

#ifdef __CUDACC__
// ---- per-sample input fetch ------------------------------------------------------------------
__device__ __forceinline__ float brief_raw_value(const NetDev& n, long long idx) {
  if (n.dtype == 1) return (float)__ldg((const unsigned short*)n.raw + idx);
  if (n.dtype == 0) return (float)__ldg((const unsigned char*)n.raw + idx);
  return __ldg((const float*)n.raw + idx);
}
// normalize_data, utils/io.py:74-77: three separately rounded fp32 ops after an IEEE division.
__device__ __forceinline__ float brief_normalize(const NetDev& n, float raw) {
  float t = __fdiv_rn(__fsub_rn(raw, n.vmin), __fsub_rn(n.vmax, n.vmin));
  t = __fmul_rn(t, __fsub_rn(n.hi, n.lo));
  return __fadd_rn(t, n.lo);
}
// parse_weight 'value' rules, utils/misc.py:293-297, later rules override earlier ones.
__device__ __forceinline__ float brief_weight(const NetDev& n, long long idx, float raw) {
  if (n.weight) return __ldg(n.weight + idx);
  float wv = 1.0f;
#pragma unroll
  for (int r = 0; r < BRIEF_MAX_RULES; ++r)
    if (r < n.n_rules && raw >= n.rule_lo[r] && raw <= n.rule_hi[r]) wv = n.rule_s[r];
  return wv;
}
// voxel index -> coordinates through the per-axis tables; channel order (d,h,w) / (h,w).
__device__ __forceinline__ void brief_coords(const NetDev& n, const float* __restrict__ axes, long long idx,
                                             float& c0, float& c1, float& c2) {
  const float* ax = axes + n.axis_off;
  int x, y, z;
  if (n.n_vox <= 0x7fffffffLL) {  // 32-bit index arithmetic (a 64-bit div/mod pair costs ~10x as many instructions)
    const unsigned i = (unsigned)idx, w = (unsigned)n.w, h = (unsigned)n.h;
    const unsigned r = i / w;
    x = (int)(i - r * w);
    z = (int)(r / h);
    y = (int)(r - (unsigned)z * h);
  } else {
    x = (int)(idx % n.w);
    const long long r = idx / n.w;
    y = (int)(r % n.h);
    z = (int)(r / n.h);
  }
  if (n.in_dim == 3) {
    c0 = __ldg(ax + z);
    c1 = __ldg(ax + n.d + y);
    c2 = __ldg(ax + n.d + n.h + x);
  } else {
    c0 = __ldg(ax + n.d + y);
    c1 = __ldg(ax + n.d + n.h + x);
    c2 = 0.0f;
  }
}
// invnormalize_data, utils/io.py:136-147 + truncating cast (np.array(tensor, dtype)).
__device__ __forceinline__ float brief_denorm(const NetDev& n, float y) {
  float t = __fsub_rn(y, n.dn_lo);
  t = __fdiv_rn(t, __fsub_rn(n.dn_hi, n.dn_lo));
  t = fminf(fmaxf(t, 0.0f), 1.0f);
  // torch computes (max-min) in double from the yaml floats, then rounds the scalar to fp32 (dn_range)
  t = __fmul_rn(t, n.dn_range);
  return __fadd_rn(t, n.dn_vmin);
}
#endif  // __CUDACC__
