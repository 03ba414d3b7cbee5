// brief_preprocess.cu — the reference's `preprocess` (utils/misc.py:244-254) on the device, in place:
//
//     data[ binary_opening(data <= level, structure = ones(sz, sy, sx)) ] = 0 ;   data = clip(data, lo, hi)
//
// called on every block before the fit (main.py:336) and on every decoded block (main.py:295).  The reference runs
// scipy.ndimage on the host: a threshold pass, an erosion, a dilation and a fancy-index store, each a full numpy pass
// over the block.  Here the mask lives as ONE BIT per voxel (32 voxels of an x-row per word), so the morphology is a
// handful of word ANDs / ORs / funnel shifts per 32 voxels, and the only volume-sized traffic is
//   pass 1  mask_*_kernel read the block once (sizeof(T) B/voxel), write 1 bit/voxel;
//   pass 2  open_kernel   bitmask -> opened bitmask (1/16 of the block, L2-resident; skipped for a 1x1x1 structure);
//   pass 3  apply_kernel  write zeros where the opened bit is set; the block is READ again only when the clip range
//                         is not the whole dtype range (then every voxel is read, zeroed / clipped and written).
// Algorithmic bytes per voxel: sizeof(T) read + sizeof(T) written for the voxels that change.  All three are HBM-bound
// byte kernels with three row paths: flat (W % 32 == 0: no index arithmetic, warp-coalesced words), vector
// (W % (16/sizeof(T)) == 0: 16-byte accesses per row) and scalar (any W).
//
// Opening with a box structure (what np.ones(close) is): erosion E[q] = AND of the mask over the box anchored at q
// (boxes leaving the block count as 0 — scipy's border_value = 0), opening O[p] = OR of E over the boxes containing p.
// The box anchor is immaterial for an opening; tests pin this against scipy for every structure size.
#include "brief_kernels.h"

namespace brief {

constexpr int kPreThreads = 256;

// Threshold of VPT voxels held in one 16-byte vector -> VPT mask bits.
template <typename T>
__device__ __forceinline__ unsigned int vec_threshold(const uint4& q, unsigned int thr) {
  const unsigned int w[4] = {q.x, q.y, q.z, q.w};
  unsigned int bits = 0;
  if (sizeof(T) == 2) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      bits |= ((w[i] & 0xffffu) <= thr ? 1u : 0u) << (2 * i);
      bits |= ((w[i] >> 16) <= thr ? 1u : 0u) << (2 * i + 1);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) bits |= (((w[i] >> (8 * j)) & 0xffu) <= thr ? 1u : 0u) << (4 * i + j);
  }
  return bits;
}

// ---- pass 1: threshold -> bitmask -------------------------------------------------------------------------------------
// Mask layout: [D][H][WW] 32-bit words, WW = ceil(W / 32); bit b of word wx <-> x = 32 wx + b; bits with x >= W are 0.
template <typename T>
__global__ void __launch_bounds__(kPreThreads) mask_vec_kernel(const T* __restrict__ vol, long long rows, int W, int WW,
                                                                unsigned int thr, unsigned char* __restrict__ mask) {
  // one thread = one 16-byte vector = VPT voxels = VPT mask bits (1 byte for uint16, 2 bytes for uint8)
  constexpr int VPT = 16 / sizeof(T);
  const int vec_per_row = W / VPT;
  const long long total = rows * vec_per_row;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += stride) {
    const long long row = v / vec_per_row;
    const int iv = (int)(v - row * vec_per_row);
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(vol + row * W) + iv);
    const unsigned int bits = vec_threshold<T>(q, thr);
    if (sizeof(T) == 2) mask[row * WW * 4 + iv] = (unsigned char)bits;
    else reinterpret_cast<unsigned short*>(mask + row * WW * 4)[iv] = (unsigned short)bits;
  }
}

// Rows that are whole mask words (W % 32 == 0, the usual block shapes): volume and mask are both flat, no index
// arithmetic at all.  One warp = 32 mask words = 1024 voxels per iteration: every lane has LPW independent 16-byte loads
// in flight, the VPT-bit pieces are assembled into words by butterfly shuffles and stored as one coalesced 128-byte line.
template <typename T>
__global__ void __launch_bounds__(kPreThreads) mask_flat_kernel(const uint4* __restrict__ vol, long long nwords,
                                                                 unsigned int thr, unsigned int* __restrict__ mask) {
  constexpr int VPT = 16 / sizeof(T), LPW = 32 / VPT, PER = 32 / LPW;  // lanes per word (4 / 2), words per 32 vectors
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long w0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; w0 < nwords; w0 += nwarps * 32) {
    uint4 q[LPW];
    bool ok[LPW];
#pragma unroll
    for (int j = 0; j < LPW; ++j) {
      const int lv = 32 * j + lane;  // vector within the warp's 32 words
      ok[j] = w0 + lv / LPW < nwords;
      if (ok[j]) q[j] = __ldg(vol + w0 * LPW + lv);
    }
    unsigned int r = 0;
#pragma unroll
    for (int j = 0; j < LPW; ++j) {
      unsigned int x = ok[j] ? vec_threshold<T>(q[j], thr) << (VPT * (lane % LPW)) : 0u;
      x |= __shfl_xor_sync(0xffffffffu, x, 1);
      if (LPW == 4) x |= __shfl_xor_sync(0xffffffffu, x, 2);
      // every lane of a group now holds word PER * j + lane / LPW; lane k takes word k
      const unsigned int t = __shfl_sync(0xffffffffu, x, (lane % PER) * LPW);
      if (lane / PER == j) r = t;
    }
    if (w0 + lane < nwords) mask[w0 + lane] = r;
  }
}

template <typename T>
__global__ void __launch_bounds__(kPreThreads) mask_scalar_kernel(const T* __restrict__ vol, long long rows, int W, int WW,
                                                                   unsigned int thr, unsigned int* __restrict__ mask) {
  // one warp = one mask word per iteration
  const int lane = threadIdx.x & 31;
  const long long total = rows * WW;
  const long long stride = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long wi = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); wi < total; wi += stride) {
    const long long row = wi / WW;
    const int x = (int)(wi - row * WW) * 32 + lane;
    const bool in = x < W && (unsigned int)vol[row * W + x] <= thr;
    const unsigned int word = __ballot_sync(0xffffffffu, in);
    if (lane == 0) mask[wi] = word;
  }
}

// ---- pass 2: opening on the bitmask ---------------------------------------------------------------------------------------
// One thread = one mask word column (z, wx) over `yr` consecutive rows.  Per row it x-erodes the words of the 2 SZ - 1
// planes around z (64-bit window: bit 32 + b <-> x = 32 wx + b, bit b <-> x = 32 (wx - 1) + b), ANDs them over z for the SZ
// box anchors, and slides the result through a register window of 2 SY - 1 rows: each output word costs (2 SZ - 1) x 3
// word loads instead of (2 SZ - 1)(2 SY - 1) x 3.  Plane pointers and all the z / x border predicates are loop-invariant.
template <int SZ, int SY, int SX>
__global__ void __launch_bounds__(kPreThreads) open_kernel(const unsigned int* __restrict__ m, int D, int H, int WW, int yr,
                                                           unsigned int* __restrict__ out) {
  constexpr int NP = 2 * SZ - 1, NR = 2 * SY - 1;
  const unsigned int ygroups = (unsigned int)((H + yr - 1) / yr);
  const unsigned int total = (unsigned int)D * ygroups * (unsigned int)WW;
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int wx = (int)(i % (unsigned int)WW);
    const unsigned int t = i / (unsigned int)WW;
    const int y0 = (int)(t % ygroups) * yr, z = (int)(t / ygroups);
    const bool has_p = wx > 0, has_n = wx + 1 < WW;
    // word (z', 0, wx) of the planes z - (SZ-1) .. z + (SZ-1).  Planes, rows and neighbour words outside the block are
    // read from a clamped (valid) address and masked to 0 afterwards: every load of a row is unconditional, so all
    // 3 (2 SZ - 1) of them are in flight together instead of one guarded group after the other.
    const unsigned int* plane[NP];
    unsigned long long zmask[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const int zz = z - (SZ - 1) + k;
      plane[k] = m + (size_t)min(max(zz, 0), D - 1) * H * WW + wx;
      zmask[k] = (zz >= 0 && zz < D) ? ~0ull : 0ull;
    }
    const int dp = has_p ? -1 : 0, dn = has_n ? 1 : 0;
    const unsigned int pm = has_p ? ~0u : 0u, nm = has_n ? ~0u : 0u;
    // R[a] of row y: AND over the SZ planes of box anchor a of the x-eroded words; 0 for rows outside the block
    auto load_row = [&](int y, unsigned long long (&R)[SZ]) {
      unsigned long long X[NP];
      const unsigned long long ymask = (y >= 0 && y < H) ? ~0ull : 0ull;
      const int off = min(max(y, 0), H - 1) * WW;
      unsigned int c[NP], pw[NP], nw[NP];
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        const unsigned int* r = plane[k] + off;
        c[k] = __ldg(r);
        pw[k] = __ldg(r + dp);
        nw[k] = __ldg(r + dn);
      }
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        const unsigned int n = nw[k] & nm;
        const unsigned long long lo = (unsigned long long)(pw[k] & pm) | ((unsigned long long)c[k] << 32);
        unsigned long long e = lo;
#pragma unroll
        for (int s = 1; s < SX; ++s) e &= (lo >> s) | ((unsigned long long)n << (64 - s));
        X[k] = e & zmask[k] & ymask;
      }
#pragma unroll
      for (int a = 0; a < SZ; ++a) {
        R[a] = X[a];
#pragma unroll
        for (int s = 1; s < SZ; ++s) R[a] &= X[a + s];
      }
    };
    unsigned long long R[NR][SZ];  // window rows y - (SY-1) .. y + (SY-1)
#pragma unroll
    for (int k = 1; k < NR; ++k) load_row(y0 - SY + k, R[k]);
    const int y_end = min(y0 + yr, H);
    unsigned int* o = out + ((size_t)z * H + y0) * WW + wx;
    for (int y = y0; y < y_end; ++y, o += WW) {
#pragma unroll
      for (int k = 0; k < NR - 1; ++k)
#pragma unroll
        for (int a = 0; a < SZ; ++a) R[k][a] = R[k + 1][a];
      load_row(y + SY - 1, R[NR - 1]);
      unsigned long long T = 0ull;
#pragma unroll
      for (int b = 0; b < SY; ++b)  // box anchor row y - (SY-1) + b covers window rows b .. b + SY - 1
#pragma unroll
        for (int a = 0; a < SZ; ++a) {
          unsigned long long e = R[b][a];
#pragma unroll
          for (int j = 1; j < SY; ++j) e &= R[b + j][a];
          T |= e;
        }
      unsigned long long O = T;  // x-dilation: O[x] = OR_s E[x - s]
#pragma unroll
      for (int s = 1; s < SX; ++s) O |= T << s;
      *o = (unsigned int)(O >> 32);
    }
  }
}

// ---- pass 3: apply ------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ unsigned int clip1(unsigned int v, unsigned int lo, unsigned int hi) {
  return min(max(v, lo), hi);
}

template <typename T, bool CLIP>
__device__ __forceinline__ void apply_vec(uint4* p, unsigned int bits, unsigned int lo, unsigned int hi) {
  constexpr int VPT = 16 / sizeof(T);
  if (!CLIP) {
    if (bits == 0) return;
    if (bits == (1u << VPT) - 1u) { *p = make_uint4(0, 0, 0, 0); return; }
  }
  const uint4 q = *p;
  unsigned int w[4] = {q.x, q.y, q.z, q.w};
  if (sizeof(T) == 2) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      unsigned int a = w[i] & 0xffffu, b = w[i] >> 16;
      if (bits & (1u << (2 * i))) a = 0;
      if (bits & (1u << (2 * i + 1))) b = 0;
      if (CLIP) { a = clip1<T>(a, lo, hi); b = clip1<T>(b, lo, hi); }
      w[i] = a | (b << 16);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      unsigned int r = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        unsigned int a = (w[i] >> (8 * j)) & 0xffu;
        if (bits & (1u << (4 * i + j))) a = 0;
        if (CLIP) a = clip1<T>(a, lo, hi);
        r |= a << (8 * j);
      }
      w[i] = r;
    }
  }
  *p = make_uint4(w[0], w[1], w[2], w[3]);
}

template <typename T, bool CLIP>
__global__ void __launch_bounds__(kPreThreads) apply_vec_kernel(T* __restrict__ vol, long long rows, int W, int WW,
                                                                 const unsigned char* __restrict__ mask, unsigned int lo,
                                                                 unsigned int hi) {
  constexpr int VPT = 16 / sizeof(T);
  const int vec_per_row = W / VPT;
  const long long total = rows * vec_per_row;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += stride) {
    const long long row = v / vec_per_row;
    const int iv = (int)(v - row * vec_per_row);
    unsigned int bits;
    if (sizeof(T) == 2) bits = mask[row * WW * 4 + iv];
    else bits = reinterpret_cast<const unsigned short*>(mask + row * WW * 4)[iv];
    apply_vec<T, CLIP>(reinterpret_cast<uint4*>(vol + row * W) + iv, bits, lo, hi);
  }
}

// flat twin of mask_flat_kernel: one coalesced 128-byte load of 32 mask words per warp, the VPT-bit pieces handed to the
// lanes by shuffle, LPW coalesced 512-byte stores.
template <typename T, bool CLIP>
__global__ void __launch_bounds__(kPreThreads) apply_flat_kernel(uint4* __restrict__ vol, long long nwords,
                                                                  const unsigned int* __restrict__ mask, unsigned int lo,
                                                                  unsigned int hi) {
  constexpr int VPT = 16 / sizeof(T), LPW = 32 / VPT;
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long w0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; w0 < nwords; w0 += nwarps * 32) {
    const unsigned int word = w0 + lane < nwords ? __ldg(mask + w0 + lane) : 0u;
    if (!CLIP && __ballot_sync(0xffffffffu, word != 0) == 0) continue;
#pragma unroll
    for (int j = 0; j < LPW; ++j) {
      const int lv = 32 * j + lane;
      const unsigned int bits = (__shfl_sync(0xffffffffu, word, lv / LPW) >> (VPT * (lv % LPW))) & ((1u << VPT) - 1u);
      if (w0 + lv / LPW < nwords) apply_vec<T, CLIP>(vol + w0 * LPW + lv, bits, lo, hi);
    }
  }
}

template <typename T, bool CLIP>
__global__ void __launch_bounds__(kPreThreads) apply_scalar_kernel(T* __restrict__ vol, long long rows, int W, int WW,
                                                                    const unsigned int* __restrict__ mask, unsigned int lo,
                                                                    unsigned int hi) {
  const int lane = threadIdx.x & 31;
  const long long total = rows * WW;
  const long long stride = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long wi = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); wi < total; wi += stride) {
    const unsigned int word = __ldg(mask + wi);
    if (!CLIP && word == 0) continue;
    const long long row = wi / WW;
    const int x = (int)(wi - row * WW) * 32 + lane;
    if (x >= W) continue;
    T* p = vol + row * W + x;
    const bool z = (word >> lane) & 1u;
    if (CLIP) {
      unsigned int a = z ? 0u : (unsigned int)*p;
      *p = (T)clip1<T>(a, lo, hi);
    } else if (z) {
      *p = (T)0;
    }
  }
}

// ---- host side ------------------------------------------------------------------------------------------------------------
size_t preprocess_scratch_bytes(int D, int H, int W) {
  const size_t words = (size_t)D * H * ((W + 31) / 32);
  return 2 * words * sizeof(unsigned int);  // threshold mask + opened mask
}

template <int SZ, int SY>
static bool open_dispatch_x(int sx, const unsigned int* m, int D, int H, int WW, int yr, unsigned int* out, int grid,
                            cudaStream_t st) {
  switch (sx) {
    case 1: open_kernel<SZ, SY, 1><<<grid, kPreThreads, 0, st>>>(m, D, H, WW, yr, out); return true;
    case 2: open_kernel<SZ, SY, 2><<<grid, kPreThreads, 0, st>>>(m, D, H, WW, yr, out); return true;
    case 3: open_kernel<SZ, SY, 3><<<grid, kPreThreads, 0, st>>>(m, D, H, WW, yr, out); return true;
    case 4: open_kernel<SZ, SY, 4><<<grid, kPreThreads, 0, st>>>(m, D, H, WW, yr, out); return true;
  }
  return false;
}
template <int SZ>
static bool open_dispatch_y(int sy, int sx, const unsigned int* m, int D, int H, int WW, int yr, unsigned int* out,
                            int grid, cudaStream_t st) {
  switch (sy) {
    case 1: return open_dispatch_x<SZ, 1>(sx, m, D, H, WW, yr, out, grid, st);
    case 2: return open_dispatch_x<SZ, 2>(sx, m, D, H, WW, yr, out, grid, st);
    case 3: return open_dispatch_x<SZ, 3>(sx, m, D, H, WW, yr, out, grid, st);
    case 4: return open_dispatch_x<SZ, 4>(sx, m, D, H, WW, yr, out, grid, st);
  }
  return false;
}

template <typename T>
static cudaError_t preprocess_t(T* vol, int D, int H, int W, unsigned int thr, bool any_mask, int sz, int sy, int sx,
                                unsigned int lo, unsigned int hi, bool clip, unsigned int* scratch, int num_sms,
                                int* launches, cudaStream_t st) {
  constexpr int VPT = 16 / sizeof(T);
  const int WW = (W + 31) / 32;
  const long long rows = (long long)D * H, words = rows * WW;
  const bool vec = (W % VPT == 0) && ((reinterpret_cast<uintptr_t>(vol) & 15) == 0);
  unsigned int* mask = scratch;
  unsigned int* opened = scratch + words;
  const int cap = num_sms * 8;  // 8 CTAs of 256 threads per SM: full occupancy, grid-stride over the rest
  auto grid_for = [&](long long items, int per_cta) { return (int)std::max<long long>(1, std::min<long long>(cap, (items + per_cta - 1) / per_cta)); };
  *launches = 0;
  const unsigned int* final_mask = nullptr;
  if (words >= (1ll << 32)) return cudaErrorInvalidValue;  // 2^37 voxels per block: far beyond any 180 GB device
  const bool flat = vec && (W % 32 == 0);
  if (any_mask) {
    if (flat) {
      mask_flat_kernel<T><<<grid_for(words, kPreThreads), kPreThreads, 0, st>>>(reinterpret_cast<const uint4*>(vol), words, thr,
                                                                               mask);
    } else if (vec) {
      cudaError_t e = cudaMemsetAsync(mask, 0, words * sizeof(unsigned int), st);  // row tails past W/VPT vectors stay 0
      if (e != cudaSuccess) return e;
      mask_vec_kernel<T><<<grid_for(rows * (W / VPT), kPreThreads), kPreThreads, 0, st>>>(
          vol, rows, W, WW, thr, reinterpret_cast<unsigned char*>(mask));
    } else {
      mask_scalar_kernel<T><<<grid_for(words, kPreThreads / 32), kPreThreads, 0, st>>>(vol, rows, W, WW, thr, mask);
    }
    ++*launches;
    final_mask = mask;
    if (sz * sy * sx > 1) {
      // rows per thread: 8 when that still leaves two full waves of threads, fewer for small blocks
      int yr = 8;
      while (yr > 1 && (long long)D * ((H + yr - 1) / yr) * WW < 2ll * num_sms * 2048) yr >>= 1;
      const int grid = grid_for((long long)D * ((H + yr - 1) / yr) * WW, kPreThreads);
      bool ok = false;
      switch (sz) {
        case 1: ok = open_dispatch_y<1>(sy, sx, mask, D, H, WW, yr, opened, grid, st); break;
        case 2: ok = open_dispatch_y<2>(sy, sx, mask, D, H, WW, yr, opened, grid, st); break;
        case 3: ok = open_dispatch_y<3>(sy, sx, mask, D, H, WW, yr, opened, grid, st); break;
        case 4: ok = open_dispatch_y<4>(sy, sx, mask, D, H, WW, yr, opened, grid, st); break;
      }
      if (!ok) return cudaErrorInvalidValue;
      ++*launches;
      final_mask = opened;
    }
  }
  if (!any_mask && !clip) return cudaGetLastError();
  if (!any_mask) {  // clip only: an all-zero mask
    cudaError_t e = cudaMemsetAsync(mask, 0, words * sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    final_mask = mask;
  }
  if (flat) {
    const int grid = grid_for(words, kPreThreads);
    uint4* v4 = reinterpret_cast<uint4*>(vol);
    if (clip) apply_flat_kernel<T, true><<<grid, kPreThreads, 0, st>>>(v4, words, final_mask, lo, hi);
    else apply_flat_kernel<T, false><<<grid, kPreThreads, 0, st>>>(v4, words, final_mask, lo, hi);
  } else if (vec) {
    const int grid = grid_for(rows * (W / VPT), kPreThreads);
    const unsigned char* mb = reinterpret_cast<const unsigned char*>(final_mask);
    if (clip) apply_vec_kernel<T, true><<<grid, kPreThreads, 0, st>>>(vol, rows, W, WW, mb, lo, hi);
    else apply_vec_kernel<T, false><<<grid, kPreThreads, 0, st>>>(vol, rows, W, WW, mb, lo, hi);
  } else {
    const int grid = grid_for(words, kPreThreads / 32);
    if (clip) apply_scalar_kernel<T, true><<<grid, kPreThreads, 0, st>>>(vol, rows, W, WW, final_mask, lo, hi);
    else apply_scalar_kernel<T, false><<<grid, kPreThreads, 0, st>>>(vol, rows, W, WW, final_mask, lo, hi);
  }
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_preprocess(void* vol, int dtype, int D, int H, int W, unsigned int thr, bool any_mask, int sz, int sy,
                              int sx, unsigned int lo, unsigned int hi, bool clip, void* scratch, int num_sms, int* launches,
                              cudaStream_t st) {
  if (dtype == 0)
    return preprocess_t<unsigned char>(reinterpret_cast<unsigned char*>(vol), D, H, W, thr, any_mask, sz, sy, sx, lo, hi, clip,
                                       reinterpret_cast<unsigned int*>(scratch), num_sms, launches, st);
  return preprocess_t<unsigned short>(reinterpret_cast<unsigned short*>(vol), D, H, W, thr, any_mask, sz, sy, sx, lo, hi, clip,
                                      reinterpret_cast<unsigned int*>(scratch), num_sms, launches, st);
}

}  // namespace brief
