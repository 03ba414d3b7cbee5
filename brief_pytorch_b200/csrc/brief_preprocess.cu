// brief_preprocess.cu — the reference's `preprocess` (utils/misc.py:244-254) on the device, in place:
//
//     data[ binary_opening(data <= level, structure = ones(sz, sy, sx)) ] = 0 ;   data = clip(data, lo, hi)
//
// called on every block before the fit (main.py:336) and on every decoded block (main.py:295).  The reference runs
// scipy.ndimage on the host: a threshold pass, an erosion, a dilation and a fancy-index store, each a full numpy pass
// over the block.  Here the mask lives as ONE BIT per voxel (32 voxels of an x-row per word), so the morphology is a
// handful of word ANDs / ORs / funnel shifts per 32 voxels, and the only volume-sized traffic is
//   pass 1  mask_kernel   read the block once (sizeof(T) B/voxel), write 1 bit/voxel;
//   pass 2  open_kernel   bitmask -> opened bitmask (1/16 of the block, L2-resident; skipped for a 1x1x1 structure);
//   pass 3  apply_kernel  write zeros where the opened bit is set; the block is READ again only when the clip range
//                         is not the whole dtype range (then every voxel is read, zeroed / clipped and written).
// Algorithmic bytes per voxel: sizeof(T) read + sizeof(T) written for the voxels that change.  All three are HBM-bound
// byte kernels: 16-byte vector accesses when the row length allows it (W % (16/sizeof(T)) == 0), scalar otherwise.
//
// Opening with a box structure (what np.ones(close) is): erosion E[q] = AND of the mask over the box anchored at q
// (boxes leaving the block count as 0 — scipy's border_value = 0), opening O[p] = OR of E over the boxes containing p.
// The box anchor is immaterial for an opening; tests pin this against scipy for every structure size.
#include "brief_kernels.h"

namespace brief {

constexpr int kPreThreads = 256;

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
    const unsigned int w[4] = {q.x, q.y, q.z, q.w};
    unsigned int bits = 0;
    if (sizeof(T) == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        bits |= ((w[i] & 0xffffu) <= thr ? 1u : 0u) << (2 * i);
        bits |= ((w[i] >> 16) <= thr ? 1u : 0u) << (2 * i + 1);
      }
      mask[row * WW * 4 + iv] = (unsigned char)bits;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) bits |= (((w[i] >> (8 * j)) & 0xffu) <= thr ? 1u : 0u) << (4 * i + j);
      reinterpret_cast<unsigned short*>(mask + row * WW * 4)[iv] = (unsigned short)bits;
    }
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
// x-erosion of one row around word wx: bit (32 + b) of the result <-> E_x at x = 32 wx + b, bit b <-> x = 32 (wx-1) + b.
template <int SX>
__device__ __forceinline__ unsigned long long row_erode_x(const unsigned int* __restrict__ m, int z, int y, int wx, int D,
                                                          int H, int WW) {
  if (z < 0 || z >= D || y < 0 || y >= H) return 0ull;
  const unsigned int* r = m + ((long long)z * H + y) * WW;
  const unsigned int p = wx > 0 ? __ldg(r + wx - 1) : 0u, c = __ldg(r + wx), n = wx + 1 < WW ? __ldg(r + wx + 1) : 0u;
  const unsigned long long lo = (unsigned long long)p | ((unsigned long long)c << 32);
  unsigned long long e = lo;
#pragma unroll
  for (int k = 1; k < SX; ++k) e &= (lo >> k) | ((unsigned long long)n << (64 - k));
  return e;
}

template <int SZ, int SY, int SX>
__global__ void __launch_bounds__(kPreThreads) open_kernel(const unsigned int* __restrict__ m, int D, int H, int WW,
                                                           unsigned int* __restrict__ out) {
  const long long total = (long long)D * H * WW;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int wx = (int)(i % WW);
    const long long zy = i / WW;
    const int y = (int)(zy % H), z = (int)(zy / H);
    // E[zq][yq] for the SZ x SY box anchors (zq, yq) in [z-SZ+1, z] x [y-SY+1, y] whose boxes contain (z, y)
    unsigned long long E[SZ][SY];
#pragma unroll
    for (int a = 0; a < SZ; ++a)
#pragma unroll
      for (int b = 0; b < SY; ++b) E[a][b] = ~0ull;
#pragma unroll
    for (int dz = -(SZ - 1); dz <= SZ - 1; ++dz) {
      // y-erosion of plane z + dz for the SY anchors
      unsigned long long Y[SY];
#pragma unroll
      for (int b = 0; b < SY; ++b) Y[b] = ~0ull;
#pragma unroll
      for (int dy = -(SY - 1); dy <= SY - 1; ++dy) {
        const unsigned long long X = row_erode_x<SX>(m, z + dz, y + dy, wx, D, H, WW);
#pragma unroll
        for (int b = 0; b < SY; ++b) {  // anchor yq = y - (SY-1) + b covers rows yq .. yq + SY - 1
          const int off = dy + (SY - 1) - b;
          if (off >= 0 && off < SY) Y[b] &= X;
        }
      }
#pragma unroll
      for (int a = 0; a < SZ; ++a) {
        const int off = dz + (SZ - 1) - a;
        if (off >= 0 && off < SZ) {
#pragma unroll
          for (int b = 0; b < SY; ++b) E[a][b] &= Y[b];
        }
      }
    }
    unsigned long long T = 0ull;
#pragma unroll
    for (int a = 0; a < SZ; ++a)
#pragma unroll
      for (int b = 0; b < SY; ++b) T |= E[a][b];
    // x-dilation: O[x] = OR_k E[x - k]
    unsigned long long O = T;
#pragma unroll
    for (int k = 1; k < SX; ++k) O |= T << k;
    out[i] = (unsigned int)(O >> 32);
  }
}

// ---- pass 3: apply ------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ unsigned int clip1(unsigned int v, unsigned int lo, unsigned int hi) {
  return min(max(v, lo), hi);
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
    uint4* p = reinterpret_cast<uint4*>(vol + row * W) + iv;
    if (!CLIP) {
      if (bits == 0) continue;
      if (bits == (1u << VPT) - 1u) { *p = make_uint4(0, 0, 0, 0); continue; }
    }
    uint4 q = *p;
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
static bool open_dispatch_x(int sx, const unsigned int* m, int D, int H, int WW, unsigned int* out, int grid, cudaStream_t st) {
  switch (sx) {
    case 1: open_kernel<SZ, SY, 1><<<grid, kPreThreads, 0, st>>>(m, D, H, WW, out); return true;
    case 2: open_kernel<SZ, SY, 2><<<grid, kPreThreads, 0, st>>>(m, D, H, WW, out); return true;
    case 3: open_kernel<SZ, SY, 3><<<grid, kPreThreads, 0, st>>>(m, D, H, WW, out); return true;
    case 4: open_kernel<SZ, SY, 4><<<grid, kPreThreads, 0, st>>>(m, D, H, WW, out); return true;
  }
  return false;
}
template <int SZ>
static bool open_dispatch_y(int sy, int sx, const unsigned int* m, int D, int H, int WW, unsigned int* out, int grid,
                            cudaStream_t st) {
  switch (sy) {
    case 1: return open_dispatch_x<SZ, 1>(sx, m, D, H, WW, out, grid, st);
    case 2: return open_dispatch_x<SZ, 2>(sx, m, D, H, WW, out, grid, st);
    case 3: return open_dispatch_x<SZ, 3>(sx, m, D, H, WW, out, grid, st);
    case 4: return open_dispatch_x<SZ, 4>(sx, m, D, H, WW, out, grid, st);
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
  if (any_mask) {
    if (vec) {
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
      const int grid = grid_for(words, kPreThreads);
      bool ok = false;
      switch (sz) {
        case 1: ok = open_dispatch_y<1>(sy, sx, mask, D, H, WW, opened, grid, st); break;
        case 2: ok = open_dispatch_y<2>(sy, sx, mask, D, H, WW, opened, grid, st); break;
        case 3: ok = open_dispatch_y<3>(sy, sx, mask, D, H, WW, opened, grid, st); break;
        case 4: ok = open_dispatch_y<4>(sy, sx, mask, D, H, WW, opened, grid, st); break;
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
  if (vec) {
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
