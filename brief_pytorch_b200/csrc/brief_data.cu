// brief_data.cu — data prologue of the fit on the device: per-block min / max / sum / sum of squares of the raw voxels.
//
// Reference work replaced: normalize_data's data.min() / data.max() (utils/io.py:67-80) and the per-chunk np.var of
// alloc_param 'by_var' (utils/misc.py:402-422) — full host numpy passes over every block.  Here: ONE launch for all
// blocks of a rank, HBM-bound (algorithmic bytes = the raw volume, read once): 16-byte vector loads, 4 independent
// loads in flight per thread, warp-shuffle + shared-memory reduction, one set of atomics per CTA.
// min / max are exact (returned as the float of the raw value, like numpy's astype(float32)); sum and sum of squares
// are accumulated in double.
#include "brief_kernels.h"

namespace brief {

constexpr int kStatThreads = 256;

// order-preserving map float -> uint32 so that atomicMin / atomicMax on the integer give the float min / max
__device__ __forceinline__ unsigned int f2ord(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float ord2f(unsigned int o) {
  const unsigned int u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

struct Acc {
  float mn, mx;            // float data
  unsigned int imn, imx;   // integer data: packed SIMD lanes (4 x u8 / 2 x u16) until the final fold
  double s, ss;            // float data
  unsigned long long is, iss;  // integer data: exact
  __device__ __forceinline__ void add(float v) {
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
    s += (double)v;
    ss += (double)v * (double)v;
  }
};

// one 16-byte vector.  Integer dtypes use the SIMD-in-a-word instructions: 4 (u8) / 2 (u16) lanes per min / max, dp4a
// for the byte sums and sums of squares — ~4 instructions per uint16 voxel instead of ~8 scalar ones, which is what
// lets one SM keep up with its share of HBM.
template <typename T>
__device__ __forceinline__ void acc_vec16(Acc& a, const uint4& q) {
  const unsigned int w[4] = {q.x, q.y, q.z, q.w};
  if (sizeof(T) == 1) {
    unsigned int s = 0, ss = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a.imn = __vminu4(a.imn, w[i]);
      a.imx = __vmaxu4(a.imx, w[i]);
      s = __dp4a(w[i], 0x01010101u, s);
      ss = __dp4a(w[i], w[i], ss);
    }
    a.is += s;
    a.iss += ss;
  } else if (sizeof(T) == 2) {
    unsigned int s = 0;
    unsigned long long ss = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a.imn = __vminu2(a.imn, w[i]);
      a.imx = __vmaxu2(a.imx, w[i]);
      const unsigned int lo = w[i] & 0xffffu, hi = w[i] >> 16;
      s += lo + hi;
      ss += (unsigned long long)(lo * lo) + (unsigned long long)(hi * hi);  // 65535^2 < 2^32
    }
    a.is += s;
    a.iss += ss;
  } else {
    const float* e = reinterpret_cast<const float*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) a.add(e[i]);
  }
}
template <typename T>
__device__ __forceinline__ void acc_scalar(Acc& a, T v) {
  if (sizeof(T) == 4) {
    a.add((float)v);
  } else {
    const unsigned int u = (unsigned int)v;
    const unsigned int rep = sizeof(T) == 1 ? u * 0x01010101u : u * 0x00010001u;  // replicate into every SIMD lane
    a.imn = sizeof(T) == 1 ? __vminu4(a.imn, rep) : __vminu2(a.imn, rep);
    a.imx = sizeof(T) == 1 ? __vmaxu4(a.imx, rep) : __vmaxu2(a.imx, rep);
    a.is += u;
    a.iss += (unsigned long long)u * u;
  }
}
// fold the integer lanes / exact sums into the float / double fields used by the reduction
template <typename T>
__device__ __forceinline__ void acc_finish(Acc& a) {
  if (sizeof(T) == 4) return;
  unsigned int mn, mx;
  if (sizeof(T) == 1) {
    mn = min(min(a.imn & 0xff, (a.imn >> 8) & 0xff), min((a.imn >> 16) & 0xff, a.imn >> 24));
    mx = max(max(a.imx & 0xff, (a.imx >> 8) & 0xff), max((a.imx >> 16) & 0xff, a.imx >> 24));
  } else {
    mn = min(a.imn & 0xffff, a.imn >> 16);
    mx = max(a.imx & 0xffff, a.imx >> 16);
  }
  if (mn > mx) { a.mn = INFINITY; a.mx = -INFINITY; }  // this thread saw no element (all lanes untouched)
  else { a.mn = (float)mn; a.mx = (float)mx; }
  a.s = (double)a.is;    // exact below 2^53
  a.ss = (double)a.iss;
}

// out[b] = {ord(min), ord(max)} as uint32 in stat_ord[2b..], {sum, sumsq} in stat_sum[2b..]
template <typename T>
__global__ void __launch_bounds__(kStatThreads) stats_kernel(const void* const* ptrs, const long long* sizes,
                                                             unsigned int* stat_ord, double* stat_sum) {
  const int b = blockIdx.y;
  const T* p = reinterpret_cast<const T*>(ptrs[b]);
  const long long n = sizes[b];
  constexpr int VE = 16 / sizeof(T);
  Acc a{INFINITY, -INFINITY, 0xffffffffu, 0u, 0.0, 0.0, 0ull, 0ull};
  // head: elements before the first 16-byte boundary; body: vectors; tail: the rest
  const long long mis = (long long)((16 - ((uintptr_t)p & 15)) & 15) / (long long)sizeof(T);
  const long long head = mis < n ? mis : n;
  const long long n_vec = (n - head) / VE;
  const uint4* pv = reinterpret_cast<const uint4*>(p + head);
  const long long tid = (long long)blockIdx.x * kStatThreads + threadIdx.x, stride = (long long)gridDim.x * kStatThreads;
  long long i = tid;
  for (; i + 3 * stride < n_vec; i += 4 * stride) {  // four independent 16-byte loads in flight
    const uint4 q0 = __ldg(pv + i), q1 = __ldg(pv + i + stride), q2 = __ldg(pv + i + 2 * stride), q3 = __ldg(pv + i + 3 * stride);
    acc_vec16<T>(a, q0); acc_vec16<T>(a, q1); acc_vec16<T>(a, q2); acc_vec16<T>(a, q3);
  }
  for (; i < n_vec; i += stride) acc_vec16<T>(a, __ldg(pv + i));
  if (blockIdx.x == 0) {
    for (long long k = threadIdx.x; k < head; k += kStatThreads) acc_scalar<T>(a, p[k]);
    for (long long k = head + n_vec * VE + threadIdx.x; k < n; k += kStatThreads) acc_scalar<T>(a, p[k]);
  }
  acc_finish<T>(a);
  // warp, then CTA reduction
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    a.mn = fminf(a.mn, __shfl_down_sync(0xffffffffu, a.mn, off));
    a.mx = fmaxf(a.mx, __shfl_down_sync(0xffffffffu, a.mx, off));
    a.s += __shfl_down_sync(0xffffffffu, a.s, off);
    a.ss += __shfl_down_sync(0xffffffffu, a.ss, off);
  }
  __shared__ float s_mn[kStatThreads / 32], s_mx[kStatThreads / 32];
  __shared__ double s_s[kStatThreads / 32], s_ss[kStatThreads / 32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s_mn[w] = a.mn; s_mx[w] = a.mx; s_s[w] = a.s; s_ss[w] = a.ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < kStatThreads / 32; ++k) {
      a.mn = fminf(a.mn, s_mn[k]); a.mx = fmaxf(a.mx, s_mx[k]); a.s += s_s[k]; a.ss += s_ss[k];
    }
    if (a.mn <= a.mx) {
      atomicMin(stat_ord + 2 * b, f2ord(a.mn));
      atomicMax(stat_ord + 2 * b + 1, f2ord(a.mx));
    }
    atomicAdd(stat_sum + 2 * b, a.s);
    atomicAdd(stat_sum + 2 * b + 1, a.ss);
  }
}

cudaError_t launch_block_stats(const void* const* dev_ptrs, const long long* dev_sizes, int n_blocks, long long max_size,
                               int dtype, unsigned int* stat_ord, double* stat_sum, int num_sms, cudaStream_t st) {
  if (n_blocks < 1) return cudaSuccess;
  // enough CTAs to fill the machine (8 resident per SM) without leaving most of them idle on small blocks
  const int esz = dtype == 0 ? 1 : dtype == 1 ? 2 : 4;
  long long per_cta = (long long)kStatThreads * 4 * (16 / esz);
  long long gx = (max_size + per_cta - 1) / per_cta;
  const long long cap = std::max<long long>(1, (long long)num_sms * 8 / n_blocks);
  gx = std::max<long long>(1, std::min<long long>(gx, cap));
  dim3 grid((unsigned)gx, (unsigned)n_blocks);
  if (dtype == 0) stats_kernel<unsigned char><<<grid, kStatThreads, 0, st>>>(dev_ptrs, dev_sizes, stat_ord, stat_sum);
  else if (dtype == 1) stats_kernel<unsigned short><<<grid, kStatThreads, 0, st>>>(dev_ptrs, dev_sizes, stat_ord, stat_sum);
  else stats_kernel<float><<<grid, kStatThreads, 0, st>>>(dev_ptrs, dev_sizes, stat_ord, stat_sum);
  return cudaGetLastError();
}

float stats_ord_to_float(unsigned int o) { return ord2f(o); }

// ---- %nsmid: the wide fit kernel indexes its per-SM scratch by %smid, whose range is [0, %nsmid) — not the SM count ----
__global__ void nsmid_kernel(unsigned int* out) {
  unsigned int v;
  asm volatile("mov.u32 %0, %%nsmid;" : "=r"(v));
  *out = v;
}
cudaError_t query_nsmid(int* out, cudaStream_t st) {
  unsigned int* d = nullptr;
  cudaError_t e = cudaMalloc(&d, sizeof(unsigned int));
  if (e != cudaSuccess) return e;
  nsmid_kernel<<<1, 1, 0, st>>>(d);
  unsigned int h = 0;
  e = cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d);
  *out = (int)h;
  return e;
}

// ---- value histogram of a uint8 / uint16 block (quantile weight rules, utils/misc.py:298-305) -------------------------
// hist[v] += number of voxels equal to v.  Biomedical volumes put most voxels into a handful of background values, so
// equal values inside a warp are merged first (match.any) and one lane adds the whole count: a hot bin costs one
// atomic per warp instead of 32.  Algorithmic bytes = the block, read once; not on the per-step path.
template <typename T>
__global__ void __launch_bounds__(256) hist_kernel(const T* __restrict__ p, long long n, unsigned long long* __restrict__ hist) {
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long base = (long long)blockIdx.x * blockDim.x + (threadIdx.x - lane); base < n; base += stride) {
    const long long i = base + lane;
    const bool ok = i < n;
    const unsigned int v = ok ? (unsigned int)__ldg(p + i) : 0xffffffffu;
    const unsigned int peers = __match_any_sync(0xffffffffu, v);
    if (ok && lane == __ffs(peers) - 1) atomicAdd(hist + v, (unsigned long long)__popc(peers));
  }
}

cudaError_t launch_histogram(const void* dev_raw, long long n, int dtype, unsigned long long* dev_hist, int num_sms,
                             cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const long long want = (n + 255) / 256;
  const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)num_sms * 16));
  if (dtype == 0) hist_kernel<unsigned char><<<grid, 256, 0, st>>>(reinterpret_cast<const unsigned char*>(dev_raw), n, dev_hist);
  else if (dtype == 1) hist_kernel<unsigned short><<<grid, 256, 0, st>>>(reinterpret_cast<const unsigned short*>(dev_raw), n, dev_hist);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace brief
