// brief_simt.cu — fp32 CUDA-core kernels: the exact-arithmetic mode of the hot path
// (BRIEF_PREC_FP32) and the on-GPU cross-check of the tcgen05 kernels.
//
// Reference work replaced (file:line relative to the reference root):
//   SIREN.forward                utils/Networks.py:269-271
//   datal2 loss                  main.py:176-182
//   autograd backward            main.py:396
//   RandompointSampler/Cube      main.py:38-163      (gather fused: index -> raw voxel -> normalise)
//   reconstruct_flattened        utils/misc.py:59-92 (dense grid generated on chip)
//   invnormalize_data            utils/io.py:136-147 (epilogue)
//
// One CTA = 128 .. 512 threads (512 when the tile is <= 32 rows, i.e. wide networks) = one tile of TM samples of one
// network; thread t serves sample
// m = t % TM and output-feature group g = t / TM.  Activations live in shared memory, sample-major
// [TM][S] with S/4 odd so that per-sample float4 row reads are bank-conflict free.  Weights are
// read with warp-uniform 16-byte __ldg (L1 resident).  Each layer is a register-blocked
// (4 outputs x 4 k) mini-GEMM.  Roofline: fp32 FMA pipe (this path is not the performance path).
#include "brief_common.cuh"
#include "brief_kernels.h"

namespace brief {

constexpr int kThreads = 128;      // default CTA size
constexpr int kMaxThreads = 512;   // wide networks (small TM): more output-feature groups per sample row

__host__ __device__ inline int row_stride(int F4) { return F4 + (((F4 >> 2) & 1) ? 0 : 4); }

__device__ __forceinline__ int find_work(const int* __restrict__ prefix, int n, int b) {
  int lo = 0, hi = n;  // largest i with prefix[i] <= b
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(prefix + mid) <= b) lo = mid; else hi = mid;
  }
  return lo;
}

// Stage the network descriptor in shared memory (its fields are read in every inner loop).
__device__ __forceinline__ void load_net(NetDev& dst, const NetDev& src) {
  static_assert(sizeof(NetDev) % 4 == 0, "NetDev must be word-sized");
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&src);
  uint32_t* d = reinterpret_cast<uint32_t*>(&dst);
  for (int i = threadIdx.x; i < (int)(sizeof(NetDev) / 4); i += blockDim.x) d[i] = __ldg(s + i);
  __syncthreads();
}

// ---- layer primitives --------------------------------------------------------------------------
// First layer: a0[m][o] = sin(w0 * (W0[o] . x + b0[o])), optionally cos and raw z.
template <bool WITH_COS>
__device__ __forceinline__ void first_layer(const NetDev& n, const float* __restrict__ P, const float* X,
                                            float* A0, float* C0, int S, int TM, int m, int g, int G,
                                            float* zdump, long long zrow) {
  const float4 x = *reinterpret_cast<const float4*>(X + 4 * m);
  const float* W0 = P + dl_W0(n);
  const float* b0 = P + dl_b0(n);
  for (int o4 = g; o4 < (n.F4 >> 2); o4 += G) {
    float zs[4], s[4], c[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int o = 4 * o4 + i;
      const float4 w = __ldg(reinterpret_cast<const float4*>(W0 + 4 * o));
      float z = __ldg(b0 + o);
      z = fmaf(w.x, x.x, z); z = fmaf(w.y, x.y, z); z = fmaf(w.z, x.z, z);
      zs[i] = z;
      sincosf(__fmul_rn(n.w0, z), &s[i], &c[i]);
    }
    *reinterpret_cast<float4*>(A0 + m * S + 4 * o4) = make_float4(s[0], s[1], s[2], s[3]);
    if (WITH_COS) *reinterpret_cast<float4*>(C0 + m * S + 4 * o4) = make_float4(c[0], c[1], c[2], c[3]);
    if (zdump && zrow >= 0)
      for (int i = 0; i < 4; ++i)
        if (4 * o4 + i < n.f) zdump[zrow * n.f + 4 * o4 + i] = zs[i];
  }
}

// Hidden layer l: Aout[m][o] = sin(wh * (W[o] . Ain[m] + b[o])).
template <bool WITH_COS>
__device__ __forceinline__ void hidden_layer(const NetDev& n, const float* __restrict__ W, const float* __restrict__ b,
                                             const float* Ain, float* Aout, float* Cout, int S, int m, int g,
                                             int G, float* zdump, long long zrow) {
  const int F4 = n.F4;
  const float* arow = Ain + m * S;
  for (int o4 = g; o4 < (F4 >> 2); o4 += G) {
    float acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = __ldg(b + 4 * o4 + i);
    for (int k = 0; k < F4; k += 4) {
      const float4 a = *reinterpret_cast<const float4*>(arow + k);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(W + (4 * o4 + i) * F4 + k));
        acc[i] = fmaf(w.x, a.x, acc[i]); acc[i] = fmaf(w.y, a.y, acc[i]);
        acc[i] = fmaf(w.z, a.z, acc[i]); acc[i] = fmaf(w.w, a.w, acc[i]);
      }
    }
    float s[4], c[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) sincosf(__fmul_rn(n.wh, acc[i]), &s[i], &c[i]);
    *reinterpret_cast<float4*>(Aout + m * S + 4 * o4) = make_float4(s[0], s[1], s[2], s[3]);
    if (WITH_COS) *reinterpret_cast<float4*>(Cout + m * S + 4 * o4) = make_float4(c[0], c[1], c[2], c[3]);
    if (zdump && zrow >= 0)
      for (int i = 0; i < 4; ++i)
        if (4 * o4 + i < n.f) zdump[zrow * n.f + 4 * o4 + i] = acc[i];
  }
}

// Last (linear) layer, one output channel: y[m] = Wlast . A[m] + blast.
__device__ __forceinline__ float last_layer(const NetDev& n, const float* __restrict__ P, const float* A, int S, int m) {
  const float* Wl = P + dl_Wlast(n);
  float acc = __ldg(P + dl_blast(n));
  const float* arow = A + m * S;
  for (int k = 0; k < n.F4; k += 4) {
    const float4 a = *reinterpret_cast<const float4*>(arow + k);
    const float4 w = __ldg(reinterpret_cast<const float4*>(Wl + k));
    acc = fmaf(w.x, a.x, acc); acc = fmaf(w.y, a.y, acc); acc = fmaf(w.z, a.z, acc); acc = fmaf(w.w, a.w, acc);
  }
  return acc;
}

// ---- forward / decompress ----------------------------------------------------------------------
__global__ void __launch_bounds__(kMaxThreads) simt_eval_kernel(EvalArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ NetDev sn;
  int net_id;
  long long tile;
  if (a.single_net >= 0) {
    net_id = a.single_net;
    tile = blockIdx.x;
  } else {
    const int wi = find_work(a.work_prefix, a.n_work, blockIdx.x);
    net_id = a.work_net[wi];
    tile = blockIdx.x - a.work_prefix[wi];
    if (a.out_ptrs[net_id] == nullptr) return;  // this network is not part of the call
  }
  load_net(sn, a.nets[net_id]);
  const NetDev& n = sn;
  const int NT = blockDim.x;
  const int TM = a.TM, G = NT / TM;
  const int t = threadIdx.x, m = t % TM, g = t / TM;
  const int S = row_stride(n.F4);
  const float* P = a.params + n.param_off;
  float* bufA = smem;
  float* bufB = smem + TM * S;
  float* X = smem + 2 * TM * S;

  const long long total = a.coords ? a.n_coords : n.n_vox;
  const long long s = tile * TM + m;
  const bool valid = s < total;
  if (g == 0) {
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
    if (valid) {
      if (a.coords) {
        c0 = a.coords[s * n.in_dim];
        c1 = a.coords[s * n.in_dim + 1];
        c2 = n.in_dim == 3 ? a.coords[s * n.in_dim + 2] : 0.f;
      } else {
        brief_coords(n, a.axes, s, c0, c1, c2);
      }
    }
    *reinterpret_cast<float4*>(X + 4 * m) = make_float4(c0, c1, c2, 0.f);
  }
  __syncthreads();
  float* zd = a.layers_out;
  const long long zrow = valid ? s : -1;
  first_layer<false>(n, P, X, bufA, nullptr, S, TM, m, g, G, zd, zrow);
  __syncthreads();
  float* in = bufA;
  float* out = bufB;
  for (int l = 1; l <= n.L - 2; ++l) {
    hidden_layer<false>(n, P + dl_W(n, l), P + dl_b(n, l), in, out, nullptr, S, m, g, G,
                        zd ? zd + (long long)l * total * n.f : nullptr, zrow);
    __syncthreads();
    float* tmp = in; in = out; out = tmp;
  }
  if (g == 0 && valid) {
    const float y = last_layer(n, P, in, S, m);
    if (a.out_f32) {
      a.out_f32[s] = y;
    } else {
      void* dst = a.out_ptrs[net_id];
      if (a.out_dtype == 2) {
        reinterpret_cast<float*>(dst)[s] = y;
      } else {
        const float v = brief_denorm(n, y);
        if (a.out_dtype == 1) reinterpret_cast<unsigned short*>(dst)[s] = (unsigned short)(int)v;
        else reinterpret_cast<unsigned char*>(dst)[s] = (unsigned char)(int)v;
      }
    }
  }
}

// ---- fit: gather + forward + weighted L2 + backward -> per-slice gradient partials ---------------
__global__ void __launch_bounds__(kMaxThreads) simt_fit_kernel(FitArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ NetDev sn;
  const int wi = find_work(a.work_prefix, a.n_work, blockIdx.x);
  const int net_id = a.work_net[wi];
  load_net(sn, a.nets[net_id]);
  const NetDev& n = sn;
  const int slice = blockIdx.x - a.work_prefix[wi];
  const int NT = blockDim.x;
  const int TM = a.TM, G = NT / TM;
  const int t = threadIdx.x, m = t % TM, g = t / TM;
  const int F4 = n.F4, S = row_stride(F4), nl = n.L - 1;  // nl sine layers
  const float* P = a.params + n.param_off;
  float* part = a.partials + n.part_off + (long long)slice * n.P_dev;

  float* A = smem;                         // [nl][TM][S]  sin activations
  float* C = smem + (size_t)nl * TM * S;   // [nl][TM][S]  cos, overwritten in place by dz
  float* X = C + (size_t)nl * TM * S;      // [TM][4]
  float* DY = X + 4 * TM;                  // [TM]
  float* LS = DY + TM;                     // [TM] per-sample loss terms

  const long long s_begin = (long long)slice * n.slice_len;
  const long long s_end = min((long long)n.batch, s_begin + n.slice_len);
  const float inv_count = 1.0f / ((float)n.batch * (float)n.out_dim);
  float loss_acc = 0.f;  // thread 0 only

  for (long long tile0 = s_begin, it = 0; tile0 < s_end; tile0 += TM, ++it) {
    const long long s = tile0 + m;
    const bool valid = s < s_end;
    float yv = 0.f, wv = 0.f;
    if (g == 0) {
      float c0 = 0.f, c1 = 0.f, c2 = 0.f;
      if (valid) {
        long long idx;
        if (n.mode == 0) idx = s;
        else if (a.idx) idx = a.idx[n.idx_off + s];
        else idx = brief_sample_index(a.seed, a.state ? a.state->step : a.step, n.stream_id, (uint64_t)s, (uint64_t)n.n_vox);
        brief_coords(n, a.axes, idx, c0, c1, c2);
        const float raw = brief_raw_value(n, idx);
        yv = brief_normalize(n, raw);
        wv = brief_weight(n, idx, raw);
      }
      *reinterpret_cast<float4*>(X + 4 * m) = make_float4(c0, c1, c2, 0.f);
    }
    __syncthreads();
    // ---- forward
    first_layer<true>(n, P, X, A, C, S, TM, m, g, G, nullptr, -1);
    __syncthreads();
    for (int l = 1; l <= n.L - 2; ++l) {
      hidden_layer<true>(n, P + dl_W(n, l), P + dl_b(n, l), A + (size_t)(l - 1) * TM * S, A + (size_t)l * TM * S,
                         C + (size_t)l * TM * S, S, m, g, G, nullptr, -1);
      __syncthreads();
    }
    const float* Alast = A + (size_t)(nl - 1) * TM * S;
    if (g == 0) {
      float dy = 0.f, lt = 0.f;
      if (valid) {
        const float yhat = last_layer(n, P, Alast, S, m);
        const float e = yhat - yv;
        const float wt = (n.tau != 0.f && yhat <= n.tau) ? 1.0f : wv;  // main.py:178-179
        lt = wt * e * e;
        dy = 2.0f * wt * e * inv_count;
      }
      DY[m] = dy;
      LS[m] = lt;
    }
    __syncthreads();
    if (t == 0) {
      float acc = 0.f;
      for (int i = 0; i < TM; ++i) acc += LS[i];
      loss_acc += acc;
    }
    // ---- backward.  dz of the last sine layer, in place into C[nl-1]
    {
      const float* Wl = P + dl_Wlast(n);
      float* Cl = C + (size_t)(nl - 1) * TM * S;
      const float om = (nl - 1 == 0) ? n.w0 : n.wh;
      const float dy = DY[m];
      for (int k4 = g; k4 < (F4 >> 2); k4 += G) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(Wl + 4 * k4));
        float4 c = *reinterpret_cast<float4*>(Cl + m * S + 4 * k4);
        c.x = dy * w.x * om * c.x; c.y = dy * w.y * om * c.y; c.z = dy * w.z * om * c.z; c.w = dy * w.w * om * c.w;
        *reinterpret_cast<float4*>(Cl + m * S + 4 * k4) = c;
      }
      // dWlast[k] = sum_m dy[m] * A[m][k] ; dblast = sum_m dy[m]
      if (t < F4) {
        float acc = 0.f;
        for (int i = 0; i < TM; ++i) acc = fmaf(DY[i], Alast[i * S + t], acc);
        float* dst = part + dl_Wlast(n) + t;
        *dst = it ? *dst + acc : acc;
      }
      if (t == NT - 1) {
        float acc = 0.f;
        for (int i = 0; i < TM; ++i) acc += DY[i];
        float* dst = part + dl_blast(n);
        dst[0] = it ? dst[0] + acc : acc;
        if (!it) { dst[1] = 0.f; dst[2] = 0.f; dst[3] = 0.f; }
      }
    }
    __syncthreads();
    for (int l = n.L - 2; l >= 1; --l) {
      const float* DZ = C + (size_t)l * TM * S;           // dz_l
      const float* Ain = A + (size_t)(l - 1) * TM * S;    // a_{l-1}
      float* Cprev = C + (size_t)(l - 1) * TM * S;        // cos_{l-1} -> dz_{l-1}
      const float* W = P + dl_W(n, l);
      const int nb = F4 >> 2;
      // dW_l[o][k] = sum_m dz[m][o] a[m][k]   (4x4 register tiles, lanes consecutive in k)
      for (int id = t; id < nb * nb; id += NT) {
        const int o4 = id / nb, k4 = id - o4 * nb;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int i2 = 0; i2 < TM; ++i2) {
          const float4 dz = *reinterpret_cast<const float4*>(DZ + i2 * S + 4 * o4);
          const float4 av = *reinterpret_cast<const float4*>(Ain + i2 * S + 4 * k4);
          const float d[4] = {dz.x, dz.y, dz.z, dz.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[i][0] = fmaf(d[i], av.x, acc[i][0]); acc[i][1] = fmaf(d[i], av.y, acc[i][1]);
            acc[i][2] = fmaf(d[i], av.z, acc[i][2]); acc[i][3] = fmaf(d[i], av.w, acc[i][3]);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4* dst = reinterpret_cast<float4*>(part + dl_W(n, l) + (4 * o4 + i) * F4 + 4 * k4);
          float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
          if (it) { const float4 p = *dst; v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w; }
          *dst = v;
        }
      }
      // db_l[o] = sum_m dz[m][o]
      if (t < F4) {
        float acc = 0.f;
        for (int i = 0; i < TM; ++i) acc += DZ[i * S + t];
        float* dst = part + dl_b(n, l) + t;
        *dst = it ? *dst + acc : acc;
      }
      // dz_{l-1}[m][k] = (sum_o dz[m][o] W[o][k]) * omega_{l-1} * cos_{l-1}[m][k]
      const float om = (l - 1 == 0) ? n.w0 : n.wh;
      const float* drow = DZ + m * S;
      for (int k4 = g; k4 < nb; k4 += G) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int o = 0; o < F4; o += 4) {
          const float4 dz = *reinterpret_cast<const float4*>(drow + o);
          const float d[4] = {dz.x, dz.y, dz.z, dz.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(W + (o + j) * F4 + 4 * k4));
            acc[0] = fmaf(d[j], w.x, acc[0]); acc[1] = fmaf(d[j], w.y, acc[1]);
            acc[2] = fmaf(d[j], w.z, acc[2]); acc[3] = fmaf(d[j], w.w, acc[3]);
          }
        }
        float4 c = *reinterpret_cast<float4*>(Cprev + m * S + 4 * k4);
        c.x = acc[0] * om * c.x; c.y = acc[1] * om * c.y; c.z = acc[2] * om * c.z; c.w = acc[3] * om * c.w;
        *reinterpret_cast<float4*>(Cprev + m * S + 4 * k4) = c;
      }
      __syncthreads();
    }
    // layer 0: dW0[o][c] = sum_m dz0[m][o] x[m][c]; db0[o] = sum_m dz0[m][o]
    if (t < F4) {
      const float* DZ0 = C;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, ab = 0.f;
      for (int i = 0; i < TM; ++i) {
        const float dz = DZ0[i * S + t];
        const float4 x = *reinterpret_cast<const float4*>(X + 4 * i);
        a0 = fmaf(dz, x.x, a0); a1 = fmaf(dz, x.y, a1); a2 = fmaf(dz, x.z, a2); ab += dz;
      }
      float4* dw = reinterpret_cast<float4*>(part + dl_W0(n) + 4 * t);
      float* db = part + dl_b0(n) + t;
      float4 v = make_float4(a0, a1, a2, 0.f);
      if (it) { const float4 p = *dw; v.x += p.x; v.y += p.y; v.z += p.z; ab += *db; }
      *dw = v;
      *db = ab;
    }
    __syncthreads();
  }
  if (t == 0) a.loss_partials[n.slice_off + slice] = loss_acc * inv_count;
}

// ---- reference sampler outputs materialised (main.py:156-160) ---------------------------------------
// HBM-bound on random sectors: the descriptor is staged in shared memory once, and every thread keeps FOUR samples'
// dependent chains (index -> voxel, index -> axis tables) in flight before it normalises and stores.
__global__ void __launch_bounds__(256) gather_kernel(const NetDev* nets, int net_id, const float* __restrict__ axes,
                                                     const long long* __restrict__ idx, long long batch, float* coords,
                                                     float* data, float* weight) {
  __shared__ NetDev sn;
  load_net(sn, nets[net_id]);
  const NetDev& n = sn;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < batch; i0 += 4 * stride) {
    long long v[4];
    bool ok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long i = i0 + k * stride;
      ok[k] = i < batch;
      v[k] = ok[k] ? (idx ? __ldg(idx + i) : i) : 0;
    }
    float raw[4], c[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      raw[k] = brief_raw_value(n, v[k]);
      brief_coords(n, axes, v[k], c[k][0], c[k][1], c[k][2]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!ok[k]) continue;
      const long long i = i0 + k * stride;
      if (coords) {
        coords[i * n.in_dim] = c[k][0];
        coords[i * n.in_dim + 1] = c[k][1];
        if (n.in_dim == 3) coords[i * n.in_dim + 2] = c[k][2];
      }
      if (data) data[i] = brief_normalize(n, raw[k]);
      if (weight) weight[i] = brief_weight(n, v[k], raw[k]);
    }
  }
}

__global__ void sample_indices_kernel(uint64_t seed, uint64_t step, uint32_t net, long long batch, long long pop,
                                      long long* out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < batch;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = brief_sample_index(seed, step, net, (uint64_t)i, (uint64_t)pop);
}

// One network's slice of the step's explicit index buffer, written when a group holds sliding-cube samplers (the fit
// kernels then run in their replayed-index mode).  cube_vox == 0: `batch` point indices, the stream the fit kernels draw
// on chip.  cube_vox > 0: batch / cube_vox cubes of cube_vox voxels each; cube c is cube_ids[c] when the caller replays
// the reference's torch.randint draws (main.py:114), else draw c of the network's Philox stream over the `pop` cubes.
__global__ void gen_indices_kernel(uint64_t seed, uint64_t step, const StepState* state, uint32_t stream, long long batch,
                                   long long pop, int h, int w, int ch, int cw, long long cube_vox,
                                   const long long* __restrict__ cube_ids, long long* out) {
  if (state) step = state->step;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < batch; i += (long long)gridDim.x * blockDim.x) {
    if (cube_vox == 0) {
      out[i] = brief_sample_index(seed, step, stream, (uint64_t)i, (uint64_t)pop);
    } else {
      const long long c = i / cube_vox, o = i - c * cube_vox;
      const long long cube = cube_ids ? __ldg(cube_ids + c) : brief_sample_index(seed, step, stream, (uint64_t)c, (uint64_t)pop);
      out[i] = brief_cube_voxel(h, w, ch, cw, cube, o);
    }
  }
}

// ---- host launchers ------------------------------------------------------------------------------------
// A tile of TM <= 32 rows (wide networks: the activations of all layers fill shared memory, one CTA per SM) is served
// by 512 threads — 16 output-feature groups per row instead of 4 — so that the SM has 16 warps to hide latency with.
static int simt_threads(int TM) { return TM <= 32 ? kMaxThreads : TM <= 64 ? 256 : kThreads; }

int simt_pick_tm(int F4, int L, bool fit, size_t smem_limit) {
  const int S = row_stride(F4);
  for (int TM = 128; TM >= 8; TM >>= 1) {
    const size_t need = fit ? simt_fit_smem(F4, L, TM) : simt_eval_smem(F4, TM);
    (void)S;
    if (need <= smem_limit) return TM;
  }
  return 0;
}
size_t simt_eval_smem(int F4, int TM) { return (size_t)(2 * TM * row_stride(F4) + 4 * TM) * sizeof(float); }
size_t simt_fit_smem(int F4, int L, int TM) {
  return ((size_t)2 * (L - 1) * TM * row_stride(F4) + 6 * (size_t)TM) * sizeof(float);
}

cudaError_t launch_simt_eval(const EvalArgs& a, int n_blocks, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(simt_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  simt_eval_kernel<<<n_blocks, simt_threads(a.TM), smem, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_simt_fit(const FitArgs& a, int n_blocks, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(simt_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  simt_fit_kernel<<<n_blocks, simt_threads(a.TM), smem, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_gather(const NetDev* nets, int net_id, const float* axes, const long long* idx, long long batch,
                          float* coords, float* data, float* weight, cudaStream_t st) {
  const int threads = 256;
  long long blocks = (batch + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  gather_kernel<<<(int)blocks, threads, 0, st>>>(nets, net_id, axes, idx, batch, coords, data, weight);
  return cudaGetLastError();
}
cudaError_t launch_sample_indices(uint64_t seed, uint64_t step, uint32_t net, long long batch, long long pop,
                                  long long* out, cudaStream_t st) {
  const int threads = 256;
  long long blocks = (batch + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  sample_indices_kernel<<<(int)blocks, threads, 0, st>>>(seed, step, net, batch, pop, out);
  return cudaGetLastError();
}

cudaError_t launch_gen_indices(uint64_t seed, uint64_t step, const StepState* state, uint32_t stream, long long batch,
                               long long pop, int h, int w, int ch, int cw, long long cube_vox, const long long* cube_ids,
                               long long* out, cudaStream_t st) {
  const int threads = 256;
  long long blocks = (batch + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  gen_indices_kernel<<<(int)blocks, threads, 0, st>>>(seed, step, state, stream, batch, pop, h, w, ch, cw, cube_vox, cube_ids, out);
  return cudaGetLastError();
}

}  // namespace brief
