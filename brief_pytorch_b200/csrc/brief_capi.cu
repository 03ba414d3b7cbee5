// brief_capi.cu — the extern "C" boundary of libbrief_b200 (include/brief_b200.h) and the host-side
// group bookkeeping: arenas, padded<->packed parameter conversion, work tables, launch sequencing.
// No torch types cross this boundary; no CPU compute path exists behind it.
#include <algorithm>
#include <array>
#include <map>
#include <set>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/brief_b200.h"
#include "brief_common.cuh"
#include "brief_kernels.h"

using namespace brief;

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(expr)                                                                                     \
  do {                                                                                               \
    cudaError_t e__ = (expr);                                                                        \
    if (e__ != cudaSuccess)                                                                          \
      return fail(BRIEF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                  __LINE__);                                                                         \
  } while (0)
#define LAUNCH(expr)      \
  do {                    \
    CU(expr);             \
    g_launches.fetch_add(1); \
  } while (0)
#define RC(expr)          \
  do {                    \
    int rc__ = (expr);    \
    if (rc__) return rc__; \
  } while (0)

constexpr size_t kSmemLimit = 200 * 1024;       // dynamic smem budget of the fp32 kernels
constexpr size_t kPartialCapBytes = 16u << 20;  // gradient-partial budget per network
constexpr int kTcTile = 128;                    // samples per tcgen05 tile (UMMA M)
// tiles per decompress CTA: a multiple of 8 (the kernel's tile slots), sized for >= ~4 CTAs per SM when the work allows
int tc_eval_tpb(long long total_tiles, int num_sms) {
  long long t = total_tiles / (4LL * num_sms);
  t = (t + 7) / 8 * 8;
  return (int)std::min<long long>(128, std::max<long long>(8, t));
}
constexpr int kBuckets = 9;  // F_PAD / 16 in 1..8
constexpr int kHostStepSlots = 4;  // distinct (host index buffer, host loss buffer) pairs with a cached step graph

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t ensure(size_t count) {
    if (count <= n && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) n = std::max<size_t>(count, 1);
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

struct WorkTable {  // block -> (network, local work index) for one kernel family
  int n = 0;        // networks
  int blocks = 0;
  int off_prefix = 0, off_net = 0;  // int offsets into the device table buffer
};

}  // namespace

// One cached CUDA graph of a whole training step driven from HOST buffers (brief_fit_step_host):
//   [H2D step scalars] -> [H2D sampler indices] -> fit kernel(s) -> optimiser kernel -> [D2H per-network loss]
struct HostStepGraph {
  const void* host_idx = nullptr;
  float* host_loss = nullptr;
  cudaStream_t stream = nullptr;
  BriefOptConfig cfg{};
  uint64_t seed = 0;
  cudaGraphExec_t exec = nullptr;
  cudaEvent_t done = nullptr;      // completion of the slot's last submission (its pinned scalars may be rewritten after it)
  cudaEvent_t copied = nullptr;    // this submission's inputs have arrived (copy stream)
  StepState* h_state = nullptr;    // pinned
  StepState* d_state = nullptr;
  long long* d_idx = nullptr;
  float* d_loss = nullptr;
  int kernels = 0;                 // kernel nodes per launch (gpu_launches accounting)
};

// One pass of the layer-wise path: networks `nets` (all of padded width F and depth L) whose tiles share the scratch arena.
struct LwPass {
  int F = 0, L = 0;
  long long T = 0;            // tiles of the pass
  std::vector<int> nets, tile_count, tile_first, tile_prefix, cta_prefix;
  std::vector<long long> tile_base;
  int off = 0, off_base = 0;  // offsets of this pass's tables in the device table buffers
  int max_slices = 1;
  int ctas() const { return cta_prefix.empty() ? 0 : cta_prefix.back(); }
};

struct BriefGroup {
  int device = 0;
  int num_sms = 148;
  int nsmid = 0;            // PTX %nsmid, queried when a wide tensor-core bucket first needs its per-SM scratch
  HostStepGraph host_steps[kHostStepSlots];
  int host_step_next = 0;
  cudaStream_t copy_stream = nullptr;     // host -> device copies of brief_fit_step_host (under the previous step's kernels)
  cudaStream_t capture_stream = nullptr;  // step graphs are captured here (the caller's stream may be the legacy default
                                          // stream, which cannot be captured) and launched into the caller's stream
  int n_nets = 0;
  std::vector<NetDev> nets;
  long long total_P = 0, total_axis = 0;
  size_t total_wpack = 0;
  DevBuf<NetDev> d_nets;
  DevBuf<float> d_params, d_grads, d_m, d_v, d_axes, d_partials, d_loss_partials, d_loss_scratch;
  DevBuf<unsigned char> d_wpack, d_stash;
  size_t stash_stride = 0;  // wide tensor-core fit kernel: activation stash bytes per CTA
  DevBuf<int> d_fit_tables, d_eval_tables;
  DevBuf<void*> d_outptrs;
  bool nets_dirty = true;   // host NetDev array differs from the device copy
  bool work_dirty = true;   // sampler changed: rebuild the fit decomposition
  int slicing = 0;          // BriefSlicing
  bool wpack_dirty = true;  // params changed since the fp16 operand image was packed
  bool any_tc = false, any_simt = false;
  // fit
  int simt_fit_tm = 0;
  size_t simt_fit_smem = 0;
  WorkTable simt_fit, tc_fit[kBuckets], opt;
  // eval (static)
  int simt_eval_tm = 0;
  size_t simt_eval_smem = 0;
  WorkTable simt_eval, tc_eval[kBuckets];
  int tc_eval_tpb[kBuckets] = {0};  // decompress tiles per CTA, per bucket
  int tc_L[kBuckets] = {0};      // deepest network per tensor-core bucket over ALL eval_tc networks (decompress launches)
  int tc_fit_L[kBuckets] = {0};  // deepest network per bucket over the networks whose FIT runs on the tensor core
  // layer-wise tensor-core path (128 < F_PAD <= 256, brief_tc_lw.cu): passes of networks of one (F_PAD, L)
  std::vector<LwPass> lw_fit, lw_eval;
  DevBuf<int> d_lw_fit_tab, d_lw_eval_tab;
  DevBuf<long long> d_lw_fit_base, d_lw_eval_base;
  DevBuf<unsigned char> d_lw_scratch;
  // general sliding-cube samplers (brief_group_set_cube_sampler).  Host-side only: on the device such a network is a
  // RANDOM_POINTS network of cube_count * cube voxels per step whose indices are read from the step's generated index
  // buffer (gen_indices_kernel), so the fit kernels are the ones of the replayed-index mode
  struct CubeCfg { int count = 0; int len[3] = {0, 0, 0}; };
  std::vector<CubeCfg> cubes;
  bool any_cube = false;
  long long idx_total = 0;         // entries of a step's index array (all RANDOM_POINTS networks, finalize)
  DevBuf<long long> d_gen_idx;
};

namespace {

int f4_of(int f) { return (f + 3) & ~3; }
inline bool is_lw(const NetDev& n) { return n.eval_tc && n.F_PAD > 128; }  // layer-wise tensor-core path (brief_tc_lw.cu)
constexpr size_t kLwScratchBudget = (size_t)6 << 30;  // scratch arena of one layer-wise pass

// Device tables of a set of passes: per pass [tile_prefix (n+1) | cta_prefix (n+1) | work_net (n) | tile_count (n) | tile_first (n)]
int upload_lw_tables(std::vector<LwPass>& passes, DevBuf<int>& tab, DevBuf<long long>& base, int num_sms, cudaStream_t st) {
  std::vector<int> t;
  std::vector<long long> b;
  for (auto& p : passes) {
    const int n = (int)p.nets.size();
    p.tile_prefix.assign(1, 0);
    p.cta_prefix.assign(1, 0);
    p.tile_base.clear();
    long long T = 0;
    for (int i = 0; i < n; ++i) {
      p.tile_base.push_back(T);
      T += p.tile_count[i];
      p.tile_prefix.push_back((int)T);
    }
    p.T = T;
    for (int i = 0; i < n; ++i) {  // GEMM CTAs in proportion to the tiles, at least one per network
      long long c = (long long)num_sms * p.tile_count[i] / std::max<long long>(T, 1);
      c = std::max<long long>(1, std::min<long long>(c, p.tile_count[i]));
      p.cta_prefix.push_back(p.cta_prefix.back() + (int)c);
    }
    p.off = (int)t.size();
    t.insert(t.end(), p.tile_prefix.begin(), p.tile_prefix.end());
    t.insert(t.end(), p.cta_prefix.begin(), p.cta_prefix.end());
    t.insert(t.end(), p.nets.begin(), p.nets.end());
    t.insert(t.end(), p.tile_count.begin(), p.tile_count.end());
    t.insert(t.end(), p.tile_first.begin(), p.tile_first.end());
    p.off_base = (int)b.size();
    b.insert(b.end(), p.tile_base.begin(), p.tile_base.end());
  }
  CU(tab.ensure(t.size()));
  CU(base.ensure(b.size()));
  if (!t.empty()) CU(cudaMemcpyAsync(tab.p, t.data(), t.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  if (!b.empty()) CU(cudaMemcpyAsync(base.p, b.data(), b.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));
  return 0;
}
void lw_bind_tables(LwArgs& a, const LwPass& p, const int* tab, const long long* base, bool ctas) {
  const int n = (int)p.nets.size();
  const int* t = tab + p.off;
  a.work_prefix = ctas ? t + (n + 1) : t;
  a.work_net = t + 2 * (n + 1);
  a.tile_count = t + 2 * (n + 1) + n;
  a.tile_first = t + 2 * (n + 1) + 2 * n;
  a.tile_base = base + p.off_base;
  a.n_work = n;
  a.F = p.F;
  a.T = p.T;
  a.max_slices = p.max_slices;
}

// reference packed order (utils/ModelSave.py:32-51) <-> padded device layout ----------------------
void packed_to_dev(const NetDev& n, const float* src, float* dst) {
  std::fill(dst, dst + n.P_dev, 0.f);
  const int f = n.f, F4 = n.F4, in = n.in_dim;
  const float* s = src;
  for (int o = 0; o < f; ++o)
    for (int c = 0; c < in; ++c) dst[dl_W0(n) + 4 * o + c] = *s++;
  for (int o = 0; o < f; ++o) dst[dl_b0(n) + o] = *s++;
  for (int l = 1; l <= n.L - 2; ++l) {
    for (int o = 0; o < f; ++o)
      for (int k = 0; k < f; ++k) dst[dl_W(n, l) + o * F4 + k] = *s++;
    for (int o = 0; o < f; ++o) dst[dl_b(n, l) + o] = *s++;
  }
  for (int k = 0; k < f; ++k) dst[dl_Wlast(n) + k] = *s++;
  dst[dl_blast(n)] = *s++;
}
void dev_to_packed(const NetDev& n, const float* src, float* dst) {
  const int f = n.f, F4 = n.F4, in = n.in_dim;
  float* d = dst;
  for (int o = 0; o < f; ++o)
    for (int c = 0; c < in; ++c) *d++ = src[dl_W0(n) + 4 * o + c];
  for (int o = 0; o < f; ++o) *d++ = src[dl_b0(n) + o];
  for (int l = 1; l <= n.L - 2; ++l) {
    for (int o = 0; o < f; ++o)
      for (int k = 0; k < f; ++k) *d++ = src[dl_W(n, l) + o * F4 + k];
    for (int o = 0; o < f; ++o) *d++ = src[dl_b(n, l) + o];
  }
  for (int k = 0; k < f; ++k) *d++ = src[dl_Wlast(n) + k];
  *d++ = src[dl_blast(n)];
}

int check_net(const BriefGroup* g, int net) {
  if (!g) return fail(BRIEF_ERR_INVALID, "null group");
  if (net < 0 || net >= g->n_nets)
    return fail(BRIEF_ERR_INVALID, "network index %d out of range [0,%d)", net, g->n_nets);
  return 0;
}

int use_device(const BriefGroup* g) {
  CU(cudaSetDevice(g->device));
  return 0;
}

int sync_nets(BriefGroup* g, cudaStream_t st) {
  if (!g->nets_dirty) return 0;
  CU(cudaMemcpyAsync(g->d_nets.p, g->nets.data(), sizeof(NetDev) * g->n_nets, cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));
  g->nets_dirty = false;
  return 0;
}

int upload_tables(DevBuf<int>& buf, const std::vector<int>& tab, cudaStream_t st) {
  CU(buf.ensure(tab.size()));
  if (!tab.empty()) {
    CU(cudaMemcpyAsync(buf.p, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
  }
  return 0;
}

struct TableBuilder {
  std::vector<int> tab;
  void add(WorkTable& w, const std::vector<int>& prefix, const std::vector<int>& nets) {
    w.n = (int)nets.size();
    w.blocks = prefix.empty() ? 0 : prefix.back();
    w.off_prefix = (int)tab.size();
    tab.insert(tab.end(), prefix.begin(), prefix.end());
    w.off_net = (int)tab.size();
    tab.insert(tab.end(), nets.begin(), nets.end());
  }
};

void drop_host_step_graph(HostStepGraph& h) {
  if (h.exec) cudaGraphExecDestroy(h.exec);
  if (h.done) cudaEventDestroy(h.done);
  if (h.copied) cudaEventDestroy(h.copied);
  if (h.h_state) cudaFreeHost(h.h_state);
  if (h.d_state) cudaFree(h.d_state);
  if (h.d_idx) cudaFree(h.d_idx);
  if (h.d_loss) cudaFree(h.d_loss);
  h = HostStepGraph{};
}
void drop_host_step_graphs(BriefGroup* g) {
  for (auto& h : g->host_steps) drop_host_step_graph(h);
}

// static decomposition of the dense-grid evaluation (decompress)
int build_eval_tables(BriefGroup* g, cudaStream_t st) {
  int tm = 128;
  for (auto& n : g->nets) {
    if (n.eval_tc) continue;
    const int t = simt_pick_tm(n.F4, n.L, false, kSmemLimit);
    if (t == 0)
      return fail(BRIEF_ERR_UNSUPPORTED, "features=%d exceed the fp32 kernel's shared-memory budget", n.f);
    tm = std::min(tm, t);
  }
  g->simt_eval_tm = tm;
  g->simt_eval_smem = 0;
  std::vector<int> sp{0}, sn, tp[kBuckets], tn[kBuckets];
  for (int b = 0; b < kBuckets; ++b) tp[b].push_back(0);
  long long bucket_tiles[kBuckets] = {0};
  for (const auto& n : g->nets)
    if (n.eval_tc && !is_lw(n)) bucket_tiles[n.F_PAD / 16] += (n.n_vox + kTcTile - 1) / kTcTile;
  for (int b = 1; b < kBuckets; ++b) g->tc_eval_tpb[b] = tc_eval_tpb(bucket_tiles[b], g->num_sms);
  for (int i = 0; i < g->n_nets; ++i) {
    const NetDev& n = g->nets[i];
    if (is_lw(n)) continue;  // decoded by the layer-wise path, see below
    if (n.eval_tc) {
      const int b = n.F_PAD / 16;
      const long long tiles = (n.n_vox + kTcTile - 1) / kTcTile;
      const long long blocks = (tiles + g->tc_eval_tpb[b] - 1) / g->tc_eval_tpb[b];
      if (tp[b].back() + blocks > 0x7fffffffLL) return fail(BRIEF_ERR_UNSUPPORTED, "group too large for one launch");
      tn[b].push_back(i);
      tp[b].push_back((int)(tp[b].back() + blocks));
      g->tc_L[b] = std::max(g->tc_L[b], n.L);
    } else {
      const long long tiles = (n.n_vox + tm - 1) / tm;
      if (sp.back() + tiles > 0x7fffffffLL) return fail(BRIEF_ERR_UNSUPPORTED, "group too large for one launch");
      sn.push_back(i);
      sp.push_back((int)(sp.back() + tiles));
      g->simt_eval_smem = std::max(g->simt_eval_smem, simt_eval_smem(n.F4, tm));
    }
  }
  TableBuilder tb;
  tb.add(g->simt_eval, sp, sn);
  for (int b = 1; b < kBuckets; ++b) tb.add(g->tc_eval[b], tp[b], tn[b]);
  RC(upload_tables(g->d_eval_tables, tb.tab, st));
  // wide networks: the dense grid of a block is decoded in passes of at most `cap` tiles (two activation buffers)
  g->lw_eval.clear();
  for (int i = 0; i < g->n_nets; ++i) {
    const NetDev& n = g->nets[i];
    if (!is_lw(n)) continue;
    const long long tiles = (n.n_vox + kTcTile - 1) / kTcTile;
    const long long cap = std::max<long long>(g->num_sms, (long long)((size_t)1 << 30) / (long long)lw_tile_bytes_host(n.F_PAD));
    for (long long first = 0; first < tiles; first += cap) {
      LwPass p;
      p.F = n.F_PAD;
      p.L = n.L;
      p.nets = {i};
      p.tile_count = {(int)std::min<long long>(cap, tiles - first)};
      if (first > 0x7fffffffLL) return fail(BRIEF_ERR_UNSUPPORTED, "block too large for the layer-wise decode");
      p.tile_first = {(int)first};
      g->lw_eval.push_back(p);
    }
  }
  return upload_lw_tables(g->lw_eval, g->d_lw_eval_tab, g->d_lw_eval_base, g->num_sms, st);
}

// (re)build the fit decomposition: slices per network, partial arenas, block -> work tables
int finalize(BriefGroup* g, cudaStream_t st) {
  if (!g->work_dirty) return sync_nets(g, st);
  int tm = 128;
  g->any_simt = g->any_tc = false;
  for (auto& n : g->nets) {
    if (n.prec == BRIEF_PREC_F16) { g->any_tc = true; continue; }
    g->any_simt = true;
    const int t = simt_pick_tm(n.F4, n.L, true, kSmemLimit);
    if (t == 0)
      return fail(BRIEF_ERR_UNSUPPORTED, "features=%d layers=%d exceed the fp32 kernel's shared-memory budget", n.f, n.L);
    tm = std::min(tm, t);
  }
  g->simt_fit_tm = tm;
  g->simt_fit_smem = 0;
  drop_host_step_graphs(g);  // the captured launches carry the old decomposition
  for (int b = 0; b < kBuckets; ++b) g->tc_fit_L[b] = 0;
  for (auto& n : g->nets)
    if (n.prec == BRIEF_PREC_F16 && !is_lw(n)) g->tc_fit_L[n.F_PAD / 16] = std::max(g->tc_fit_L[n.F_PAD / 16], n.L);
  // tensor-core networks: one CTA per slice; each width bucket is its own launch, so its slices are sized to fill
  // one wave of CTAs (SMs x resident CTAs of that bucket's kernel) by themselves
  long long tc_tiles[kBuckets] = {0}, tc_tps[kBuckets] = {0};
  for (auto& n : g->nets)
    if (n.prec == BRIEF_PREC_F16 && !is_lw(n)) {
      const long long b = n.mode == BRIEF_SAMPLE_FULL_BLOCK ? n.n_vox : (long long)n.batch;
      tc_tiles[n.F_PAD / 16] += (b + kTcTile - 1) / kTcTile;
    }
  for (int b = 1; b < kBuckets; ++b) {
    if (tc_tiles[b] == 0) continue;
    const long long wave = (long long)g->num_sms * tc_fit_ctas_per_sm(16 * b, g->tc_fit_L[b]);
    tc_tps[b] = std::max<long long>(1, (tc_tiles[b] + wave - 1) / wave);
  }
  long long slice_total = 0, part_total = 0, idx_total = 0;
  std::vector<int> sp{0}, sn, tp[kBuckets], tn[kBuckets], op{0}, on, lw_ids;
  for (int b = 0; b < kBuckets; ++b) tp[b].push_back(0);
  int n_lw = 0;
  for (const auto& n : g->nets) n_lw += (n.prec == BRIEF_PREC_F16 && is_lw(n)) ? 1 : 0;
  for (int i = 0; i < g->n_nets; ++i) {
    NetDev& n = g->nets[i];
    if (n.mode == BRIEF_SAMPLE_FULL_BLOCK) {
      if (n.n_vox > 0x7fffffffLL) return fail(BRIEF_ERR_UNSUPPORTED, "block of %lld voxels exceeds the full-block sampler limit", n.n_vox);
      n.batch = (int)n.n_vox;
      n.idx_off = 0;
    } else {
      n.idx_off = idx_total;
      idx_total += n.batch;
    }
    const bool lw = n.prec == BRIEF_PREC_F16 && is_lw(n);
    const bool tc = n.prec == BRIEF_PREC_F16 && !lw;
    const int tile = (tc || lw) ? kTcTile : tm;
    const long long n_tiles = ((long long)n.batch + tile - 1) / tile;
    const long long max_slices = std::max<long long>(1, (long long)(kPartialCapBytes / ((size_t)n.P_dev * 4)));
    long long tps = (n_tiles + max_slices - 1) / max_slices;
    if (lw) {
      // layer-wise path: a slice is one dW CTA per output half.  PER_NETWORK: 8 slices, fixed by the network alone;
      // FILL_WAVE: enough slices for the wide networks of the group to fill the SMs together (2 CTAs per slice)
      long long want = 8;
      if (g->slicing != BRIEF_SLICING_PER_NETWORK)
        want = std::min<long long>(24, std::max<long long>(8, (g->num_sms + 2 * n_lw - 1) / (2 * std::max(1, n_lw))));
      want = std::min<long long>(want, std::max<long long>(1, (long long)((size_t)(32u << 20) / ((size_t)n.P_dev * 4))));
      tps = std::max<long long>((n_tiles + want - 1) / want, 1);
    }
    if (tc) {
      // PER_NETWORK: the network fills one wave by itself, so its slice boundaries (and with them the fp32
      // summation order of its gradients) do not depend on what else shares the GPU
      const long long wave = (long long)g->num_sms * tc_fit_ctas_per_sm(n.F_PAD, g->tc_fit_L[n.F_PAD / 16]);
      const long long own = g->slicing == BRIEF_SLICING_PER_NETWORK ? std::max<long long>(1, (n_tiles + wave - 1) / wave)
                                                                    : tc_tps[n.F_PAD / 16];
      tps = std::max<long long>(tps, own);
    }
    tps = std::max<long long>(tps, 1);
    n.slice_len = (int)(tps * tile);
    n.n_slices = (int)((n_tiles + tps - 1) / tps);
    n.slice_off = slice_total;
    n.part_off = part_total;
    slice_total += n.n_slices;
    part_total += (long long)n.n_slices * n.P_dev;
    if (lw) {
      lw_ids.push_back(i);
    } else if (tc) {
      const int b = n.F_PAD / 16;
      tn[b].push_back(i);
      tp[b].push_back(tp[b].back() + n.n_slices);
    } else {
      sn.push_back(i);
      sp.push_back(sp.back() + n.n_slices);
      g->simt_fit_smem = std::max(g->simt_fit_smem, simt_fit_smem(n.F4, n.L, tm));
    }
    on.push_back(i);
    op.push_back(op.back() + (n.P_dev + 255) / 256);
  }
  TableBuilder tb;
  tb.add(g->simt_fit, sp, sn);
  for (int b = 1; b < kBuckets; ++b) tb.add(g->tc_fit[b], tp[b], tn[b]);
  tb.add(g->opt, op, on);
  RC(upload_tables(g->d_fit_tables, tb.tab, st));
  // layer-wise passes: networks of one (F_PAD, L), greedily packed under the scratch budget
  g->lw_fit.clear();
  std::sort(lw_ids.begin(), lw_ids.end(), [&](int x, int y) {
    const NetDev &a = g->nets[x], &b = g->nets[y];
    return a.F_PAD != b.F_PAD ? a.F_PAD < b.F_PAD : a.L != b.L ? a.L < b.L : x < y;
  });
  size_t lw_bytes = 0;
  for (int i : lw_ids) {
    const NetDev& n = g->nets[i];
    const int tiles = (int)(((long long)n.batch + kTcTile - 1) / kTcTile);
    LwPass* p = g->lw_fit.empty() ? nullptr : &g->lw_fit.back();
    long long cur = 0;
    if (p) for (int c : p->tile_count) cur += c;
    if (!p || p->F != n.F_PAD || p->L != n.L || (int)p->nets.size() >= g->num_sms ||
        lw_scratch_bytes(n.F_PAD, n.L, cur + tiles) > kLwScratchBudget) {
      g->lw_fit.emplace_back();
      p = &g->lw_fit.back();
      p->F = n.F_PAD;
      p->L = n.L;
      cur = 0;
    }
    p->nets.push_back(i);
    p->tile_count.push_back(tiles);
    p->tile_first.push_back(0);
    p->max_slices = std::max(p->max_slices, n.n_slices);
    lw_bytes = std::max(lw_bytes, lw_scratch_bytes(n.F_PAD, n.L, cur + tiles));
  }
  RC(upload_lw_tables(g->lw_fit, g->d_lw_fit_tab, g->d_lw_fit_base, g->num_sms, st));
  for (const auto& p : g->lw_eval) lw_bytes = std::max(lw_bytes, lw_eval_scratch_bytes(p.F, p.T));
  if (lw_bytes > 0) CU(g->d_lw_scratch.ensure(lw_bytes));
  const size_t part_before = g->d_partials.n;
  CU(g->d_partials.ensure((size_t)part_total));
  // the layer-wise kernels write only the real entries of a slot: the pads of the padded device layout must be zero
  if (!lw_ids.empty() || g->d_partials.n != part_before) CU(cudaMemsetAsync(g->d_partials.p, 0, g->d_partials.n * sizeof(float), st));
  // wide tensor-core buckets: one activation stash + dW scratch per SM (buckets launch one after the other and share it)
  size_t stash_total = 0;
  g->stash_stride = 0;
  for (int b = 5; b < kBuckets; ++b)
    if (g->tc_fit[b].blocks > 0) g->stash_stride = std::max(g->stash_stride, tc_fit_stash_bytes(16 * b, g->tc_fit_L[b]));
  if (g->stash_stride > 0) {
    // indexed by %smid, whose range is [0, %nsmid): ask the device instead of assuming it equals the SM count
    if (g->nsmid == 0) CU(query_nsmid(&g->nsmid, st));
    if (g->nsmid < g->num_sms || g->nsmid > 4096)
      return fail(BRIEF_ERR_UNSUPPORTED, "device reports %%nsmid = %d for %d SMs", g->nsmid, g->num_sms);
    stash_total = (size_t)g->nsmid * g->stash_stride;
  }
  if (stash_total > 0) CU(g->d_stash.ensure(stash_total));
  CU(g->d_loss_partials.ensure((size_t)slice_total));
  CU(g->d_loss_scratch.ensure((size_t)g->n_nets));
  g->idx_total = idx_total;
  g->any_cube = false;
  for (const auto& c : g->cubes) g->any_cube |= c.count > 0;
  if (g->any_cube) CU(g->d_gen_idx.ensure((size_t)idx_total));
  g->nets_dirty = true;
  RC(sync_nets(g, st));
  g->work_dirty = false;
  return 0;
}

int ensure_wpack(BriefGroup* g, cudaStream_t st) {
  if (!g->wpack_dirty || g->total_wpack == 0) { g->wpack_dirty = false; return 0; }
  RC(sync_nets(g, st));
  LAUNCH(launch_pack(g->d_nets.p, g->n_nets, g->d_params.p, g->d_wpack.p, st));
  g->wpack_dirty = false;
  return 0;
}

int launch_lw_fit(BriefGroup* g, const int64_t* dev_idx, uint64_t seed, uint64_t step, cudaStream_t st, const StepState* state);

// A step's explicit index array (layout: NetDev::idx_off) for a group that holds sliding-cube samplers: cube networks
// expand the cubes their Philox stream draws, point networks get the stream they would have drawn on chip.
int generate_step_indices(BriefGroup* g, uint64_t seed, uint64_t step, cudaStream_t st, const StepState* state) {
  for (int i = 0; i < g->n_nets; ++i) {
    const NetDev& n = g->nets[i];
    if (n.mode != BRIEF_SAMPLE_RANDOM_POINTS) continue;
    const BriefGroup::CubeCfg& c = g->cubes[i];
    long long* out = g->d_gen_idx.p + n.idx_off;
    if (c.count == 0) {
      LAUNCH(launch_gen_indices(seed, step, state, n.stream_id, n.batch, n.n_vox, n.h, n.w, 1, 1, 0, nullptr, out, st));
    } else {
      const long long pop = (long long)(n.d - c.len[0] + 1) * (n.h - c.len[1] + 1) * (n.w - c.len[2] + 1);
      const long long cube_vox = (long long)c.len[0] * c.len[1] * c.len[2];
      LAUNCH(launch_gen_indices(seed, step, state, n.stream_id, n.batch, pop, n.h, n.w, c.len[1], c.len[2], cube_vox, nullptr, out, st));
    }
  }
  return 0;
}

int launch_fit_kernels(BriefGroup* g, const int64_t* dev_idx, uint64_t seed, uint64_t step, cudaStream_t st,
                       const StepState* state = nullptr) {
  if (!dev_idx && g->any_cube) {
    RC(generate_step_indices(g, seed, step, st, state));
    dev_idx = reinterpret_cast<const int64_t*>(g->d_gen_idx.p);
  }
  FitArgs a{};
  a.nets = g->d_nets.p;
  a.params = g->d_params.p;
  a.axes = g->d_axes.p;
  a.idx = reinterpret_cast<const long long*>(dev_idx);
  a.seed = seed;
  a.step = step;
  a.partials = g->d_partials.p;
  a.loss_partials = g->d_loss_partials.p;
  a.wpack = g->d_wpack.p;
  a.stash = g->d_stash.p;
  a.stash_stride = g->stash_stride;
  a.stash_slots = g->nsmid;
  a.state = state;
  if (g->simt_fit.blocks > 0) {
    a.work_prefix = g->d_fit_tables.p + g->simt_fit.off_prefix;
    a.work_net = g->d_fit_tables.p + g->simt_fit.off_net;
    a.n_work = g->simt_fit.n;
    a.TM = g->simt_fit_tm;
    LAUNCH(launch_simt_fit(a, g->simt_fit.blocks, g->simt_fit_smem, st));
  }
  for (int b = 1; b < kBuckets; ++b) {
    if (g->tc_fit[b].blocks == 0) continue;
    a.work_prefix = g->d_fit_tables.p + g->tc_fit[b].off_prefix;
    a.work_net = g->d_fit_tables.p + g->tc_fit[b].off_net;
    a.n_work = g->tc_fit[b].n;
    a.TM = kTcTile;
    LAUNCH(launch_tc_fit(a, 16 * b, g->tc_fit_L[b], g->tc_fit[b].blocks, st));
  }
  if (!g->lw_fit.empty()) RC(launch_lw_fit(g, dev_idx, seed, step, st, state));
  return 0;
}

// learning rate of 1-based step t under MultiStepLR (utils/misc.py:187-188): scheduler.step() runs after
// optimizer.step(), so step t uses lr0 * gamma^(#milestones <= t-1), accumulated by chained double multiplications
double lr_at(const BriefOptConfig* cfg, int64_t t) {
  double lr = cfg->lr;
  for (int i = 0; i < cfg->n_milestones; ++i)
    if (cfg->milestones[i] <= t - 1) lr *= (double)cfg->gamma;
  return lr;
}
// torch: bias_correction = 1 - beta1 ** step ; clr = lr / bias_correction   (python doubles)
void step_scalars(int kind, double lr, double b1, double b2, long long t, float* neg_clr, float* bc2_sqrt) {
  if (kind == BRIEF_OPT_SGD) {
    *neg_clr = (float)(-lr);
    *bc2_sqrt = 1.f;
  } else {
    *neg_clr = (float)(-(lr / (1.0 - std::pow(b1, (double)t))));
    *bc2_sqrt = (float)std::sqrt(1.0 - std::pow(b2, (double)t));
  }
}

// one training step of the wide networks, layer by layer (brief_tc_lw.cu)
int launch_lw_fit(BriefGroup* g, const int64_t* dev_idx, uint64_t seed, uint64_t step, cudaStream_t st, const StepState* state) {
  for (const LwPass& p : g->lw_fit) {
    const int NH = p.L - 2, n = (int)p.nets.size();
    LwArgs a{};
    a.nets = g->d_nets.p;
    a.scratch = g->d_lw_scratch.p;
    a.wpack = g->d_wpack.p;
    a.axes = g->d_axes.p;
    a.idx = reinterpret_cast<const long long*>(dev_idx);
    a.seed = seed;
    a.step = step;
    a.state = state;
    a.partials = g->d_partials.p;
    a.loss_partials = g->d_loss_partials.p;
    auto off = [&](int what, int j) { return lw_fit_offset(p.F, p.L, p.T, what, j); };
    lw_bind_tables(a, p, g->d_lw_fit_tab.p, g->d_lw_fit_base.p, false);
    LAUNCH(launch_lw_sample_l0(a, (int)p.T, st));
    for (int j = 1; j <= NH; ++j) {  // theta_j = ACT_{j-1} W'_j^T
      lw_bind_tables(a, p, g->d_lw_fit_tab.p, g->d_lw_fit_base.p, true);
      a.layer = j - 1;
      a.in_off = off(0, j - 1);
      a.out_off = off(0, j);
      a.out2_off = off(1, j);
      LAUNCH(launch_lw_gemm(a, 0, p.ctas(), st));
    }
    lw_bind_tables(a, p, g->d_lw_fit_tab.p, g->d_lw_fit_base.p, false);
    a.in_off = off(0, NH);
    a.out_off = off(2, 0);
    LAUNCH(launch_lw_last(a, (int)p.T, st));
    a.kind = 2;  // dWlast, dblast = ACT_NH^T [dy']
    a.nb = 16;
    a.in_off = off(0, NH);
    a.in2_off = off(4, 0);
    LAUNCH(launch_lw_dw(a, n, st));
    int cur = 0;
    for (int l = NH; l >= 1; --l) {
      a.kind = 0;  // dW_l = DZ_l^T ACT_{l-1}
      a.nb = p.F;
      a.layer = l;
      a.in_off = off(2, cur);
      a.in2_off = off(0, l - 1);
      LAUNCH(launch_lw_dw(a, n, st));
      lw_bind_tables(a, p, g->d_lw_fit_tab.p, g->d_lw_fit_base.p, true);
      a.layer = l - 1;  // dz_{l-1} = (DZ_l W'_l) * COS_{l-1}
      a.in_off = off(2, cur);
      a.in2_off = off(1, l - 1);
      a.out_off = off(2, cur ^ 1);
      LAUNCH(launch_lw_gemm(a, 2, p.ctas(), st));
      lw_bind_tables(a, p, g->d_lw_fit_tab.p, g->d_lw_fit_base.p, false);
      cur ^= 1;
    }
    a.kind = 1;  // dW0, db0 = DZ_0^T [x_hi 1 x_lo ...]
    a.nb = 16;
    a.in_off = off(2, cur);
    a.in2_off = off(3, 0);
    LAUNCH(launch_lw_dw(a, n, st));
    LAUNCH(launch_lw_loss_reduce(a, n, st));
  }
  return 0;
}

// forward of one pass of the wide networks' decode (dense grid or explicit coordinates)
int launch_lw_eval(BriefGroup* g, const LwPass& p, const int* tab, const long long* base, const float* coords, long long n_coords,
                   float* out_f32, float* layers_out, void* const* out_ptrs, int out_dtype, cudaStream_t st) {
  const int NH = p.L - 2;
  LwArgs a{};
  a.nets = g->d_nets.p;
  a.scratch = g->d_lw_scratch.p;
  a.wpack = g->d_wpack.p;
  a.axes = g->d_axes.p;
  a.eval = 1;
  a.coords = coords;
  a.n_coords = n_coords;
  a.out_f32 = out_f32;
  a.layers_out = layers_out;
  a.out_ptrs = out_ptrs;
  a.out_dtype = out_dtype;
  const size_t buf = (size_t)p.T * lw_tile_bytes_host(p.F);
  lw_bind_tables(a, p, tab, base, false);
  LAUNCH(launch_lw_sample_l0(a, (int)p.T, st));  // -> buffer 0
  for (int j = 1; j <= NH; ++j) {
    lw_bind_tables(a, p, tab, base, true);
    a.layer = j - 1;
    a.in_off = (size_t)((j - 1) & 1) * buf;
    a.out_off = (size_t)(j & 1) * buf;
    LAUNCH(launch_lw_gemm(a, 1, p.ctas(), st));
  }
  lw_bind_tables(a, p, tab, base, false);
  a.in_off = (size_t)(NH & 1) * buf;
  LAUNCH(launch_lw_last(a, (int)p.T, st));
  return 0;
}

int launch_opt_kernel(BriefGroup* g, bool from_partials, bool apply, int kind, double lr, double b1, double b2,
                      double eps, long long t, float* loss_out, cudaStream_t st, const StepState* state = nullptr) {
  OptArgs o{};
  o.state = state;
  o.nets = g->d_nets.p;
  o.blk_prefix = g->d_fit_tables.p + g->opt.off_prefix;
  o.n_nets = g->n_nets;
  o.params = g->d_params.p;
  o.grads = g->d_grads.p;
  o.m = g->d_m.p;
  o.v = g->d_v.p;
  o.partials = from_partials ? g->d_partials.p : nullptr;
  o.loss_partials = g->d_loss_partials.p;
  o.loss_out = loss_out;
  o.kind = kind;
  o.apply = apply ? 1 : 0;
  if (apply) {
    step_scalars(kind, lr, b1, b2, t, &o.neg_clr, &o.bc2_sqrt);
    if (kind != BRIEF_OPT_SGD) {
      o.w1 = (float)(1.0 - b1);
      o.beta2 = (float)b2;
      o.w2 = (float)(1.0 - b2);
      o.eps = (float)eps;
    }
  }
  // the optimiser refreshes the fp16 operand image itself (image_scatter) — unless the image is already stale, in
  // which case the next ensure_wpack() rebuilds it from the parameters anyway
  if (apply && g->total_wpack > 0 && !g->wpack_dirty) o.wpack = g->d_wpack.p;
  LAUNCH(launch_opt(o, g->opt.blocks, st));
  if (apply && !o.wpack) g->wpack_dirty = true;
  return 0;
}

int check_bound(const BriefGroup* g) {
  for (int i = 0; i < g->n_nets; ++i)
    if (!g->nets[i].bound) return fail(BRIEF_ERR_STATE, "network %d has no volume bound (brief_group_bind_volume)", i);
  return 0;
}

void linspace_host(float lo, float hi, int n, float* out) {
  // torch CPU linspace, scalar form: step = (end-start)/(steps-1); first half from start, second from end
  if (n == 1) { out[0] = lo; return; }
  const float step = (hi - lo) / (float)(n - 1);
  const int half = n / 2;
  // torch's vectorised CPU kernel contracts the multiply-add (AVX2/AVX-512 fmadd), so fmaf reproduces it
  for (int i = 0; i < n; ++i)
    out[i] = i < half ? std::fmaf(step, (float)i, lo) : std::fmaf(-step, (float)(n - i - 1), hi);
}

}  // namespace

extern "C" {

int brief_abi_version(void) { return BRIEF_B200_ABI_VERSION; }
const char* brief_last_error(void) { return g_err.c_str(); }
int64_t brief_launch_count(void) { return g_launches.load(); }
void brief_reset_launch_count(void) { g_launches.store(0); }

int brief_device_count(int* out_count) {
  if (!out_count) return fail(BRIEF_ERR_INVALID, "null out_count");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *out_count = 0;
    return fail(BRIEF_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  }
  *out_count = n;
  return 0;
}

int brief_linspace(float lo, float hi, int32_t n, float* host_out) {
  if (n < 1 || !host_out) return fail(BRIEF_ERR_INVALID, "brief_linspace: n=%d", n);
  linspace_host(lo, hi, n, host_out);
  return 0;
}

int brief_group_create(const BriefNetDesc* descs, int32_t n_nets, int32_t device, int32_t precision,
                       BriefGroup** out) {
  if (!descs || n_nets < 1 || !out) return fail(BRIEF_ERR_INVALID, "brief_group_create: bad arguments");
  if (precision < 0 || precision > 2) return fail(BRIEF_ERR_INVALID, "unknown precision %d", precision);
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(BRIEF_ERR_CUDA, "CUDA device %d not available (%d devices)", device, ndev);
  CU(cudaSetDevice(device));
  BriefGroup* g = new BriefGroup();
  g->device = device;
  cudaDeviceGetAttribute(&g->num_sms, cudaDevAttrMultiProcessorCount, device);
  if (g->num_sms < 1) g->num_sms = 148;
  g->n_nets = n_nets;
  g->nets.resize(n_nets);
  g->cubes.assign(n_nets, BriefGroup::CubeCfg{});
  auto bail = [&](int rc) { brief_group_destroy(g); return rc; };
  for (int i = 0; i < n_nets; ++i) {
    const BriefNetDesc& d = descs[i];
    if (d.coords_channel != 2 && d.coords_channel != 3)
      return bail(fail(BRIEF_ERR_UNSUPPORTED, "net %d: coords_channel=%d (2 or 3 supported)", i, d.coords_channel));
    if (d.data_channel != 1)
      return bail(fail(BRIEF_ERR_UNSUPPORTED, "net %d: data_channel=%d (the fused kernels support 1)", i, d.data_channel));
    if (d.features < 1 || d.layers < 2 || d.layers > BRIEF_MAX_LAYERS)
      return bail(fail(BRIEF_ERR_INVALID, "net %d: features=%d layers=%d", i, d.features, d.layers));
    if (d.dims[0] < 1 || d.dims[1] < 1 || d.dims[2] < 1 || (d.coords_channel == 2 && d.dims[0] != 1))
      return bail(fail(BRIEF_ERR_INVALID, "net %d: dims=(%d,%d,%d)", i, d.dims[0], d.dims[1], d.dims[2]));
    NetDev n{};
    n.in_dim = d.coords_channel;
    n.out_dim = d.data_channel;
    n.f = d.features;
    n.L = d.layers;
    n.F4 = f4_of(d.features);
    n.w0 = d.w0;
    n.wh = d.w_hidden;
    n.d = d.dims[0];
    n.h = d.dims[1];
    n.w = d.dims[2];
    n.n_vox = (long long)n.d * n.h * n.w;
    n.P_dev = dl_total(n.F4, n.L);
    n.P_ref = n.in_dim * n.f + n.f + (n.L - 2) * (n.f * n.f + n.f) + n.f * n.out_dim + n.out_dim;
    n.param_off = g->total_P;
    g->total_P += n.P_dev;
    n.axis_off = (int)g->total_axis;
    g->total_axis += n.d + n.h + n.w;
    if (g->total_axis > 0x7fffffffLL) return bail(fail(BRIEF_ERR_UNSUPPORTED, "axis table arena too large"));
    const bool tc_ok = tc_supported(n.f, n.L, n.in_dim, n.out_dim);
    if (precision == BRIEF_PREC_F16 && !tc_ok)
      return bail(fail(BRIEF_ERR_UNSUPPORTED,
                       "net %d: features=%d layers=%d is outside the fused tcgen05 kernel's TMEM/SMEM budget; "
                       "use BRIEF_PREC_FP32 or BRIEF_PREC_AUTO", i, n.f, n.L));
    n.prec = (precision == BRIEF_PREC_FP32 || !tc_ok) ? BRIEF_PREC_FP32 : BRIEF_PREC_F16;
    // forward / decompress have a wider tensor-core envelope than the fit (no activation ring): AUTO uses it
    n.eval_tc = (n.prec == BRIEF_PREC_F16 ||
                 (precision == BRIEF_PREC_AUTO && (tc_eval_supported(n.f, n.L, n.in_dim, n.out_dim) ||
                                                   tc_lw_supported(n.f, n.L, n.in_dim, n.out_dim)))) ? 1 : 0;
    if (n.eval_tc) {
      n.F_PAD = tc_fpad(n.f);
      n.wpack_off = (long long)g->total_wpack;
      g->total_wpack += (tc_wpack_bytes(n.F_PAD, n.L) + 127) & ~(size_t)127;
    }
    n.lo = 0.f; n.hi = 100.f; n.vmin = 0.f; n.vmax = 1.f;
    n.dn_lo = 0.f; n.dn_hi = 100.f; n.dn_vmin = 0.f; n.dn_vmax = 1.f; n.dn_range = 1.f;
    // main.py:332-334: whole-block sampling only for blocks of at most 80^3 voxels
    n.stream_id = (unsigned int)i;
    n.mode = n.n_vox <= 80LL * 80 * 80 ? BRIEF_SAMPLE_FULL_BLOCK : BRIEF_SAMPLE_RANDOM_POINTS;
    n.batch = n.mode == BRIEF_SAMPLE_FULL_BLOCK ? (int)n.n_vox : 100000;
    g->nets[i] = n;
  }
  auto cu_bail = [&](cudaError_t e, const char* what) {
    return bail(fail(BRIEF_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e)));
  };
  cudaError_t e;
  if ((e = g->d_nets.ensure(n_nets)) != cudaSuccess) return cu_bail(e, "alloc nets");
  if ((e = g->d_params.ensure(g->total_P)) != cudaSuccess) return cu_bail(e, "alloc params");
  if ((e = g->d_grads.ensure(g->total_P)) != cudaSuccess) return cu_bail(e, "alloc grads");
  if ((e = g->d_m.ensure(g->total_P)) != cudaSuccess) return cu_bail(e, "alloc m");
  if ((e = g->d_v.ensure(g->total_P)) != cudaSuccess) return cu_bail(e, "alloc v");
  if ((e = g->d_axes.ensure(g->total_axis)) != cudaSuccess) return cu_bail(e, "alloc axes");
  if ((e = g->d_wpack.ensure(g->total_wpack + 128)) != cudaSuccess) return cu_bail(e, "alloc wpack");
  if ((e = g->d_outptrs.ensure(n_nets)) != cudaSuccess) return cu_bail(e, "alloc outptrs");
  const size_t pb = (size_t)g->total_P * sizeof(float);
  if ((e = cudaMemset(g->d_params.p, 0, pb)) != cudaSuccess) return cu_bail(e, "memset");
  if ((e = cudaMemset(g->d_grads.p, 0, pb)) != cudaSuccess) return cu_bail(e, "memset");
  if ((e = cudaMemset(g->d_m.p, 0, pb)) != cudaSuccess) return cu_bail(e, "memset");
  if ((e = cudaMemset(g->d_v.p, 0, pb)) != cudaSuccess) return cu_bail(e, "memset");
  if ((e = cudaMemset(g->d_wpack.p, 0, g->total_wpack + 128)) != cudaSuccess) return cu_bail(e, "memset");
  std::vector<float> axes((size_t)g->total_axis);
  for (auto& n : g->nets) {
    linspace_host(-1.f, 1.f, n.d, axes.data() + n.axis_off);
    linspace_host(-1.f, 1.f, n.h, axes.data() + n.axis_off + n.d);
    linspace_host(-1.f, 1.f, n.w, axes.data() + n.axis_off + n.d + n.h);
  }
  if ((e = cudaMemcpy(g->d_axes.p, axes.data(), axes.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess)
    return cu_bail(e, "upload axes");
  int rc = build_eval_tables(g, 0);
  if (rc) return bail(rc);
  *out = g;
  return 0;
}

void brief_group_destroy(BriefGroup* g) {
  if (!g) return;
  cudaSetDevice(g->device);
  drop_host_step_graphs(g);
  if (g->capture_stream) cudaStreamDestroy(g->capture_stream);
  if (g->copy_stream) cudaStreamDestroy(g->copy_stream);
  g->d_nets.release(); g->d_params.release(); g->d_grads.release(); g->d_m.release(); g->d_v.release();
  g->d_axes.release(); g->d_partials.release(); g->d_loss_partials.release(); g->d_loss_scratch.release();
  g->d_lw_fit_tab.release(); g->d_lw_eval_tab.release(); g->d_lw_fit_base.release(); g->d_lw_eval_base.release();
  g->d_lw_scratch.release(); g->d_gen_idx.release();
  g->d_wpack.release(); g->d_stash.release(); g->d_fit_tables.release(); g->d_eval_tables.release(); g->d_outptrs.release();
  delete g;
}

int brief_group_num_nets(const BriefGroup* g) { return g ? g->n_nets : fail(BRIEF_ERR_INVALID, "null group"); }
int brief_group_param_count(const BriefGroup* g, int32_t net) {
  RC(check_net(g, net));
  return g->nets[net].P_ref;
}
int brief_group_precision(const BriefGroup* g, int32_t net) {
  RC(check_net(g, net));
  return g->nets[net].prec;
}
int brief_group_batch(const BriefGroup* g, int32_t net) {
  RC(check_net(g, net));
  return g->nets[net].batch;
}

static int arena_put(BriefGroup* g, int net, float* arena, const float* host_packed, cudaStream_t st) {
  RC(check_net(g, net));
  if (!host_packed) return fail(BRIEF_ERR_INVALID, "null host pointer");
  RC(use_device(g));
  const NetDev& n = g->nets[net];
  std::vector<float> tmp(n.P_dev);
  packed_to_dev(n, host_packed, tmp.data());
  CU(cudaMemcpyAsync(arena + n.param_off, tmp.data(), sizeof(float) * n.P_dev, cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));
  return 0;
}
static int arena_get(BriefGroup* g, int net, const float* arena, float* host_packed, cudaStream_t st) {
  RC(check_net(g, net));
  if (!host_packed) return fail(BRIEF_ERR_INVALID, "null host pointer");
  RC(use_device(g));
  const NetDev& n = g->nets[net];
  std::vector<float> tmp(n.P_dev);
  CU(cudaMemcpyAsync(tmp.data(), arena + n.param_off, sizeof(float) * n.P_dev, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  dev_to_packed(n, tmp.data(), host_packed);
  return 0;
}

int brief_group_set_params(BriefGroup* g, int32_t net, const float* host_packed, void* stream) {
  RC(arena_put(g, net, g ? g->d_params.p : nullptr, host_packed, (cudaStream_t)stream));
  g->wpack_dirty = true;
  return 0;
}
int brief_group_get_params(BriefGroup* g, int32_t net, float* host_packed, void* stream) {
  return arena_get(g, net, g ? g->d_params.p : nullptr, host_packed, (cudaStream_t)stream);
}
int brief_group_get_grads(BriefGroup* g, int32_t net, float* host_packed, void* stream) {
  return arena_get(g, net, g ? g->d_grads.p : nullptr, host_packed, (cudaStream_t)stream);
}
int brief_group_set_grads(BriefGroup* g, int32_t net, const float* host_packed, void* stream) {
  return arena_put(g, net, g ? g->d_grads.p : nullptr, host_packed, (cudaStream_t)stream);
}
int brief_group_get_opt_state(BriefGroup* g, int32_t net, float* host_m, float* host_v, void* stream) {
  RC(arena_get(g, net, g ? g->d_m.p : nullptr, host_m, (cudaStream_t)stream));
  return arena_get(g, net, g->d_v.p, host_v, (cudaStream_t)stream);
}
int brief_group_reset_opt_state(BriefGroup* g, void* stream) {
  if (!g) return fail(BRIEF_ERR_INVALID, "null group");
  RC(use_device(g));
  CU(cudaMemsetAsync(g->d_m.p, 0, sizeof(float) * g->total_P, (cudaStream_t)stream));
  CU(cudaMemsetAsync(g->d_v.p, 0, sizeof(float) * g->total_P, (cudaStream_t)stream));
  return 0;
}

int brief_group_set_axes(BriefGroup* g, int32_t net, const float* host_d, const float* host_h, const float* host_w,
                         void* stream) {
  RC(check_net(g, net));
  if (!host_d || !host_h || !host_w) return fail(BRIEF_ERR_INVALID, "null axis table");
  RC(use_device(g));
  const NetDev& n = g->nets[net];
  std::vector<float> tmp((size_t)n.d + n.h + n.w);
  std::copy(host_d, host_d + n.d, tmp.begin());
  std::copy(host_h, host_h + n.h, tmp.begin() + n.d);
  std::copy(host_w, host_w + n.w, tmp.begin() + n.d + n.h);
  CU(cudaMemcpyAsync(g->d_axes.p + n.axis_off, tmp.data(), tmp.size() * sizeof(float), cudaMemcpyHostToDevice,
                     (cudaStream_t)stream));
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

int brief_group_bind_volume(BriefGroup* g, int32_t net, const void* dev_raw, int32_t dtype, float vmin, float vmax,
                            float lo, float hi, const float* dev_weight, const BriefWeightRule* rules,
                            int32_t n_rules, float tau) {
  RC(check_net(g, net));
  if (!dev_raw) return fail(BRIEF_ERR_INVALID, "null volume pointer");
  if (dtype < 0 || dtype > 2) return fail(BRIEF_ERR_INVALID, "unknown dtype %d", dtype);
  if (n_rules < 0 || n_rules > BRIEF_MAX_RULES) return fail(BRIEF_ERR_UNSUPPORTED, "at most %d weight rules", BRIEF_MAX_RULES);
  if (n_rules > 0 && !rules) return fail(BRIEF_ERR_INVALID, "null rules");
  NetDev& n = g->nets[net];
  n.raw = dev_raw;
  n.dtype = dtype;
  n.weight = dev_weight;
  n.vmin = vmin; n.vmax = vmax; n.lo = lo; n.hi = hi;
  n.dn_vmin = vmin; n.dn_vmax = vmax; n.dn_lo = lo; n.dn_hi = hi;
  n.dn_range = (float)((double)vmax - (double)vmin);
  n.n_rules = n_rules;
  for (int r = 0; r < n_rules; ++r) { n.rule_lo[r] = rules[r].lo; n.rule_hi[r] = rules[r].hi; n.rule_s[r] = rules[r].scale; }
  n.tau = tau;
  n.bound = 1;
  g->nets_dirty = true;
  return 0;
}

int brief_group_set_denorm(BriefGroup* g, int32_t net, float vmin, float vmax, float lo, float hi) {
  RC(check_net(g, net));
  NetDev& n = g->nets[net];
  n.dn_vmin = vmin; n.dn_vmax = vmax; n.dn_lo = lo; n.dn_hi = hi;
  n.dn_range = (float)((double)vmax - (double)vmin);
  g->nets_dirty = true;
  return 0;
}

int brief_group_set_sampler(BriefGroup* g, int32_t net, int32_t mode, int32_t batch) {
  RC(check_net(g, net));
  if (mode != BRIEF_SAMPLE_FULL_BLOCK && mode != BRIEF_SAMPLE_RANDOM_POINTS)
    return fail(BRIEF_ERR_INVALID, "unknown sampler mode %d", mode);
  if (mode == BRIEF_SAMPLE_RANDOM_POINTS && batch < 1) return fail(BRIEF_ERR_INVALID, "batch=%d", batch);
  NetDev& n = g->nets[net];
  n.mode = mode;
  n.batch = mode == BRIEF_SAMPLE_FULL_BLOCK ? (int)std::min<long long>(n.n_vox, 0x7fffffffLL) : batch;
  g->cubes[net] = BriefGroup::CubeCfg{};
  g->work_dirty = true;
  g->nets_dirty = true;
  return 0;
}

int brief_group_set_cube_sampler(BriefGroup* g, int32_t net, int32_t cube_count, const int32_t* host_cube_len) {
  RC(check_net(g, net));
  if (!host_cube_len || cube_count < 1) return fail(BRIEF_ERR_INVALID, "cube_count=%d", cube_count);
  NetDev& n = g->nets[net];
  const int dims[3] = {n.d, n.h, n.w};
  BriefGroup::CubeCfg c;
  long long vox = 1, pop = 1;
  for (int k = 0; k < 3; ++k) {
    if (host_cube_len[k] < 1) return fail(BRIEF_ERR_INVALID, "cube_len[%d]=%d", k, host_cube_len[k]);
    c.len[k] = std::min(host_cube_len[k], dims[k]);  // main.py:49-50
    vox *= c.len[k];
    pop *= dims[k] - c.len[k] + 1;
  }
  if (pop == 1 && cube_count == 1) return brief_group_set_sampler(g, net, BRIEF_SAMPLE_FULL_BLOCK, 0);
  if (vox * cube_count > 0x7fffffffLL)
    return fail(BRIEF_ERR_UNSUPPORTED, "%d cubes of %lld voxels exceed the per-step sample limit", cube_count, vox);
  if (pop > 0xffffffffLL) return fail(BRIEF_ERR_UNSUPPORTED, "cube population %lld exceeds the sampler stream's range", pop);
  c.count = cube_count;
  g->cubes[net] = c;
  n.mode = BRIEF_SAMPLE_RANDOM_POINTS;
  n.batch = (int)(vox * cube_count);
  g->work_dirty = true;
  g->nets_dirty = true;
  return 0;
}

int brief_cube_indices(BriefGroup* g, int32_t net, const int64_t* dev_cube_ids, uint64_t seed, uint64_t step, int64_t* dev_out,
                       void* stream) {
  RC(check_net(g, net));
  if (!dev_out) return fail(BRIEF_ERR_INVALID, "null output");
  const BriefGroup::CubeCfg& c = g->cubes[net];
  if (c.count == 0) return fail(BRIEF_ERR_STATE, "net %d has no sliding-cube sampler (brief_group_set_cube_sampler)", net);
  RC(use_device(g));
  const NetDev& n = g->nets[net];
  const long long pop = (long long)(n.d - c.len[0] + 1) * (n.h - c.len[1] + 1) * (n.w - c.len[2] + 1);
  LAUNCH(launch_gen_indices(seed, step, nullptr, n.stream_id, n.batch, pop, n.h, n.w, c.len[1], c.len[2],
                            (long long)c.len[0] * c.len[1] * c.len[2], reinterpret_cast<const long long*>(dev_cube_ids),
                            reinterpret_cast<long long*>(dev_out), (cudaStream_t)stream));
  return 0;
}

int brief_group_set_stream(BriefGroup* g, int32_t net, uint32_t stream_id) {
  RC(check_net(g, net));
  g->nets[net].stream_id = stream_id;
  g->nets_dirty = true;
  return 0;
}

int brief_group_set_slicing(BriefGroup* g, int32_t mode) {
  if (!g) return fail(BRIEF_ERR_INVALID, "null group");
  if (mode != BRIEF_SLICING_FILL_WAVE && mode != BRIEF_SLICING_PER_NETWORK) return fail(BRIEF_ERR_INVALID, "unknown slicing mode %d", mode);
  g->slicing = mode;
  g->work_dirty = true;
  return 0;
}

int brief_fit_step(BriefGroup* g, const int64_t* dev_idx, uint64_t seed, uint64_t step, float* dev_loss, void* stream) {
  if (!g) return fail(BRIEF_ERR_INVALID, "null group");
  cudaStream_t st = (cudaStream_t)stream;
  RC(use_device(g));
  RC(check_bound(g));
  RC(finalize(g, st));
  RC(ensure_wpack(g, st));
  RC(launch_fit_kernels(g, dev_idx, seed, step, st));
  return launch_opt_kernel(g, true, false, 0, 0, 0, 0, 0, 0, dev_loss ? dev_loss : g->d_loss_scratch.p, st);
}

int brief_fit_kernels(BriefGroup* g, const int64_t* dev_idx, uint64_t seed, uint64_t step, void* stream) {
  if (!g) return fail(BRIEF_ERR_INVALID, "null group");
  cudaStream_t st = (cudaStream_t)stream;
  RC(use_device(g));
  RC(check_bound(g));
  RC(finalize(g, st));
  RC(ensure_wpack(g, st));
  return launch_fit_kernels(g, dev_idx, seed, step, st);
}

int brief_opt_step(BriefGroup* g, int32_t kind, float lr, float beta1, float beta2, float eps, int64_t t, void* stream) {
  if (!g) return fail(BRIEF_ERR_INVALID, "null group");
  if (kind < 0 || kind > 2) return fail(BRIEF_ERR_INVALID, "unknown optimiser %d", kind);
  if (t < 1) return fail(BRIEF_ERR_INVALID, "optimiser step count t=%lld must be >= 1", (long long)t);
  cudaStream_t st = (cudaStream_t)stream;
  RC(use_device(g));
  RC(finalize(g, st));
  return launch_opt_kernel(g, false, true, kind, lr, beta1, beta2, eps, t, nullptr, st);
}

int brief_fit_run(BriefGroup* g, const BriefOptConfig* cfg, uint64_t seed, int64_t steps_done, int64_t n_steps,
                  float* dev_loss_hist, void* stream) {
  if (!g || !cfg) return fail(BRIEF_ERR_INVALID, "null argument");
  if (cfg->kind < 0 || cfg->kind > 2) return fail(BRIEF_ERR_INVALID, "unknown optimiser %d", cfg->kind);
  if (cfg->n_milestones < 0 || cfg->n_milestones > 8) return fail(BRIEF_ERR_INVALID, "n_milestones=%d", cfg->n_milestones);
  if (steps_done < 0 || n_steps < 0) return fail(BRIEF_ERR_INVALID, "negative step count");
  cudaStream_t st = (cudaStream_t)stream;
  RC(use_device(g));
  RC(check_bound(g));
  RC(finalize(g, st));
  for (int64_t s = 0; s < n_steps; ++s) {
    const int64_t t = steps_done + s + 1;  // 1-based optimiser step
    const double lr = lr_at(cfg, t);
    RC(ensure_wpack(g, st));
    RC(launch_fit_kernels(g, nullptr, seed, (uint64_t)(t - 1), st));
    float* loss_out = dev_loss_hist ? dev_loss_hist + (size_t)s * g->n_nets : g->d_loss_scratch.p;
    RC(launch_opt_kernel(g, true, true, cfg->kind, lr, cfg->beta1, cfg->beta2, cfg->eps, t, loss_out, st));
  }
  return 0;
}

// ---- one training step driven from HOST buffers, replayed as a CUDA graph --------------------------------------------
// kernels + loss read-back of one step (this is what the graph holds); the inputs arrive on the copy stream
static int enqueue_host_step(BriefGroup* g, HostStepGraph& h, long long n_idx, cudaStream_t st, int* kernels) {
  const BriefOptConfig& c = h.cfg;
  const long long before = g_launches.load();
  RC(launch_fit_kernels(g, n_idx > 0 ? reinterpret_cast<const int64_t*>(h.d_idx) : nullptr, h.seed, 0, st, h.d_state));
  RC(launch_opt_kernel(g, true, true, c.kind, c.lr, c.beta1, c.beta2, c.eps, 1, h.d_loss, st, h.d_state));
  *kernels = (int)(g_launches.load() - before);
  if (h.host_loss) CU(cudaMemcpyAsync(h.host_loss, h.d_loss, sizeof(float) * g->n_nets, cudaMemcpyDeviceToHost, st));
  return 0;
}

int brief_fit_step_host(BriefGroup* g, const int64_t* host_idx, const BriefOptConfig* cfg, uint64_t seed, int64_t steps_done,
                        float* host_loss, void* stream) {
  if (!g || !cfg) return fail(BRIEF_ERR_INVALID, "null argument");
  if (cfg->kind < 0 || cfg->kind > 2) return fail(BRIEF_ERR_INVALID, "unknown optimiser %d", cfg->kind);
  if (cfg->n_milestones < 0 || cfg->n_milestones > 8) return fail(BRIEF_ERR_INVALID, "n_milestones=%d", cfg->n_milestones);
  if (steps_done < 0) return fail(BRIEF_ERR_INVALID, "negative step count");
  cudaStream_t st = (cudaStream_t)stream;
  RC(use_device(g));
  RC(check_bound(g));
  RC(finalize(g, st));       // drops the cached graphs when the decomposition changed
  RC(ensure_wpack(g, st));   // before capture: the captured optimiser launch then keeps the operand image current
  long long n_idx = 0;
  for (const auto& n : g->nets)
    if (n.mode == BRIEF_SAMPLE_RANDOM_POINTS) n_idx += n.batch;
  if (!host_idx) n_idx = 0;  // on-device sampler stream (Philox, keyed by seed and step)
  if (!g->capture_stream) CU(cudaStreamCreateWithFlags(&g->capture_stream, cudaStreamNonBlocking));
  if (!g->copy_stream) CU(cudaStreamCreateWithFlags(&g->copy_stream, cudaStreamNonBlocking));
  // slot lookup: same buffers, stream and optimiser constants -> the cached graph is valid
  HostStepGraph* h = nullptr;
  for (auto& c : g->host_steps)
    if (c.d_state && c.host_idx == host_idx && c.host_loss == host_loss && c.stream == st && c.seed == seed &&
        memcmp(&c.cfg, cfg, sizeof(BriefOptConfig)) == 0) { h = &c; break; }
  if (!h) {
    h = &g->host_steps[g->host_step_next];
    g->host_step_next = (g->host_step_next + 1) % kHostStepSlots;
    if (h->done) cudaEventSynchronize(h->done);
    drop_host_step_graph(*h);
    h->host_idx = host_idx; h->host_loss = host_loss; h->stream = st; h->cfg = *cfg; h->seed = seed;
    CU(cudaMallocHost(&h->h_state, sizeof(StepState)));
    CU(cudaMalloc(&h->d_state, sizeof(StepState)));
    CU(cudaMalloc(&h->d_loss, sizeof(float) * g->n_nets));
    if (n_idx > 0) CU(cudaMalloc(&h->d_idx, (size_t)n_idx * sizeof(long long)));
    CU(cudaEventCreateWithFlags(&h->done, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->copied, cudaEventDisableTiming));
    static const bool no_graph = getenv("BRIEF_NO_GRAPH") != nullptr;
    if (!no_graph) {
      // capture the sequence a plain submission would enqueue (relaxed mode: the launchers call cudaFuncSetAttribute), on
      // an internal stream: the caller's stream may be the legacy default stream, which cannot be captured
      cudaGraph_t graph = nullptr;
      CU(cudaStreamSynchronize(st));  // finalize / pack above ran on the caller's stream
      cudaStream_t cs = g->capture_stream;
      cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed);
      if (e != cudaSuccess) return fail(BRIEF_ERR_CUDA, "cudaStreamBeginCapture: %s", cudaGetErrorString(e));
      int rc = enqueue_host_step(g, *h, n_idx, cs, &h->kernels);
      e = cudaStreamEndCapture(cs, &graph);
      if (rc != 0 || e != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return rc ? rc : fail(BRIEF_ERR_CUDA, "capturing the step graph failed: %s", cudaGetErrorString(e));
      }
      g_launches.fetch_add(-(long long)h->kernels);  // captured, not launched
      e = cudaGraphInstantiate(&h->exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) return fail(BRIEF_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    }
  }
  // this step's scalars: sampler stream position and the optimiser's step-dependent factors (host doubles like torch)
  const int64_t t = steps_done + 1;
  CU(cudaEventSynchronize(h->done));  // the slot's previous submission has consumed its device buffers and pinned scalars
  h->h_state->step = (unsigned long long)(t - 1);
  step_scalars(cfg->kind, lr_at(cfg, t), cfg->beta1, cfg->beta2, t, &h->h_state->neg_clr, &h->h_state->bc2_sqrt);
  // inputs cross PCIe on the copy stream, i.e. under the kernels of the step before (the caller enqueues step s while
  // step s-1 runs); the kernels of this step wait for them
  CU(cudaMemcpyAsync(h->d_state, h->h_state, sizeof(StepState), cudaMemcpyHostToDevice, g->copy_stream));
  if (n_idx > 0)
    CU(cudaMemcpyAsync(h->d_idx, host_idx, (size_t)n_idx * sizeof(long long), cudaMemcpyHostToDevice, g->copy_stream));
  CU(cudaEventRecord(h->copied, g->copy_stream));
  CU(cudaStreamWaitEvent(st, h->copied, 0));
  if (h->exec) {
    CU(cudaGraphLaunch(h->exec, st));
    g_launches.fetch_add(h->kernels);
  } else {
    int k = 0;
    RC(enqueue_host_step(g, *h, n_idx, st, &k));
  }
  CU(cudaEventRecord(h->done, st));
  return 0;
}

int brief_block_histogram(const void* dev_raw, int64_t n, int32_t dtype, uint64_t* host_hist, int32_t device, void* stream) {
  if (!dev_raw || !host_hist || n < 0) return fail(BRIEF_ERR_INVALID, "brief_block_histogram: bad arguments");
  if (dtype != BRIEF_U8 && dtype != BRIEF_U16) return fail(BRIEF_ERR_UNSUPPORTED, "brief_block_histogram: uint8 / uint16 blocks only");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail(BRIEF_ERR_CUDA, "CUDA device %d not available", device);
  CU(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const size_t bins = dtype == BRIEF_U8 ? 256 : 65536;
  unsigned long long* d = nullptr;
  CU(cudaMalloc(&d, bins * sizeof(unsigned long long)));
  cudaError_t e = cudaMemsetAsync(d, 0, bins * sizeof(unsigned long long), st);
  if (e == cudaSuccess && n > 0) {
    e = launch_histogram(dev_raw, n, dtype, d, sms, st);
    if (e == cudaSuccess) g_launches.fetch_add(1);
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(host_hist, d, bins * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d);
  if (e != cudaSuccess) return fail(BRIEF_ERR_CUDA, "brief_block_histogram: %s", cudaGetErrorString(e));
  return 0;
}

int brief_forward(BriefGroup* g, int32_t net, const float* dev_coords, int64_t n, float* dev_out, float* dev_layers,
                  void* stream) {
  RC(check_net(g, net));
  if (n == 0) return 0;  // zero coordinates: nothing to do (empty tensors have null pointers)
  if (!dev_coords || !dev_out || n < 0) return fail(BRIEF_ERR_INVALID, "brief_forward: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  RC(use_device(g));
  RC(sync_nets(g, st));
  const NetDev& nd = g->nets[net];
  EvalArgs a{};
  a.nets = g->d_nets.p;
  a.params = g->d_params.p;
  a.axes = g->d_axes.p;
  a.single_net = net;
  a.coords = dev_coords;
  a.n_coords = n;
  a.out_f32 = dev_out;
  a.layers_out = dev_layers;
  a.wpack = g->d_wpack.p;
  if (is_lw(nd)) {  // wide network: layer-wise forward over passes of explicit coordinates
    RC(ensure_wpack(g, st));
    const long long tiles = (n + kTcTile - 1) / kTcTile;
    const long long cap = std::max<long long>(g->num_sms, (long long)((size_t)1 << 30) / (long long)lw_tile_bytes_host(nd.F_PAD));
    std::vector<LwPass> passes;
    for (long long first = 0; first < tiles; first += cap) {
      LwPass p;
      p.F = nd.F_PAD;
      p.L = nd.L;
      p.nets = {net};
      p.tile_count = {(int)std::min<long long>(cap, tiles - first)};
      p.tile_first = {(int)first};
      passes.push_back(p);
    }
    DevBuf<int> tab;
    DevBuf<long long> base;
    int rc = upload_lw_tables(passes, tab, base, g->num_sms, st);
    size_t need = 0;
    for (const auto& p : passes) need = std::max(need, lw_eval_scratch_bytes(p.F, p.T));
    if (!rc && g->d_lw_scratch.ensure(need) != cudaSuccess) rc = fail(BRIEF_ERR_CUDA, "alloc layer-wise scratch");
    for (size_t i = 0; i < passes.size() && !rc; ++i)
      rc = launch_lw_eval(g, passes[i], tab.p, base.p, dev_coords, n, dev_out, dev_layers, nullptr, 0, st);
    cudaStreamSynchronize(st);  // the temporary tables are freed below
    tab.release();
    base.release();
    return rc;
  }
  if (nd.eval_tc) {
    RC(ensure_wpack(g, st));
    const long long tiles = (n + kTcTile - 1) / kTcTile;
    a.tiles_per_block = tc_eval_tpb(tiles, g->num_sms);
    const long long blocks = (tiles + a.tiles_per_block - 1) / a.tiles_per_block;
    a.TM = kTcTile;
    LAUNCH(launch_tc_eval(a, nd.F_PAD, nd.L, (int)blocks, st));
  } else {
    const int tm = simt_pick_tm(nd.F4, nd.L, false, kSmemLimit);
    if (tm == 0) return fail(BRIEF_ERR_UNSUPPORTED, "features=%d exceed the fp32 kernel's shared-memory budget", nd.f);
    a.TM = tm;
    const long long tiles = (n + tm - 1) / tm;
    if (tiles > 0x7fffffffLL) return fail(BRIEF_ERR_UNSUPPORTED, "too many coordinates for one launch");
    LAUNCH(launch_simt_eval(a, (int)tiles, simt_eval_smem(nd.F4, tm), st));
  }
  return 0;
}

int brief_decompress(BriefGroup* g, void* const* host_dev_out, int32_t out_dtype, void* stream) {
  if (!g || !host_dev_out) return fail(BRIEF_ERR_INVALID, "null argument");
  if (out_dtype < 0 || out_dtype > 2) return fail(BRIEF_ERR_INVALID, "unknown dtype %d", out_dtype);
  bool any = false;
  for (int i = 0; i < g->n_nets; ++i) any = any || host_dev_out[i] != nullptr;  // NULL = leave this network out
  if (!any) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  RC(use_device(g));
  RC(sync_nets(g, st));
  CU(cudaMemcpyAsync(g->d_outptrs.p, host_dev_out, sizeof(void*) * g->n_nets, cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));
  EvalArgs a{};
  a.nets = g->d_nets.p;
  a.params = g->d_params.p;
  a.axes = g->d_axes.p;
  a.single_net = -1;
  a.out_ptrs = g->d_outptrs.p;
  a.out_dtype = out_dtype;
  a.wpack = g->d_wpack.p;
  if (g->simt_eval.blocks > 0) {
    a.work_prefix = g->d_eval_tables.p + g->simt_eval.off_prefix;
    a.work_net = g->d_eval_tables.p + g->simt_eval.off_net;
    a.n_work = g->simt_eval.n;
    a.TM = g->simt_eval_tm;
    LAUNCH(launch_simt_eval(a, g->simt_eval.blocks, g->simt_eval_smem, st));
  }
  bool packed = false;
  for (int b = 1; b < kBuckets; ++b) {
    if (g->tc_eval[b].blocks == 0) continue;
    if (!packed) { RC(ensure_wpack(g, st)); packed = true; }
    a.work_prefix = g->d_eval_tables.p + g->tc_eval[b].off_prefix;
    a.work_net = g->d_eval_tables.p + g->tc_eval[b].off_net;
    a.n_work = g->tc_eval[b].n;
    a.TM = kTcTile;
    a.tiles_per_block = g->tc_eval_tpb[b];
    LAUNCH(launch_tc_eval(a, 16 * b, g->tc_L[b], g->tc_eval[b].blocks, st));
  }
  for (const LwPass& p : g->lw_eval) {  // wide networks, layer by layer
    if (host_dev_out[p.nets[0]] == nullptr) continue;
    if (!packed) { RC(ensure_wpack(g, st)); packed = true; }
    size_t need = lw_eval_scratch_bytes(p.F, p.T);
    if (g->d_lw_scratch.n < need) {
      CU(cudaStreamSynchronize(st));
      CU(g->d_lw_scratch.ensure(need));
    }
    RC(launch_lw_eval(g, p, g->d_lw_eval_tab.p, g->d_lw_eval_base.p, nullptr, 0, nullptr, nullptr, g->d_outptrs.p, out_dtype, st));
  }
  return 0;
}

int brief_gather(BriefGroup* g, int32_t net, const int64_t* dev_idx, int64_t batch, float* dev_coords, float* dev_data,
                 float* dev_weight, void* stream) {
  RC(check_net(g, net));
  if (batch < 0) return fail(BRIEF_ERR_INVALID, "batch=%lld", (long long)batch);
  if (!g->nets[net].bound && (dev_data || dev_weight)) return fail(BRIEF_ERR_STATE, "network %d has no volume bound", net);
  if (batch == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  RC(use_device(g));
  RC(sync_nets(g, st));
  LAUNCH(launch_gather(g->d_nets.p, net, g->d_axes.p, reinterpret_cast<const long long*>(dev_idx), batch, dev_coords,
                       dev_data, dev_weight, st));
  return 0;
}

int brief_deblock(void* dev_volume, int32_t depth, int32_t height, int32_t width, const int32_t* host_blocks,
                  int32_t n_blocks, int32_t index_a, int32_t index_b, int32_t thres, int32_t* host_masks, int32_t device,
                  void* stream) {
  if (!host_blocks || n_blocks < 0 || depth < 0 || height < 0 || width < 0)
    return fail(BRIEF_ERR_INVALID, "brief_deblock: bad arguments");
  // seam list (deblock.cpp:244-276): a block lists its left / right / down / up seam for every z of its range unless
  // the seam's flag is up; a flag goes up when the block's seam at z1 is already listed and never comes down again
  std::vector<DeblockBlock> blocks((size_t)n_blocks);
  // listed seams: (l, r, d, u) -> z ranges (a seam is listed for every z of its block: ranges instead of one key per z)
  std::map<std::array<int, 4>, std::vector<std::pair<int, int>>> seen;
  auto listed = [&](const int* c, int z) {
    auto it = seen.find({c[0], c[1], c[2], c[3]});
    if (it == seen.end()) return false;
    for (const auto& zr : it->second)
      if (z >= zr.first && z <= zr.second) return true;
    return false;
  };
  bool flags[4] = {false, false, false, false};
  for (int i = 0; i < n_blocks; ++i) {
    const int32_t* b = host_blocks + 6 * (size_t)i;
    DeblockBlock k{b[0], b[1], b[2], b[3], b[4], b[5], 0};
    if (k.z1 < 0 || k.z2 < k.z1 || k.y1 < 0 || k.y2 < k.y1 || k.x1 < 0 || k.x2 < k.x1 ||
        (dev_volume && (k.z2 >= depth || k.y2 >= height || k.x2 >= width)))
      return fail(BRIEF_ERR_INVALID, "brief_deblock: block %d out of range", i);
    const int cand[4][4] = {{k.x1, k.x1, k.y1, k.y2}, {k.x2, k.x2, k.y1, k.y2}, {k.x1, k.x2, k.y1, k.y1}, {k.x1, k.x2, k.y2, k.y2}};
    for (int s = 0; s < 4; ++s)
      if (listed(cand[s], k.z1)) flags[s] = true;
    for (int s = 0; s < 4; ++s) {
      if (flags[s]) continue;
      k.mask |= 1 << s;
      seen[{cand[s][0], cand[s][1], cand[s][2], cand[s][3]}].push_back({k.z1, k.z2});
    }
    blocks[(size_t)i] = k;
    if (host_masks) host_masks[i] = k.mask;
  }
  if (!dev_volume || n_blocks == 0 || depth == 0) return 0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail(BRIEF_ERR_CUDA, "CUDA device %d not available", device);
  CU(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  // thresholds exactly as the reference forms them (deblock.cpp:13-21): double arithmetic, float result
  const float xa = (float)index_a;
  const float alpha = (float)(0.8 * (std::pow(2.0, (double)(xa / 6)) - 1));
  const float beta = (float)(0.5 * (double)(float)index_b - 7);
  // the reference's seam list in its order (block by block; left, right, down, up), minus the seams its border guard
  // skips (deblock.cpp:283-286), then the wave of every seam: 1 + the highest wave among the EARLIER seams it conflicts
  // with (overlapping z range, and one seam's written voxels inside the other's 6-tap footprint)
  std::vector<DeblockSeam> seams;
  for (const DeblockBlock& k : blocks)
    for (int s = 0; s < 4; ++s) {
      if (!((k.mask >> s) & 1)) continue;
      const int l = s == 1 ? k.x2 : k.x1, r = s == 0 ? k.x1 : k.x2;
      const int dn = s == 3 ? k.y2 : k.y1, up = s == 2 ? k.y1 : k.y2;
      if (l == r && (l - 3 < 0 || l + 3 > width - 1)) continue;
      else if (dn == up && (dn - 3 < 0 || dn + 3 > height - 1)) continue;
      if (l != r && dn != up) continue;  // not a line (cannot happen for the four seams of a block)
      seams.push_back(DeblockSeam{k.z1, k.z2, l, r, dn, up});
    }
  struct Rect { int x0, x1, y0, y1; };
  auto rects = [](const DeblockSeam& q, Rect& rd, Rect& wr) {
    if (q.l == q.r) { rd = {q.l - 3, q.l + 2, q.d, q.u}; wr = {q.l - 2, q.l + 1, q.d, q.u}; }
    else { rd = {q.l, q.r, q.d - 3, q.d + 2}; wr = {q.l, q.r, q.d - 2, q.d + 1}; }
  };
  auto hit = [](const Rect& a, const Rect& b) { return a.x0 <= b.x1 && b.x0 <= a.x1 && a.y0 <= b.y1 && b.y0 <= a.y1; };
  const size_t ns = seams.size();
  std::vector<int> wave(ns, 0);
  std::vector<Rect> rd(ns), wr(ns);
  for (size_t i = 0; i < ns; ++i) rects(seams[i], rd[i], wr[i]);
  int n_waves = 0;
  {
    // candidates through a coarse xy grid (64 x 64 voxel cells): a seam is compared only with the earlier seams that
    // registered in a cell its own footprint touches
    constexpr int kCell = 64;
    std::unordered_map<long long, std::vector<int>> cells;
    std::vector<int> stamp(ns, -1);
    for (size_t j = 0; j < ns; ++j) {
      int w = 0;
      const int cx0 = std::max(rd[j].x0, 0) / kCell, cx1 = std::max(rd[j].x1, 0) / kCell;
      const int cy0 = std::max(rd[j].y0, 0) / kCell, cy1 = std::max(rd[j].y1, 0) / kCell;
      for (int cy = cy0; cy <= cy1; ++cy)
        for (int cx = cx0; cx <= cx1; ++cx) {
          auto& lst = cells[((long long)cy << 32) | (unsigned int)cx];
          for (int i : lst) {
            if (stamp[(size_t)i] == (int)j) continue;  // already compared through another cell
            stamp[(size_t)i] = (int)j;
            if (wave[(size_t)i] < w) continue;
            if (seams[(size_t)i].z1 > seams[j].z2 || seams[j].z1 > seams[(size_t)i].z2) continue;
            if (hit(wr[(size_t)i], rd[j]) || hit(rd[(size_t)i], wr[j])) w = wave[(size_t)i] + 1;
          }
          lst.push_back((int)j);
        }
      wave[j] = w;
      n_waves = std::max(n_waves, w + 1);
    }
  }
  if (ns == 0) return 0;
  std::vector<int> wave_off((size_t)n_waves + 1, 0);
  for (size_t i = 0; i < ns; ++i) wave_off[(size_t)wave[i] + 1]++;
  for (int w = 0; w < n_waves; ++w) wave_off[(size_t)w + 1] += wave_off[(size_t)w];
  std::vector<DeblockSeam> sorted(ns);
  {
    std::vector<int> cur(wave_off.begin(), wave_off.end() - 1);
    for (size_t i = 0; i < ns; ++i) sorted[(size_t)cur[(size_t)wave[i]]++] = seams[i];
  }
  unsigned char* d = nullptr;
  const size_t seam_bytes = ns * sizeof(DeblockSeam), off_bytes = wave_off.size() * sizeof(int);
  CU(cudaMalloc(&d, seam_bytes + off_bytes));
  cudaError_t e = cudaMemcpyAsync(d, sorted.data(), seam_bytes, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + seam_bytes, wave_off.data(), off_bytes, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess)
    e = launch_deblock(reinterpret_cast<unsigned short*>(dev_volume), depth, height, width, reinterpret_cast<const DeblockSeam*>(d),
                       reinterpret_cast<const int*>(d + seam_bytes), n_waves, alpha, beta, thres, st);
  if (e == cudaSuccess) { g_launches.fetch_add(1); e = cudaStreamSynchronize(st); }
  cudaFree(d);
  if (e != cudaSuccess) return fail(BRIEF_ERR_CUDA, "brief_deblock: %s", cudaGetErrorString(e));
  return 0;
}

int brief_volume_quality(const void* dev_a, const void* dev_b, int32_t dtype, int32_t depth, int32_t height, int32_t width,
                         double data_range, double* host_out, int32_t device, void* stream) {
  if (!dev_a || !dev_b || !host_out || depth < 1) return fail(BRIEF_ERR_INVALID, "brief_volume_quality: bad arguments");
  if (dtype < 0 || dtype > 2) return fail(BRIEF_ERR_INVALID, "unknown dtype %d", dtype);
  if (height < 11 || width < 11)
    return fail(BRIEF_ERR_UNSUPPORTED, "brief_volume_quality: slices of %dx%d are smaller than the 11-tap window", height, width);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail(BRIEF_ERR_CUDA, "CUDA device %d not available", device);
  CU(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  // _fspecial_gauss_1d(11, 1.5), utils/ssim.py:9-24: fp32 exp, normalised by the fp32 sum
  float win[11], sum = 0.f;
  for (int k = 0; k < 11; ++k) {
    const float c = (float)(k - 5);
    win[k] = expf(-(c * c) / (2.f * 1.5f * 1.5f));
    sum += win[k];
  }
  for (int k = 0; k < 11; ++k) win[k] /= sum;
  const float c1 = (float)((0.01 * data_range) * (0.01 * data_range)), c2 = (float)((0.03 * data_range) * (0.03 * data_range));
  double* d = nullptr;
  CU(cudaMalloc(&d, 2 * sizeof(double)));
  cudaError_t e = cudaMemsetAsync(d, 0, 2 * sizeof(double), st);
  if (e == cudaSuccess) e = launch_quality(dev_a, dev_b, dtype, depth, height, width, win, c1, c2, d, st);
  if (e == cudaSuccess) { g_launches.fetch_add(1); e = cudaMemcpyAsync(host_out, d, 2 * sizeof(double), cudaMemcpyDeviceToHost, st); }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d);
  if (e != cudaSuccess) return fail(BRIEF_ERR_CUDA, "brief_volume_quality: %s", cudaGetErrorString(e));
  host_out[2] = (double)depth * (double)(height - 10) * (double)(width - 10);
  return 0;
}

int64_t brief_preprocess_scratch_bytes(int32_t depth, int32_t height, int32_t width) {
  if (depth < 1 || height < 1 || width < 1) return 0;
  return (int64_t)preprocess_scratch_bytes(depth, height, width);
}

int brief_preprocess(void* dev_volume, int32_t dtype, int32_t depth, int32_t height, int32_t width, double level,
                     const int32_t* host_close, double clip_lo, double clip_hi, void* dev_scratch, int32_t device,
                     void* stream) {
  if (!dev_volume || !dev_scratch || depth < 1 || height < 1 || width < 1)
    return fail(BRIEF_ERR_INVALID, "brief_preprocess: bad arguments");
  if (dtype != BRIEF_U8 && dtype != BRIEF_U16)
    return fail(BRIEF_ERR_UNSUPPORTED, "brief_preprocess: dtype %d (uint8 / uint16 volumes only)", dtype);
  const double tmax = dtype == BRIEF_U8 ? 255.0 : 65535.0;
  if (!(clip_lo >= 0 && clip_lo <= clip_hi && clip_hi <= tmax))  // range_limit, utils/tool.py:26-30
    return fail(BRIEF_ERR_INVALID, "Improper range setting! clip [%g, %g] outside [0, %g]", clip_lo, clip_hi, tmax);
  int sz = 1, sy = 1, sx = 1;
  if (host_close) {
    sz = host_close[0]; sy = host_close[1]; sx = host_close[2];
    if (sz < 1 || sy < 1 || sx < 1 || sz > 4 || sy > 4 || sx > 4)
      return fail(BRIEF_ERR_UNSUPPORTED, "brief_preprocess: structure %dx%dx%d (each side 1..4)", sz, sy, sx);
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail(BRIEF_ERR_CUDA, "CUDA device %d not available", device);
  CU(cudaSetDevice(device));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  // v <= level on integers: v <= floor(level); nothing qualifies below 0
  const bool any_mask = level >= 0;
  const unsigned int thr = (unsigned int)std::min(std::floor(std::max(level, 0.0)), tmax);
  // numpy clips integer data against the bounds as given: v < lo -> lo means v < ceil(lo); the stored value is the
  // bound cast to the dtype.  Configs give integers; fractional bounds are rejected rather than guessed at.
  if (clip_lo != std::floor(clip_lo) || clip_hi != std::floor(clip_hi))
    return fail(BRIEF_ERR_UNSUPPORTED, "brief_preprocess: fractional clip bounds");
  const unsigned int lo = (unsigned int)clip_lo, hi = (unsigned int)clip_hi;
  const bool clip = lo > 0 || hi < (unsigned int)tmax;
  int launches = 0;
  cudaError_t e = launch_preprocess(dev_volume, dtype, depth, height, width, thr, any_mask, sz, sy, sx, lo, hi, clip,
                                    dev_scratch, sms, &launches, (cudaStream_t)stream);
  g_launches.fetch_add(launches);
  if (e != cudaSuccess) return fail(BRIEF_ERR_CUDA, "brief_preprocess: %s", cudaGetErrorString(e));
  return 0;
}

int brief_block_stats(void* const* host_dev_raw, const int64_t* host_sizes, int32_t n_blocks, int32_t dtype,
                      int32_t device, double* host_out, void* stream) {
  if (!host_dev_raw || !host_sizes || !host_out || n_blocks < 0) return fail(BRIEF_ERR_INVALID, "brief_block_stats: bad arguments");
  if (dtype < 0 || dtype > 2) return fail(BRIEF_ERR_INVALID, "unknown dtype %d", dtype);
  if (n_blocks == 0) return 0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail(BRIEF_ERR_CUDA, "CUDA device %d not available", device);
  CU(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  long long max_size = 0;
  for (int i = 0; i < n_blocks; ++i) {
    if (!host_dev_raw[i] || host_sizes[i] < 0) return fail(BRIEF_ERR_INVALID, "block %d: null pointer or negative size", i);
    max_size = std::max<long long>(max_size, host_sizes[i]);
  }
  // one device scratch: [pointers | sizes | ord(min), ord(max) | sum, sumsq]
  const size_t nb = (size_t)n_blocks;
  const size_t off_sizes = nb * sizeof(void*), off_ord = off_sizes + nb * sizeof(long long);
  const size_t off_sum = (off_ord + nb * 2 * sizeof(unsigned int) + 7) & ~(size_t)7, total = off_sum + nb * 2 * sizeof(double);
  std::vector<unsigned char> h(total, 0);
  memcpy(h.data(), host_dev_raw, nb * sizeof(void*));
  for (size_t i = 0; i < nb; ++i) {
    reinterpret_cast<long long*>(h.data() + off_sizes)[i] = host_sizes[i];
    reinterpret_cast<unsigned int*>(h.data() + off_ord)[2 * i] = 0xffffffffu;  // ord(min) starts at the top
  }
  unsigned char* d = nullptr;
  CU(cudaMalloc(&d, total));
  cudaError_t e = cudaMemcpyAsync(d, h.data(), total, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess)
    e = launch_block_stats(reinterpret_cast<void* const*>(d), reinterpret_cast<const long long*>(d + off_sizes), n_blocks,
                           max_size, dtype, reinterpret_cast<unsigned int*>(d + off_ord), reinterpret_cast<double*>(d + off_sum),
                           sms, st);
  if (e == cudaSuccess) { g_launches.fetch_add(1); e = cudaMemcpyAsync(h.data(), d, total, cudaMemcpyDeviceToHost, st); }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d);
  if (e != cudaSuccess) return fail(BRIEF_ERR_CUDA, "brief_block_stats: %s", cudaGetErrorString(e));
  for (size_t i = 0; i < nb; ++i) {
    const unsigned int* o = reinterpret_cast<const unsigned int*>(h.data() + off_ord) + 2 * i;
    const double* s2 = reinterpret_cast<const double*>(h.data() + off_sum) + 2 * i;
    host_out[4 * i + 0] = host_sizes[i] ? (double)stats_ord_to_float(o[0]) : 0.0;
    host_out[4 * i + 1] = host_sizes[i] ? (double)stats_ord_to_float(o[1]) : 0.0;
    host_out[4 * i + 2] = s2[0];
    host_out[4 * i + 3] = s2[1];
  }
  return 0;
}

int brief_sample_indices(uint64_t seed, uint64_t step, int32_t net, int64_t batch, int64_t pop, int64_t* dev_out,
                         void* stream) {
  if (!dev_out || batch < 0 || pop < 1 || pop > 0xffffffffLL) return fail(BRIEF_ERR_INVALID, "brief_sample_indices: bad arguments");
  if (batch == 0) return 0;
  LAUNCH(launch_sample_indices(seed, step, (uint32_t)net, batch, pop, reinterpret_cast<long long*>(dev_out),
                               (cudaStream_t)stream));
  return 0;
}

}  // extern "C"
