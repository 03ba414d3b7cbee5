// brief_kernels.h — kernel argument blocks and host launchers shared between the translation
// units of libbrief_b200 (brief_simt.cu, brief_tc.cu, brief_opt.cu, brief_capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "brief_common.cuh"

namespace brief {

// Forward / decompress work: block b serves tile (b - work_prefix[i]) of network work_net[i].
struct EvalArgs {
  const NetDev* nets;
  const float* params;
  const float* axes;
  const int* work_prefix;  // n_work + 1 entries (tiles, exclusive prefix sum)
  const int* work_net;     // n_work entries, or NULL for identity
  int n_work;
  int single_net;          // >= 0: one network, block b = tile b (work tables unused)
  int TM;
  // explicit-coordinate mode (brief_forward); NULL = dense grid of the network's block
  const float* coords;
  long long n_coords;
  float* out_f32;      // explicit mode destination
  float* layers_out;   // optional z_l dump [L-1][n][f]
  // dense mode destinations
  void* const* out_ptrs;  // device array indexed by network id
  int out_dtype;
  // tensor-core path only
  const unsigned char* wpack;
  int tiles_per_block;  // 128-sample tiles served by one CTA (split round-robin over its groups)
  int eval_slots;       // tile slots per group (2, or 1 for the widest networks)
};

// Per-step scalars that live in DEVICE memory when a step is replayed as a CUDA graph (brief_fit_step_host): the graph's
// kernel arguments are frozen at capture time, so what changes from step to step is delivered by a memcpy node.
struct StepState {
  unsigned long long step;  // sampler stream position (Philox counter word)
  float neg_clr;            // OptArgs::neg_clr of this step
  float bc2_sqrt;           // OptArgs::bc2_sqrt of this step
};

// Fit work: block b serves slice (b - work_prefix[i]) of network work_net[i].
struct FitArgs {
  const NetDev* nets;
  const float* params;
  const float* axes;
  const int* work_prefix;
  const int* work_net;
  int n_work;
  int TM;
  const long long* idx;  // replayed sampler indices or NULL
  uint64_t seed, step;
  float* partials;       // per-slice gradient slots (padded device layout)
  float* loss_partials;  // per-slice loss terms
  const unsigned char* wpack;
  unsigned char* stash;     // wide tensor-core kernel: per-SM activation stash + dW scratch (stash_stride bytes per SM id)
  size_t stash_stride;
  int stash_slots;          // SM ids covered by `stash` (%nsmid of the device)
  const StepState* state;   // non-NULL: `step` is read from device memory (graph replay)
};

// Layer-wise tensor-core path for wide networks (brief_tc_lw.cu): one pass = a set of networks of one padded width whose
// tiles share a scratch arena.  Work entry i = network work_net[i], tiles [tile_first[i], tile_first[i] + tile_count[i])
// of it, stored at scratch tiles tile_base[i]...; `work_prefix` is the launch's block -> entry table (tiles for the
// per-tile kernels, CTAs for the GEMM kernel).
struct LwArgs {
  const NetDev* nets;
  const int* work_prefix;
  const int* work_net;
  const int* tile_count;
  const long long* tile_base;
  const int* tile_first;
  int n_work;
  int F;              // padded width of the bucket
  long long T;        // tiles of the pass (scratch tensor stride)
  int layer;          // hidden-layer index of this launch (weights at image offset layer * F * F * 2)
  int kind, nb, max_slices, eval;
  size_t in_off, in2_off, out_off, out2_off;  // scratch byte offsets of the launch's tile tensors
  unsigned char* scratch;
  const unsigned char* wpack;
  const float* axes;
  const long long* idx;
  uint64_t seed, step;
  const StepState* state;
  const float* coords;    // explicit-coordinate forward
  long long n_coords;
  float* out_f32;
  float* layers_out;
  void* const* out_ptrs;
  int out_dtype;
  float* partials;
  float* loss_partials;
};

// Optimiser step over the whole group (one launch).
struct OptArgs {
  const NetDev* nets;
  const int* blk_prefix;  // n_nets + 1: blocks per network (256 params each)
  int n_nets;
  float* params;
  float* grads;
  float* m;
  float* v;
  const float* partials;       // NULL: consume `grads`; else reduce slices in order first
  const float* loss_partials;  // with partials: per-network loss = ordered sum
  float* loss_out;             // n_nets floats or NULL
  int kind;
  float neg_clr;     // -(lr / (1 - beta1^t))   [Adamax/Adam] or -lr [SGD]
  float w1;          // 1 - beta1   (lerp weight)
  float beta2;
  float w2;          // 1 - beta2   (Adam addcmul value)
  float eps;
  float bc2_sqrt;    // sqrt(1 - beta2^t) (Adam)
  int apply;         // 0 = only reduce partials into grads / loss (brief_fit_step)
  unsigned char* wpack;  // refreshed fp16 operand image (tensor-core networks) or NULL
  const StepState* state;  // non-NULL: neg_clr / bc2_sqrt are read from device memory (graph replay)
};

// brief_simt.cu
int simt_pick_tm(int F4, int L, bool fit, size_t smem_limit);
size_t simt_eval_smem(int F4, int TM);
size_t simt_fit_smem(int F4, int L, int TM);
cudaError_t launch_simt_eval(const EvalArgs& a, int n_blocks, size_t smem, cudaStream_t st);
cudaError_t launch_simt_fit(const FitArgs& a, int n_blocks, size_t smem, cudaStream_t st);
cudaError_t launch_gather(const NetDev* nets, int net_id, const float* axes, const long long* idx, long long batch,
                          float* coords, float* data, float* weight, cudaStream_t st);
cudaError_t launch_sample_indices(uint64_t seed, uint64_t step, uint32_t net, long long batch, long long pop,
                                  long long* out, cudaStream_t st);
cudaError_t launch_gen_indices(uint64_t seed, uint64_t step, const StepState* state, uint32_t stream, long long batch,
                               long long pop, int h, int w, int ch, int cw, long long cube_vox, const long long* cube_ids,
                               long long* out, cudaStream_t st);

// brief_data.cu
cudaError_t launch_block_stats(const void* const* dev_ptrs, const long long* dev_sizes, int n_blocks, long long max_size,
                               int dtype, unsigned int* stat_ord, double* stat_sum, int num_sms, cudaStream_t st);
float stats_ord_to_float(unsigned int o);
cudaError_t query_nsmid(int* out, cudaStream_t st);  // PTX %nsmid: upper bound (exclusive) of %smid on this device
cudaError_t launch_histogram(const void* dev_raw, long long n, int dtype, unsigned long long* dev_hist, int num_sms,
                             cudaStream_t st);

// brief_deblock.cu
struct DeblockBlock { int z1, z2, y1, y2, x1, x2, mask; };  // inclusive ends; mask bit 0..3 = left, right, down, up seam listed
struct DeblockSeam { int z1, z2, l, r, d, u; };                // one listed seam: columns l..r, rows d..u of slices z1..z2
cudaError_t launch_deblock(unsigned short* img, int D, int H, int W, const DeblockSeam* dev_seams, const int* dev_wave_off,
                           int n_waves, float alpha, float beta, int thres, cudaStream_t st);

// brief_quality.cu
cudaError_t launch_quality(const void* a, const void* b, int dtype, int D, int H, int W, const float* win11, float c1, float c2,
                           double* dev_out, cudaStream_t st);

// brief_preprocess.cu
size_t preprocess_scratch_bytes(int D, int H, int W);
cudaError_t launch_preprocess(void* vol, int dtype, int D, int H, int W, unsigned int thr, bool any_mask, int sz, int sy,
                              int sx, unsigned int lo, unsigned int hi, bool clip, void* scratch, int num_sms, int* launches,
                              cudaStream_t st);

// brief_opt.cu
cudaError_t launch_opt(const OptArgs& a, int n_blocks, cudaStream_t st);
cudaError_t launch_pack(const NetDev* nets, int n_nets, const float* params, unsigned char* wpack, cudaStream_t st);

// brief_tc.cu (tcgen05 path)
bool tc_supported(int f, int L, int in_dim, int out_dim);
int tc_fpad(int f);
int tc_fit_ctas_per_sm(int F_PAD, int L);
size_t tc_fit_stash_bytes(int F_PAD, int L);  // wide kernel (F_PAD > 64): activation stash per CTA, else 0
size_t tc_wpack_bytes(int F_PAD, int L);
size_t tc_eval_smem(int F_PAD, int L);
int tc_eval_groups(int F_PAD, int L);
int tc_eval_slots(int F_PAD, int L);
bool tc_eval_supported(int f, int L, int in_dim, int out_dim);
size_t tc_fit_smem(int F_PAD, int L);
cudaError_t launch_tc_eval(const EvalArgs& a, int F_PAD, int L_max, int n_blocks, cudaStream_t st);
cudaError_t launch_tc_fit(const FitArgs& a, int F_PAD, int L_max, int n_blocks, cudaStream_t st);

// brief_tc_lw.cu (layer-wise tcgen05 path, 128 < F_PAD <= 256)
bool tc_lw_supported(int f, int L, int in_dim, int out_dim);
size_t lw_scratch_bytes(int F, int L, long long T);
size_t lw_eval_scratch_bytes(int F, long long T);
size_t lw_tile_bytes_host(int F);
size_t lw_fit_offset(int F, int L, long long T, int what, int j);  // what: 0 ACT_j, 1 COS_j, 2 DZ_j, 3 X block, 4 DY block
cudaError_t launch_lw_sample_l0(const LwArgs& a, int n_tiles, cudaStream_t st);
cudaError_t launch_lw_gemm(const LwArgs& a, int mode, int n_ctas, cudaStream_t st);
cudaError_t launch_lw_last(const LwArgs& a, int n_tiles, cudaStream_t st);
cudaError_t launch_lw_dw(const LwArgs& a, int n_nets, cudaStream_t st);
cudaError_t launch_lw_loss_reduce(const LwArgs& a, int n_nets, cudaStream_t st);

}  // namespace brief
