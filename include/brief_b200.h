/*
 * brief_b200.h — C ABI of the B200-native BRIEF hot path (libbrief_b200.so).
 *
 * The reference (RichealYoung/BRIEF_PyTorch) has no FFI: its boundary for this path is Python
 * (SURVEY.md section 8b).  This header is the C-ABI that sits BENEATH a Python mirror of that
 * boundary (package brief_pytorch_b200).  Every entry point names the reference interface whose
 * work it replaces (file:line relative to the reference root).
 *
 * Conventions
 *  - All functions return 0 (BRIEF_OK) on success and a negative BriefStatus otherwise;
 *    brief_last_error() returns a thread-local message for the last failure.
 *  - "dev_" pointers are device pointers owned by the caller (e.g. torch allocations),
 *    "host_" pointers are ordinary host memory.  `stream` is a cudaStream_t passed as void*.
 *    All work is enqueued on that stream; calls taking host pointers synchronise it.
 *  - A BriefGroup is a set of independent per-block SIREN networks (one per volume block,
 *    main.py:484-532) that are evaluated / fitted together in grouped kernel launches.
 *  - Parameters cross the ABI in the reference's `parameters()` order, fp32, exactly as
 *    utils/ModelSave.py:32-51 writes them: W_0[f][in], b_0[f], W_1[f][f], b_1[f], ...,
 *    W_{L-1}[out][f], b_{L-1}[out]  (row-major [out][in]).
 *  - There is no CPU fallback: without a CUDA device every compute entry point fails with
 *    BRIEF_ERR_CUDA.
 */
#ifndef BRIEF_B200_H_
#define BRIEF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BRIEF_B200_ABI_VERSION 1

typedef enum BriefStatus {
  BRIEF_OK = 0,
  BRIEF_ERR_INVALID = -1,     /* bad argument / unsupported configuration            */
  BRIEF_ERR_CUDA = -2,        /* CUDA runtime error (message has the CUDA string)    */
  BRIEF_ERR_UNSUPPORTED = -3, /* valid request outside what this build implements    */
  BRIEF_ERR_STATE = -4        /* call order violated (e.g. fit before bind_volume)   */
} BriefStatus;

/* Arithmetic of the hidden-layer contractions. First/last layer, loss, optimiser: always fp32. */
typedef enum BriefPrecision {
  BRIEF_PREC_FP32 = 0, /* CUDA-core fp32 everywhere (exact mode; any width up to the smem limit) */
  BRIEF_PREC_F16 = 1,  /* tcgen05.mma kind::f16: fp16 operands (activations live in [-1,1], weights << 1, so
                          fp16 carries 3 more mantissa bits than bf16 at the same tensor rate; gradients
                          are carried with a static power-of-two scale), fp32 accumulate in TMEM     */
  BRIEF_PREC_AUTO = 2  /* F16 where the fused tensor-core kernel supports the shape, else FP32   */
} BriefPrecision;

typedef enum BriefDType { BRIEF_U8 = 0, BRIEF_U16 = 1, BRIEF_F32 = 2 } BriefDType;

typedef enum BriefSamplerMode {
  BRIEF_SAMPLE_FULL_BLOCK = 0,   /* RandomCubeSampler with pop_size==1: every step = whole block in
                                    voxel order (main.py:38-125 with the shipped cube_len)         */
  BRIEF_SAMPLE_RANDOM_POINTS = 1 /* RandompointSampler (main.py:126-163): `batch` indices with
                                    replacement per step                                           */
} BriefSamplerMode;

typedef enum BriefOptimizer { BRIEF_OPT_ADAMAX = 0, BRIEF_OPT_ADAM = 1, BRIEF_OPT_SGD = 2 } BriefOptimizer;

/* One network = SIREN(coords_channel, data_channel, features, layers, w0) of utils/Networks.py:246-266
 * bound to one block of `dims` voxels.  Hidden-layer omega is the reference's hard-wired 30
 * (utils/Networks.py:227-229,259) unless overridden here. */
typedef struct BriefNetDesc {
  int32_t coords_channel; /* 2 or 3 */
  int32_t data_channel;   /* 1 (the fused kernels support a single output channel)        */
  int32_t features;       /* f                                                               */
  int32_t layers;         /* L >= 2: 1 input sine layer, L-2 hidden sine layers, 1 linear    */
  float w0;               /* first-layer omega                                               */
  float w_hidden;         /* hidden-layer omega, 30 in the reference                         */
  int32_t dims[3];        /* block extent (d,h,w); for 2-D data d = 1 and coords are (h,w)   */
} BriefNetDesc;

/* parse_weight 'value_l_h_s' rules (utils/misc.py:293-297), applied in order on the raw voxel. */
typedef struct BriefWeightRule {
  float lo, hi, scale;
} BriefWeightRule;
#define BRIEF_MAX_WEIGHT_RULES 4

typedef struct BriefOptConfig {
  int32_t kind;    /* BriefOptimizer; configure_optimizer utils/misc.py:174-183                */
  float lr;        /* initial lr (lr_phi)                                                      */
  float beta1, beta2, eps; /* torch defaults 0.9, 0.999, 1e-8                                  */
  int32_t n_milestones;    /* MultiStepLR (utils/misc.py:187-188); 0 = constant lr             */
  int64_t milestones[8];
  float gamma;
} BriefOptConfig;

typedef struct BriefGroup BriefGroup;

/* ---- library ------------------------------------------------------------------------------- */
int brief_abi_version(void);
const char* brief_last_error(void);
int brief_device_count(int* out_count);

/* torch.linspace(lo, hi, n) on the CPU, bit-for-bit (utils/dataset.py:22-32,47-58 call sites).
 * Used to build the per-axis coordinate tables when the caller does not supply them. */
int brief_linspace(float lo, float hi, int32_t n, float* host_out);

/* ---- group life cycle ------------------------------------------------------------------------ */
/* Replaces: init_phi / SIREN.__init__ (utils/Networks.py:246-266, 800-802) for n_nets networks and
 * the per-block process farm of main.py:547-579.  Parameters start at zero: load them with
 * brief_group_set_params (initialisation stays in Python so that it is bit-identical to torch's
 * CPU generator under the same seed). */
int brief_group_create(const BriefNetDesc* nets, int32_t n_nets, int32_t device, int32_t precision,
                       BriefGroup** out);
void brief_group_destroy(BriefGroup* g);
int brief_group_num_nets(const BriefGroup* g);
int brief_group_param_count(const BriefGroup* g, int32_t net);      /* SIREN.calc_param_count :291-297 */
int brief_group_precision(const BriefGroup* g, int32_t net);        /* resolved BriefPrecision          */
int brief_group_batch(const BriefGroup* g, int32_t net);            /* samples per step under the current sampler */

/* Replaces load_model / save_model's tensor traffic (utils/ModelSave.py:8-51). Packed fp32, host. */
int brief_group_set_params(BriefGroup* g, int32_t net, const float* host_packed, void* stream);
int brief_group_get_params(BriefGroup* g, int32_t net, float* host_packed, void* stream);
/* Gradients of the last brief_fit_step (same packed order); for parity tests. */
int brief_group_get_grads(BriefGroup* g, int32_t net, float* host_packed, void* stream);
/* Test hook: overwrite the gradient arena of one network (packed order), so that brief_opt_step can
 * be checked against torch.optim on identical gradients. */
int brief_group_set_grads(BriefGroup* g, int32_t net, const float* host_packed, void* stream);
/* Optimiser state (exp_avg, exp_inf / exp_avg_sq) of torch.optim, packed order; tests + resume. */
int brief_group_get_opt_state(BriefGroup* g, int32_t net, float* host_m, float* host_v, void* stream);
int brief_group_reset_opt_state(BriefGroup* g, void* stream);

/* Per-axis coordinate tables (create_coords, utils/dataset.py:11-35): axis k of network `net`
 * has dims[k] fp32 entries.  If never called the tables are brief_linspace(-1, 1, dims[k]). */
int brief_group_set_axes(BriefGroup* g, int32_t net, const float* host_d, const float* host_h,
                         const float* host_w, void* stream);

/* ---- data binding ----------------------------------------------------------------------------- */
/* Replaces normalize_data (utils/io.py:67-80), parse_weight (utils/misc.py:272-307) and the
 * samplers' materialised fp32 copies (main.py:134-139): the raw block stays in HBM in its own
 * dtype and is normalised on chip as ((x - vmin) / (vmax - vmin)) * (hi - lo) + lo, fp32, the
 * reference's operation order.  `dev_weight` (fp32, one per voxel) may be NULL, in which case
 * `rules` are evaluated on the raw voxel (weight 1 when no rule matches).  `tau` is the normalised
 * weight threshold of main.py:380-383 (0 disables the override, like the reference's truthiness test). */
int brief_group_bind_volume(BriefGroup* g, int32_t net, const void* dev_raw, int32_t dtype, float vmin,
                            float vmax, float lo, float hi, const float* dev_weight,
                            const BriefWeightRule* rules, int32_t n_rules, float tau);
/* Sampler of main.py:367-371 for this network. `batch` is ignored for FULL_BLOCK (= voxel count). */
int brief_group_set_sampler(BriefGroup* g, int32_t net, int32_t mode, int32_t batch);

/* The GENERAL RandomCubeSampler (main.py:38-125, == utils/sampler.py:9-57): every step draws `cube_count` windows of
 * cube_len voxels (clamped to the block, main.py:49-50; for 2-D data pass {1, len_h, len_w}) with replacement from all
 * stride-1 window positions, listed '(dc hc wc)' (main.py:61-69), and the loss is the mean over all their voxels.
 * A cube network then counts as a RANDOM_POINTS network of cube_count * prod(cube_len) samples per step: replayed
 * indices (brief_fit_step's dev_idx, brief_fit_step_host's host_idx) are VOXEL indices in cube order — brief_cube_indices
 * expands the reference's torch.randint cube draws (main.py:114) into them — and with the on-device sampler the cube
 * draws come from the network's Philox stream (draw c of the step = cube c).  cube_len >= block with cube_count 1 (every
 * shipped config) degenerates to BRIEF_SAMPLE_FULL_BLOCK.  brief_group_set_sampler on the network removes the cubes. */
int brief_group_set_cube_sampler(BriefGroup* g, int32_t net, int32_t cube_count, const int32_t* host_cube_len);
/* Voxel indices (int64, cube_count * prod(cube_len), cube after cube in 'ds hs ws' order) of one step of a cube
 * network: dev_cube_ids = cube_count window indices (int64, device memory), or NULL for the on-device stream at
 * (seed, step).  Asynchronous on `stream`. */
int brief_cube_indices(BriefGroup* g, int32_t net, const int64_t* dev_cube_ids, uint64_t seed, uint64_t step,
                       int64_t* dev_out, void* stream);

/* Key of the network's on-device sampler stream (Philox counter word).  Default: the network's index in the group;
 * a caller that shards blocks over ranks passes the block's GLOBAL index so that a block draws the same samples
 * whichever rank owns it. */
int brief_group_set_stream(BriefGroup* g, int32_t net, uint32_t stream_id);

/* How a network's step batch is cut into per-CTA slices (each slice accumulates its gradient in TMEM, the optimiser
 * kernel adds the slices in order).  FILL_WAVE (default, fastest): slices sized so that ALL networks of a width
 * bucket together fill one wave of CTAs.  PER_NETWORK: a network's slices depend only on its own batch and the
 * device, so its fp32 summation order — and therefore every bit of its result — is independent of what else shares
 * the GPU (the 1/2/4/8-GPU equality of SURVEY.md 8e); costs extra waves when many networks share a GPU. */
typedef enum BriefSlicing { BRIEF_SLICING_FILL_WAVE = 0, BRIEF_SLICING_PER_NETWORK = 1 } BriefSlicing;
int brief_group_set_slicing(BriefGroup* g, int32_t mode);

/* ---- hot path ----------------------------------------------------------------------------------- */
/* One training step for EVERY network of the group: sampler gather + forward + weighted L2
 * (datal2, main.py:176-182) + backward (main.py:385-396).  Gradients stay on the device.
 * dev_idx: NULL = on-device sampler (Philox4x32-10 keyed by seed, step), or the replayed indices of
 * the reference's torch.randint stream, int64, all networks concatenated in order (RANDOM_POINTS
 * networks contribute `batch` entries each, FULL_BLOCK networks none).
 * dev_loss: n_nets floats (mean weighted loss per network), may be NULL. */
int brief_fit_step(BriefGroup* g, const int64_t* dev_idx, uint64_t seed, uint64_t step, float* dev_loss,
                   void* stream);
/* Bench / profiling hook: ONLY the fit kernel(s) of brief_fit_step (per-slice gradient partials are produced but not
 * reduced), so that bench.py can time the dominant kernel alone with CUDA events on `stream`. */
int brief_fit_kernels(BriefGroup* g, const int64_t* dev_idx, uint64_t seed, uint64_t step, void* stream);
/* optimizer.step() + lr_scheduler.step() (main.py:399-400, utils/misc.py:174-197) for every network,
 * consuming the gradients of the preceding brief_fit_step.  `lr` is this step's learning rate and
 * `t` the 1-based optimiser step count (bias correction). */
int brief_opt_step(BriefGroup* g, int32_t kind, float lr, float beta1, float beta2, float eps, int64_t t,
                   void* stream);
/* The loop of main.py:385-400 for `n_steps` steps starting after `steps_done` completed steps, entirely
 * enqueued from C (no per-step host sync, no loss.item()).  dev_loss_hist: NULL or n_steps*n_nets floats. */
int brief_fit_run(BriefGroup* g, const BriefOptConfig* cfg, uint64_t seed, int64_t steps_done, int64_t n_steps,
                  float* dev_loss_hist, void* stream);

/* The loop body of main.py:385-401 for ONE step with HOST buffers — what a caller that draws the sampler indices with
 * the CPU generator (main.py:156) and reads loss.item() every step (main.py:401) needs — enqueued as: [host->device: step scalars,
 * sampler indices] on an internal copy stream (so that they cross PCIe under the kernels of the step before), then ONE
 * CUDA-graph launch on `stream`: fit kernel(s) -> optimiser kernel -> [device->host: per-network loss].  host_idx: PINNED host memory, int64, all RANDOM_POINTS networks' indices
 * concatenated in network order (NULL: on-device sampler stream keyed by seed and step); host_loss: PINNED host
 * memory, n_nets floats, or NULL.  `steps_done` completed steps precede this one (optimiser step count, MultiStepLR
 * position, sampler stream position).  The call returns as soon as the step is enqueued; the caller synchronises on
 * `stream` before it reads host_loss or rewrites host_idx.  The graph is built on the first call for a given
 * (host_idx, host_loss, stream, cfg) and replayed afterwards; up to 4 such buffer sets are cached (double buffering). */
int brief_fit_step_host(BriefGroup* g, const int64_t* host_idx, const BriefOptConfig* cfg, uint64_t seed,
                        int64_t steps_done, float* host_loss, void* stream);

/* SIREN.forward on caller coordinates (utils/Networks.py:269-271): dev_coords [n][coords_channel] fp32
 * -> dev_out [n][1] fp32.  dev_layers (optional, may be NULL): pre-activations z_l of every layer,
 * layout [layers-1][n][features] fp32, for per-layer parity checks. */
int brief_forward(BriefGroup* g, int32_t net, const float* dev_coords, int64_t n, float* dev_out,
                  float* dev_layers, void* stream);

/* Replaces reconstruct_flattened + invnormalize_data (utils/misc.py:59-92, utils/io.py:136-147) for every
 * network of the group in one launch: dense grid generated on chip from the axis tables, evaluated, then
 * ((y - lo) / (hi - lo)) clipped to [0,1], * (vmax - vmin) + vmin, truncating cast to out_dtype.
 * host_dev_out[i] is the device destination of network i (dims[0]*dims[1]*dims[2] elements, contiguous), or NULL to
 * leave network i out of this call (a caller that overlaps decode with device->host copies decodes block by block);
 * out_dtype BRIEF_F32 skips the inverse normalisation and stores the raw network output.
 * vmin/vmax/lo/hi default to the values given to brief_group_bind_volume, or override with
 * brief_group_set_denorm for decode-only groups. */
int brief_group_set_denorm(BriefGroup* g, int32_t net, float vmin, float vmax, float lo, float hi);
int brief_decompress(BriefGroup* g, void* const* host_dev_out, int32_t out_dtype, void* stream);

/* The reference samplers' three outputs materialised (main.py:156-160), for parity and for the
 * HBM-bandwidth measurement of the gather: coords [batch][coords_channel], data [batch], weight [batch]. */
int brief_gather(BriefGroup* g, int32_t net, const int64_t* dev_idx, int64_t batch, float* dev_coords,
                 float* dev_data, float* dev_weight, void* stream);
/* The on-device sampler's index stream for (seed, step, net): int64 [batch] in [0, pop). */
int brief_sample_indices(uint64_t seed, uint64_t step, int32_t net, int64_t batch, int64_t pop,
                         int64_t* dev_out, void* stream);

/* Replaces the host numpy passes of normalize_data's data.min() / data.max() (utils/io.py:67-80) and the per-chunk
 * variance of alloc_param 'by_var' (utils/misc.py:402-422) for n_blocks contiguous raw blocks in ONE launch:
 * host_dev_raw[i] = device pointer of block i (dtype as in brief_group_bind_volume), host_sizes[i] = its element
 * count.  host_out receives 4 doubles per block: min, max (exact values of the raw dtype), sum, sum of squares
 * (accumulated in fp64).  Synchronises `stream`. */
int brief_block_stats(void* const* host_dev_raw, const int64_t* host_sizes, int32_t n_blocks, int32_t dtype,
                      int32_t device, double* host_out, void* stream);

/* Value histogram of one raw uint8 / uint16 block in device memory (host_hist: 256 / 65536 counters): what the
 * 'quantile_ge_ql_qh_s' weight rule needs instead of np.quantile over a host copy of the block (utils/misc.py:298-305;
 * the two order statistics are read off the cumulative counts).  Synchronises `stream`. */
int brief_block_histogram(const void* dev_raw, int64_t n, int32_t dtype, uint64_t* host_hist, int32_t device, void* stream);

/* Replaces the reference's `preprocess` (utils/misc.py:244-254; called on every block before the fit, main.py:336,
 * and on every decoded block, main.py:295) on one block [depth][height][width] of uint8 / uint16 voxels in device
 * memory, IN PLACE and bit for bit:
 *     v[ binary_opening(v <= level, structure = ones(close[0], close[1], close[2])) ] = 0;  v = clip(v, clip_lo, clip_hi)
 * host_close == NULL is the reference's `denoise_close == False` (plain threshold, no opening); each close[k] must be
 * in 1..4.  For a 2-D image [height][width] pass depth = 1 and close = {1, c0, c1} (utils/misc.py:251).  clip_lo /
 * clip_hi must satisfy 0 <= lo <= hi <= dtype max (the reference's range_limit assertion) else BRIEF_ERR_INVALID.
 * dev_scratch: brief_preprocess_scratch_bytes(depth, height, width) bytes of device memory (two 1-bit-per-voxel masks).
 * Asynchronous on `stream`. */
int64_t brief_preprocess_scratch_bytes(int32_t depth, int32_t height, int32_t width);
int brief_preprocess(void* dev_volume, int32_t dtype, int32_t depth, int32_t height, int32_t width, double level,
                     const int32_t* host_close, double clip_lo, double clip_hi, void* dev_scratch, int32_t device,
                     void* stream);

/* Replaces the reference's deblocking post-filter (deblock.cpp:226-321, its only native component; deblock.py is the
 * float twin) on a decoded uint16 volume [depth][height][width] in device memory, in place, bit for bit:
 * host_blocks holds n_blocks x 6 int32 (z1, z2, y1, y2, x1, x2, inclusive ends — the chunk directory names
 * d_z1_z2-h_y1_y2-w_x1_x2 of compressed/module) IN THE ORDER the reference's readdir() loop lists them: seams are
 * filtered sequentially in place, so the order is part of the result.  index_a / index_b / thres: deblock.cpp:326
 * (51, 2000, 65535).  host_masks (optional, n_blocks int32): receives the seam mask per block (bit 0..3 = left, right,
 * down, up seam listed; the reference's sticky duplicate flags, deblock.cpp:244-276).  With dev_volume == NULL only the
 * masks are computed (no CUDA call). */
int brief_deblock(void* dev_volume, int32_t depth, int32_t height, int32_t width, const int32_t* host_blocks,
                  int32_t n_blocks, int32_t index_a, int32_t index_b, int32_t thres, int32_t* host_masks, int32_t device,
                  void* stream);

/* Replaces eval_performance / cal_mse / cal_psnr / cal_ssim (utils/misc.py:447-499) and utils/ssim.py:9-150 for two
 * volumes [depth][height][width] of the same dtype in device memory, in one pass: host_out[0] = sum of squared
 * differences, host_out[1] = sum of the SSIM map (11-tap Gaussian window sigma 1.5, valid region, K = (0.01, 0.03),
 * C1/C2 from data_range) over all valid pixels of all slices, host_out[2] = number of valid pixels
 * (depth * (height-10) * (width-10)).  MSE = out[0] / voxels, PSNR = -10 log10(MSE / data_range^2),
 * SSIM = out[1] / out[2] (all slices have the same valid area, so this is the reference's mean of slice means).
 * height, width >= 11.  Synchronises `stream`. */
int brief_volume_quality(const void* dev_a, const void* dev_b, int32_t dtype, int32_t depth, int32_t height, int32_t width,
                         double data_range, double* host_out, int32_t device, void* stream);

/* Kernel launches issued by this library since the last reset (bench.py's gpu_launches). */
int64_t brief_launch_count(void);
void brief_reset_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* BRIEF_B200_H_ */
