/* nerf_b200.h -- C ABI of libnerf_b200.so: the B200 (sm_100a) implementation of the
 * Phase2 NeRF render / train-step hot path of ColinBalfour/SfM-NeRF.
 *
 * The reference has no FFI: its boundary is the Python call surface of Phase2/train.py and
 * Phase2/NeRFModel.py.  Each entry point below replaces the arithmetic of the reference lines
 * it cites; the Python host (sfm-nerf_b200/train.py, NeRFModel.py) keeps the reference's
 * names/signatures and binds these symbols with ctypes.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns all memory; the library never allocates or frees user-visible buffers
 *     (sizes of opaque buffers come from the *_bytes() queries);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden syncs;
 *   - return value: 0 = success, <0 = NERF_ERR_* argument/state error, >0 = cudaError_t;
 *   - re-entrant and thread-safe: no state is shared between host threads (the only cached state -- device attributes, one
 *     helper stream + two events per device -- is per thread and released when the thread exits; the layout tables are
 *     immutable after a thread-safe one-time initialisation); one device (the current one) per call;
 *   - there is no CPU fallback: on a device that is not sm_100 every compute entry point
 *     returns NERF_ERR_ARCH.
 */
#ifndef NERF_B200_H_
#define NERF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NERF_ABI_VERSION 4

enum {
  NERF_OK = 0,
  NERF_ERR_ARG = -1,      /* null pointer, negative size, unsupported parameter value */
  NERF_ERR_ALIGN = -2,    /* pointer not aligned as documented */
  NERF_ERR_ARCH = -3,     /* current device is not compute capability 10.x */
  NERF_ERR_LAYOUT = -4,   /* unknown model layout id */
  NERF_ERR_SELFTEST = -5  /* a self test produced wrong results */
};

/* Model layouts. Hidden width is 256, rgb hidden 128 (NeRFModel.py:116-140). */
enum {
  NERF_LAYOUT_LIVE_PE = 0,   /* NeRFmodel(10, 4, 256, pos_encoding=True): 661,636 params */
  NERF_LAYOUT_LIVE_NOPE = 1, /* NeRFmodel(.., pos_encoding=False):      627,844 params */
  /* the same parameter layouts with density = relu(.) instead of softplus(.): what the reference's shipped (legacy)
   * checkpoints evaluate once remapped onto the live layer table (commented class NeRFModel.py:7-99; no-PE layout of
   * Phase2/checkpoints_no_pos_encoding) -- see sfm-nerf_b200/legacy.py */
  NERF_LAYOUT_LIVE_PE_RELU = 2,
  NERF_LAYOUT_LIVE_NOPE_RELU = 3
};

/* Layout id of NeRFmodel(embed_pos_L, embed_direction_L, num_channels, pos_encoding) (NeRFModel.py:102-140; the CLI's
 * --n_pos_freq / --n_dirc_freq / --hidden_size, train.py:796-797, 813): the four ids above for the defaults, a generalised id for
 * any other frequency counts 0..10 (3 * (2L + 1) <= 63 channels fit the 64-channel encoding tile) and for num_channels 128
 * (256 and 128 are built).  density_relu != 0: ReLU density head.  NERF_ERR_LAYOUT for anything else.  Every entry point that
 * takes `layout` accepts these ids. */
int nerf_layout_id(int embed_pos_L, int embed_direction_L, int pos_encoding, int density_relu, int num_channels);

int nerf_abi_version(void);
const char* nerf_error_string(int code);
/* 0 when the current CUDA device can run the kernels (sm_100), else NERF_ERR_ARCH / cudaError_t. */
int nerf_device_check(void);

/* ---------------------------------------------------------------------------------------
 * Parameter vector.  All 26 tensors live in ONE flat fp32 buffer in state_dict order
 * (fc1.weight, fc1.bias, fc_layers.0.weight, ... rgb_layer2.bias; NeRFModel.py:116-140),
 * each tensor row-major [out, in] exactly as torch.nn.Linear stores it.  Gradients and Adam
 * moments use the same layout.
 * ------------------------------------------------------------------------------------- */
int64_t nerf_param_count(int layout);                       /* 661636 / 627844, <0 on error */
/* offset (in floats) of tensor `index` (0..25, weight/bias interleaved) or <0 */
int64_t nerf_param_offset(int layout, int index);

/* ---------------------------------------------------------------------------------------
 * K1  ray generation -- replaces PixelToRay (train.py:68-102), the per-pixel loops of
 * test_image (train.py:735-742) and generateBatch (train.py:104-127).  fp64 arithmetic,
 * rounded once to fp32 like the reference (train.py:127, 741-742).
 *   poses   : [n_images,4,4] fp64 camera-to-world, row-major
 *   fx..cy  : camera_matrix[0][0], [1][1], [0][2], [1][2]
 * image mode: rays ray_begin .. ray_begin+ray_count-1 of image `pose_index`, ray id = y*W+x.
 * ------------------------------------------------------------------------------------- */
int nerf_raygen_image(const double* poses, int pose_index, double fx, double fy, double cx,
                      double cy, int width, int height, int64_t ray_begin, int64_t ray_count,
                      float* rays_o, float* rays_d, void* stream);
/* batch mode: ray i uses (img_idx[i], px[i], py[i]); `images` ([n,H,W,3] fp32) and `target`
 * ([N,3]) may both be NULL. */
int nerf_raygen_batch(const double* poses, int n_images, double fx, double fy, double cx,
                      double cy, int width, int height, const int32_t* img_idx,
                      const int32_t* px, const int32_t* py, const float* images, int64_t n_rays,
                      float* rays_o, float* rays_d, float* target, void* stream);
/* device-side replacement for the three random.randint draws per ray (train.py:120-121);
 * counter-based, so (seed, offset) fully determine the output. */
int nerf_draw_pixels(uint64_t seed, uint64_t offset, int n_images, int width, int height,
                     int64_t n_rays, int32_t* img_idx, int32_t* px, int32_t* py, void* stream);

/* One train step's inputs in ONE launch (train.py:104-127 generateBatch + train.py:205-217 depths): for ray i draws
 * (image, x, y) exactly like nerf_draw_pixels(seed, ray_offset), generates the ray and gathers its target colour like
 * nerf_raygen_batch, and writes z[n_rays, n_sample] exactly like nerf_sample_z(perturb, t_rand = NULL, z_seed, z_offset). */
int nerf_train_batch(const double* poses, int n_images, double fx, double fy, double cx, double cy, int width, int height,
                     const float* images, uint64_t seed, uint64_t ray_offset, int64_t n_rays, float* rays_o, float* rays_d,
                     float* target, int n_sample, float near, float far, int perturb, uint64_t z_seed, uint64_t z_offset,
                     float* z, void* stream);

/* The same launch for a batch that already exists: batch9 is a generateBatch() result [n_rays, 9] = o | d | rgb fp32
 * (train.py:127, 448 -- e.g. just uploaded from host memory).  Splits it into rays_o / rays_d / target [n_rays, 3] and writes
 * z exactly like nerf_sample_z(perturb, t_rand = NULL, z_seed, z_offset). */
int nerf_unpack_batch(const float* batch9, int64_t n_rays, float* rays_o, float* rays_d, float* target, int n_sample, float near,
                      float far, int perturb, uint64_t z_seed, uint64_t z_offset, float* z, void* stream);

/* ---------------------------------------------------------------------------------------
 * K2  stratified depths -- train.py:205-217.  z is [n_rays, n_sample] fp32.
 *   perturb == 0            : t = linspace(0,1,S)
 *   perturb != 0, t_rand    : jitter with the caller's uniforms [n_rays,n_sample]
 *   perturb != 0, no t_rand : jitter with counter-based uniforms from (seed, offset)
 * ------------------------------------------------------------------------------------- */
int nerf_sample_z(int64_t n_rays, int n_sample, float near, float far, int perturb,
                  const float* t_rand, uint64_t seed, uint64_t offset, float* z, void* stream);

/* ---------------------------------------------------------------------------------------
 * K3  positional encoding -- NeRFmodel.position_encoding (NeRFModel.py:142-156).
 *   x [n,3] -> out [n, 3*(2L+1)] fp32, channel order [x, sin0, cos0, sin1, cos1, ...].
 * ------------------------------------------------------------------------------------- */
int nerf_encode(const float* x, int64_t n, int L, float* out, void* stream);

/* ---------------------------------------------------------------------------------------
 * K4  fused MLP -- NeRFmodel.forward (NeRFModel.py:158-202) and its backward.
 * Weights are consumed in a packed bf16 image (UMMA operand order) built from the flat fp32
 * parameter vector; rebuild it after every optimizer step.
 * ------------------------------------------------------------------------------------- */
int64_t nerf_mlp_packed_bytes(int layout);
int nerf_mlp_pack_weights(const float* params, int layout, void* packed, void* stream);

/* activation stash written by the training forward and read by the backward; n_rows is
 * rounded up to the 128-row tile internally.  16-byte aligned. */
int64_t nerf_mlp_stash_bytes(int layout, int64_t n_rows);
int64_t nerf_mlp_dstash_bytes(int layout, int64_t n_rows);   /* backward scratch */

/* per-sample inputs: pos [n,3], dir [n,3] -> sigma_rgb [n,4] = (softplus density, sigmoid rgb).
 * stash == NULL: inference; otherwise the training forward (stash sized by the query). */
int nerf_mlp_forward(const void* packed, int layout, const float* pos, const float* dir,
                     int64_t n_rows, float* sigma_rgb, void* stash, void* stream);
/* ray inputs (fuses train.py:220-228): row r*S+s evaluates o[r] + d[r]*z[r,s], dir d[r]. */
int nerf_mlp_forward_rays(const void* packed, int layout, const float* rays_o,
                          const float* rays_d, const float* z, int64_t n_rays, int n_sample,
                          float* sigma_rgb, void* stash, void* stream);
/* Fused render of rays with 64 or 128 samples (NeRFModel.py:158-202 + train.py:248-269 in ONE kernel): the volume
 * compositing runs in the MLP kernel's epilogue -- a 128-row tile is exactly two rays or one -- so the per-sample (density, rgb)
 * never leave the SM; rgb [n_rays,3], depth / acc [n_rays] or NULL.  Bit-identical to nerf_mlp_forward_rays + nerf_composite_forward.
 * Other n_sample: NERF_ERR_ARG (use the two calls; nerf_render_image chooses by itself). */
int nerf_mlp_render_rays(const void* packed, int layout, const float* rays_o, const float* rays_d,
                         const float* z, int64_t n_rays, int n_sample, float* rgb, float* depth,
                         float* acc, void* stream);
/* backward: d_sigma_rgb [n,4] = dL/d(density, rgb) -> grads (flat fp32 parameter layout).
 * accumulate == 0 overwrites grads, != 0 adds to them.  dstash is scratch. */
int nerf_mlp_backward(const void* packed, int layout, const void* stash, const float* sigma_rgb,
                      const float* d_sigma_rgb, int64_t n_rows, void* dstash, float* grads,
                      int accumulate, void* stream);
/* the two halves of nerf_mlp_backward, callable (and timeable) on their own:
 *   dgrad: data-gradient chain through the layers; fills dstash with every layer's dY (bf16);
 *   wgrad: weight/bias gradients from (stash, dstash) into the flat fp32 gradient vector. */
int nerf_mlp_dgrad(const void* packed, int layout, const void* stash, const float* sigma_rgb,
                   const float* d_sigma_rgb, int64_t n_rows, void* dstash, void* stream);
int nerf_mlp_wgrad(int layout, const void* stash, const void* dstash, int64_t n_rows, float* grads,
                   int accumulate, void* stream);

/* ---------------------------------------------------------------------------------------
 * K5  volume compositing -- train.py:248-269 (+ depth/acc = sum w*z / sum w, which the
 * reference does not output) and its backward w.r.t. density and colour.
 *   sigma_rgb [n_rays,S,4], z [n_rays,S]; depth/acc may be NULL.
 * ------------------------------------------------------------------------------------- */
int nerf_composite_forward(const float* sigma_rgb, const float* z, int64_t n_rays, int n_sample,
                           float* rgb, float* depth, float* acc, void* stream);
int nerf_composite_backward(const float* sigma_rgb, const float* z, const float* g_rgb,
                            int64_t n_rays, int n_sample, float* d_sigma_rgb, void* stream);
/* Training form of K5 (train.py:248-269 + 460 + autograd): compositing forward, MSE over n_rays*3, its gradient and the
 * compositing backward in ONE launch: loss[0] = mean((rgb - target)^2), d_sigma_rgb[n,4] = grad_scale * dL/d(sigma, r, g, b).
 * The loss is reduced deterministically (per-block partial sums, last block adds them in order).  workspace:
 * nerf_composite_train_workspace_bytes(n_rays) bytes, zero-filled before the FIRST use (the kernel leaves it ready for the
 * next call).  Any n_sample (rays of up to 256 samples are kept in registers between the passes; longer rays are read twice). */
int64_t nerf_composite_train_workspace_bytes(int64_t n_rays);
int nerf_composite_train(const float* sigma_rgb, const float* z, const float* target, int64_t n_rays, int n_sample,
                         float grad_scale, float* loss, float* d_sigma_rgb, void* workspace, void* stream);

/* MSE over n_rays*3 (train.py:460): loss[0] = mean((rgb-target)^2) and, when g_rgb != NULL,
 * g_rgb = grad_scale * 2*(rgb-target)/(3*n_rays). */
int nerf_mse_loss(const float* rgb, const float* target, int64_t n_rays, float grad_scale,
                  float* loss, float* g_rgb, void* stream);

/* ---------------------------------------------------------------------------------------
 * Optimizer -- torch.optim.Adam defaults of train.py:401 over the flat vectors (one launch
 * instead of 26).  step is 1-based; grads are multiplied by grad_scale first (1/world for
 * data-parallel sums).  lr / beta1 / beta2 / eps are doubles like torch's Python scalars: the per-step
 * scalars (1 - beta1, 1 - beta2, lr / (1 - beta1^t), sqrt(1 - beta2^t)) are formed in fp64 and rounded
 * to fp32 once, exactly where torch.optim.Adam rounds them, and each element follows torch's
 * operation order -- parameters and moments agree with torch.optim.Adam to <= 2e-7 over a run.
 * ------------------------------------------------------------------------------------- */
int nerf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                   int64_t n, double lr, double beta1, double beta2, double eps, int64_t step,
                   float grad_scale, void* stream);

/* Data-parallel form of nerf_adam_step (train.py:401, 456-465 on N GPUs): the gradient exchange is fused into the optimiser
 * kernel.  Every rank's nerf_mlp_wgrad has written its local gradient into a buffer that is peer-mapped on all ranks
 * (symmetric memory); after a cross-rank barrier (the caller's) this kernel reads the gradient as the SUM over ranks --
 * multicast_grads != NULL: through the NVLink multicast address of the buffers (multimem.ld_reduce, the switch adds);
 * else: from peer_grads_host[0..n_ranks) (HOST array of device pointers, rank order) over NVLink -- and applies Adam with
 * grad_scale (= 1/n_ranks for the mean).  n must be a multiple of 4 and all pointers 16-byte aligned.  Every rank gets
 * bit-identical parameters.
 * The cross-rank barrier can be FUSED into the kernel: peer_flags_host[0..n_ranks) = every rank's array of 8 unsigned words
 * (symmetric memory, zero before the first exchange, peer-mapped like the gradient buffers; HOST array of device pointers),
 * rank = this rank, exchange_epoch = a device word holding the number of exchanges done so far (the kernel announces / waits for
 * *exchange_epoch + 1; the caller advances it after the call, nerf_train_step does so itself).  exchange_epoch == NULL: no barrier
 * inside, the caller's own barrier precedes the call. */
int nerf_adam_step_reduce(float* params, const float* const* peer_grads_host, int n_ranks, const float* multicast_grads,
                          float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1, double beta2, double eps,
                          int64_t step, float grad_scale, unsigned int* const* peer_flags_host, int rank,
                          const unsigned long long* exchange_epoch, void* stream);

/* ---------------------------------------------------------------------------------------
 * One frame (or a contiguous slice of its rays) by ONE call -- the body of test_image (train.py:735-757): K1 rays of ray ids
 * [ray_begin, ray_begin + ray_count) of pose `pose_index`, K2 depths (like nerf_sample_z with t_rand = NULL), K4 inference
 * forward, K5 compositing -> rgb [ray_count, 3] and optionally depth / acc [ray_count].  A multi-GPU host gives each rank its
 * parallel.shard_range slice; no collective.  workspace: nerf_render_workspace_bytes(ray_count, n_sample) bytes, 256-byte
 * aligned, the caller's (holds the rays, depths and per-sample sigma/rgb between the launches).
 * ------------------------------------------------------------------------------------- */
int64_t nerf_render_workspace_bytes(int64_t n_rays, int n_sample);
int nerf_render_image(const void* packed, int layout, const double* poses, int pose_index, double fx, double fy, double cx,
                      double cy, int width, int height, int64_t ray_begin, int64_t ray_count, int n_sample, float near,
                      float far, int perturb, uint64_t seed, uint64_t offset, float* rgb, float* depth, float* acc,
                      void* workspace, void* stream);

/* ---------------------------------------------------------------------------------------
 * One optimisation step (train.py:447-465: generateBatch -> render -> MSE -> backward -> Adam [StepLR] -> weight re-pack)
 * enqueued by ONE call, and replayable as a CUDA graph: nothing that changes from step to step is a launch argument.
 * The iteration counter lives in device memory (step_state[0], one unsigned 64-bit word, zero before the first step;
 * the last launch of the step advances it); the RNG counters (ray draw: (iteration * world + rank) * n_rays; depth jitter:
 * (iteration * world + rank) * ceil(n_rays * n_sample / 4)), the StepLR step size lr0 * lr_decay^(iteration / lr_decay_steps)
 * and Adam's bias corrections are derived from it on the device -- the same values the host forms above compute.
 *   batch9 == NULL : the batch is drawn like nerf_train_batch(seed, ...) from poses / images
 *   batch9 != NULL : [n_rays, 9] device array, a generateBatch() result (the caller refreshes it between steps)
 * The launches are exactly those of nerf_train_batch | nerf_unpack_batch, nerf_mlp_forward_rays (with stash),
 * nerf_composite_train (grad_scale 1), nerf_mlp_dgrad, nerf_mlp_wgrad (accumulate 0), nerf_adam_step(grad_scale),
 * nerf_mlp_pack_weights; every buffer is the caller's and must be sized as for those calls.  world / rank select the RNG
 * streams; with peer_grads / multicast_grads the optimiser is nerf_adam_step_reduce (see the phases below).
 * struct_bytes = sizeof(NerfTrainStep) (checked).
 * ------------------------------------------------------------------------------------- */
typedef struct NerfTrainStep {
  int64_t struct_bytes;
  int32_t layout, n_sample, perturb, world, rank, n_images, width, height;
  int64_t n_rays;
  double fx, fy, cx, cy;
  float near, far;
  uint64_t seed, z_seed;
  double lr0, lr_decay;
  int64_t lr_decay_steps;
  double beta1, beta2, eps;
  float grad_scale, pad_;
  const double* poses;
  const float* images;
  const float* batch9;
  float *rays_o, *rays_d, *target, *z, *sigma_rgb, *d_sigma_rgb, *loss;
  void* composite_ws;
  void *stash, *dstash, *packed;
  float *params, *grads, *exp_avg, *exp_avg_sq;
  unsigned long long* step_state;
  const float* peer_grads[8];      /* data parallel (world > 1), as nerf_adam_step_reduce: the ranks' gradient buffers ... */
  const float* multicast_grads;    /* ... or their NVLink multicast address; both NULL: plain Adam on `grads` */
  unsigned int* peer_flags[8];     /* data parallel: the ranks' barrier words (see nerf_adam_step_reduce) and ... */
  unsigned long long* exchange_epoch;  /* ... the device word counting the exchanges; NULL: the caller barriers before UPDATE */
} NerfTrainStep;
/* phases (bit mask): NERF_STEP_FORWARD = batch + render + loss (cfg->loss[0] is final when these launches complete);
 * NERF_STEP_GRADS = dgrad + wgrad (cfg->grads is final); NERF_STEP_UPDATE = Adam + re-pack + counter advance.
 * A host that wants the loss of every step (train.py:468) enqueues FORWARD, an async D2H copy of the loss + event, then
 * BACKWARD, and waits on the event only: the backward launches are still running while it prepares the next step.
 * A data-parallel host sets peer_grads / multicast_grads (the gradient sum over ranks is read through NVLink inside the Adam
 * kernel; grad_scale = 1/world) and either peer_flags + exchange_epoch -- then NERF_STEP_ALL is the whole step, the cross-rank
 * barrier included -- or enqueues FORWARD | GRADS, its own cross-rank barrier, then UPDATE. */
#define NERF_STEP_FORWARD 1
#define NERF_STEP_GRADS 2
#define NERF_STEP_UPDATE 4
#define NERF_STEP_BACKWARD 6
#define NERF_STEP_ALL 7
int nerf_train_step(const NerfTrainStep* cfg, int phases, void* stream);

/* ---------------------------------------------------------------------------------------
 * Self tests (device): exercise the tcgen05 descriptor encodings used by K4 on a small GEMM.
 * variant selects the operand layout hypothesis; result_host[0..3] receives max-abs-error
 * and diagnostics.  Returns 0 when the variant reproduces the fp32 reference product.
 * Not part of the hot path: synchronises the device and is the one entry point that allocates.
 * ------------------------------------------------------------------------------------- */
int nerf_selftest_umma(int variant, float* result_host);
#ifdef NERF_DIAGNOSTICS   /* only in libnerf_b200_diag.so (built with -DNERF_DIAGNOSTICS for tools/); not in the product library */
/* Diagnostics: when device_buffer != NULL, CTA 0 (the leader of cluster 0) of the fused MLP kernels writes clock64 stamps
 * [cluster iteration][step 0..11][tile slot 0..1][4] = (MMA issue start, MMA issue end, accumulator ready seen by the
 * epilogue, next-layer release) into it (int64; 8 iterations fit in 8*12*2*4 entries).  NULL disables.  tools/trace_mlp.py. */
int nerf_debug_trace(void* device_buffer);
/* Diagnostics: issue-rate microbenchmark of 128x256x16 UMMAs on one SM (variant: 0 one thread / one accumulator,
 * 1 one thread / two accumulators, 2 two threads / one accumulator each, 3 one thread / N=128 halves,
 * 4 two threads / same accumulator).  cycles_host[0] = SM cycles for n_mma MMAs, cycles_host[1] = n_mma.
 * Synchronises the device and allocates; not part of the hot path. */
int nerf_debug_umma_rate(int variant, int n_mma, long long* cycles_host);
/* Diagnostics: how fast can `warps` warps of ONE CTA per SM stream 512-byte warp stores (16 B per lane, the stash store pattern)
 * to HBM?  mode 0 st.global.cs from shared memory, 1 st.global, 2 st.global.cs from registers.  out_host[0] = max cycles over
 * CTAs, [1] = bytes per CTA, [2] = elapsed ns.  device_buffer must hold sms * iters * warps * 512 bytes. */
int nerf_debug_store_rate(int warps, int iters, int mode, void* device_buffer, long long buffer_bytes, long long* out_host);

#endif

#ifdef __cplusplus
}
#endif
#endif /* NERF_B200_H_ */
