// api.cu -- the C ABI of libnerf_b200.so (include/nerf_b200.h).  Argument checking, launch
// geometry and nothing else: every entry point enqueues sm_100a kernels on the caller's stream.
#include <atomic>
#include <map>
#include <mutex>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include <cuda.h>
#include "common.cuh"
#include "composite.cuh"
#include "mlp_fused.cuh"
#include "mlp_layout.cuh"
#include "rays.cuh"
#include "selftest.cuh"
#include "wgrad_pair.cuh"
#if !NERF_WGRAD_PAIR
#include "wgrad.cuh"            // the round-1 single-CTA kernel + CUDA-core head kernel: A/B builds only (-DNERF_WGRAD_PAIR=0)
#endif

using namespace nerf;

namespace {

struct DeviceInfo {
  int device = -1; int sms = 0; int cc_major = 0; int mlp_clusters = 0; bool attrs_set = false;
  // fork/join helpers of nerf_mlp_wgrad: the CUDA-core head-gradient kernel runs beside the tcgen05 kernel (stream-ordered)
  cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};
// Per host thread, one slot per device (indexed by device id): nothing is shared between host threads, so concurrent calls
// -- also first calls -- do not race, and switching the current device keeps the other devices' slots.  The side stream and
// the two events of a slot are destroyed when the thread exits.
constexpr int kMaxDevices = 64;
struct DeviceTable {
  DeviceInfo slot[kMaxDevices];
  ~DeviceTable() {
    for (DeviceInfo& d : slot) {
      if (d.ev_fork) (void)cudaEventDestroy(d.ev_fork);
      if (d.ev_join) (void)cudaEventDestroy(d.ev_join);
      if (d.side) (void)cudaStreamDestroy(d.side);      // errors (e.g. the runtime is already unloading at process exit) are ignored
    }
  }
};
thread_local DeviceTable g_devs;
#ifdef NERF_DIAGNOSTICS
std::atomic<long long*> g_trace{nullptr};      // nerf_debug_trace
#endif

int device_info(DeviceInfo** out) {
  int dev = 0;
  NERF_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return NERF_ERR_ARG;
  DeviceInfo& g_dev = g_devs.slot[dev];
  if (g_dev.device != dev) {
    NERF_CUDA_TRY(cudaDeviceGetAttribute(&g_dev.sms, cudaDevAttrMultiProcessorCount, dev));
    NERF_CUDA_TRY(cudaDeviceGetAttribute(&g_dev.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    g_dev.device = dev;
  }
  if (g_dev.cc_major != 10) return NERF_ERR_ARCH;
  if (!g_dev.attrs_set) {
    NERF_CUDA_TRY(cudaFuncSetAttribute(mlp_fused_kernel<MODE_FWD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MlpSmem::kTotal));
    NERF_CUDA_TRY(cudaFuncSetAttribute(mlp_fused_kernel<MODE_FWD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MlpSmem::kTotal));
    NERF_CUDA_TRY(cudaFuncSetAttribute(mlp_fused_kernel<MODE_BWD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MlpSmem::kTotal));
#if !NERF_WGRAD_PAIR
    NERF_CUDA_TRY(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes));
#endif
    NERF_CUDA_TRY(cudaFuncSetAttribute(wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kW2SmemBytes));
    NERF_CUDA_TRY(cudaFuncSetAttribute(selftest_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    NERF_CUDA_TRY(cudaFuncSetAttribute(selftest_umma_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    // how many CTA pairs of the fused MLP kernel the device can hold at once (one CTA per SM, pairs share a TPC)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(g_dev.sms & ~1); cfg.blockDim = dim3(kMlpThreads); cfg.dynamicSmemBytes = MlpSmem::kTotal;
    cudaLaunchAttribute at{};
    at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int n_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&n_clusters, mlp_fused_kernel<MODE_FWD, true>, &cfg) != cudaSuccess || n_clusters < 1) {
      (void)cudaGetLastError();
      n_clusters = g_dev.sms / 2;
    }
    g_dev.mlp_clusters = n_clusters;
    g_dev.attrs_set = true;
  }
  *out = &g_dev;
  return NERF_OK;
}

// persistent grid of the fused MLP kernels: CTA pairs, each iteration of a pair covers 4 tiles
inline int mlp_grid(const DeviceInfo& dev, int n_tiles) {
  const int n_quads = (n_tiles + 3) / 4;
  return 2 * (n_quads < dev.mlp_clusters ? n_quads : dev.mlp_clusters);
}

inline unsigned blocks_for(int64_t n, int per_block) { return (unsigned)((n + per_block - 1) / per_block); }

// every layout is built once, under the thread-safe initialisation of a function-local static (C++11)
struct LayoutCache {
  ModelLayout m[kNumLayouts];
  LayoutCache() { for (int i = 0; i < kNumLayouts; ++i) build_model_layout(i, m[i]); }
};
const ModelLayout* get_layout(int layout) {
  static const LayoutCache cache;
  if (layout >= 0 && layout < kNumLayouts) return &cache.m[layout];
  if (layout < 0 || !(layout & kLayoutCustomBit)) return nullptr;
  // other encoding frequencies (nerf_layout_id): built on first use, kept for the life of the process
  static std::mutex mu;
  static std::map<int, ModelLayout*> custom;
  std::lock_guard<std::mutex> lock(mu);
  auto it = custom.find(layout);
  if (it != custom.end()) return it->second;
  ModelLayout* m = new ModelLayout;
  if (!build_model_layout(layout, *m)) { delete m; return nullptr; }
  custom[layout] = m;
  return m;
}

// The weight images as a 2-D tensor for TMA: [rows][256] uint32 (1 KB rows); a box of 4 (2) rows is one CTA's half of an N = 256 (128)
// slab.  cuTensorMapEncodeTiled comes from the driver through the runtime (no link against libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_weight_maps(MlpArgs& a, const ModelLayout& m, const void* packed) {
  static const EncodeTiledFn encode = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  if (!encode) return NERF_ERR_ARCH;
  const cuuint64_t dims[2] = {256, (cuuint64_t)((m.fwd.image_bytes + m.bwd.image_bytes) / 1024)};
  const cuuint64_t strides[1] = {1024};
  const cuuint32_t estr[2] = {1, 1};
  for (int which = 0; which < 2; ++which) {
    const cuuint32_t box[2] = {256, which == 0 ? 4u : 2u};
    if (encode(which == 0 ? &a.tm4 : &a.tm2, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(packed), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return NERF_ERR_ARG;
  }
  return NERF_OK;
}

void fill_mlp_args(MlpArgs& a, const ModelLayout& m, const void* packed, int64_t n_rows) {
  std::memset(&a, 0, sizeof(a));
  a.packed = static_cast<const uint8_t*>(packed);
  a.small = reinterpret_cast<const float*>(a.packed + m.small_off);
  a.pos_L = m.pos_L; a.dir_L = m.dir_L; a.density_relu = m.density_relu; a.c_kcores = m.hidden / 8;
  a.sm_wd = m.sm_wd; a.sm_bd = m.sm_bd; a.sm_w2 = m.sm_w2; a.sm_b2 = m.sm_b2;
  a.n_rows = n_rows;
  a.n_tiles = (int)((n_rows + 127) / 128);
  a.stash_tile_bytes = m.stash_tile_bytes; a.dstash_tile_bytes = m.dstash_tile_bytes;
#ifdef NERF_DIAGNOSTICS
  a.trace = g_trace.load(std::memory_order_relaxed);
#endif
  a.st_pe = m.st_pe; a.st_de = m.st_de; a.st_r1 = m.st_r1; a.st_mask = m.st_mask; a.ds_y10 = m.ds_y[10]; a.ds_sr = m.ds_sr; a.ds_cot = m.ds_cot;
}

int launch_forward(const void* packed, int layout, const float* pos, const float* dir, const float* z,
                   int64_t n_rows, int n_sample, float* sigma_rgb, void* stash, void* stream,
                   float* comp_rgb = nullptr, float* comp_depth = nullptr, float* comp_acc = nullptr) {
  DeviceInfo* dev;
  int rc = device_info(&dev);
  if (rc != NERF_OK) return rc;
  const ModelLayout* m = get_layout(layout);
  if (!m) return NERF_ERR_LAYOUT;
  if (!packed || n_rows < 0 || n_rows > (int64_t)128 * 0x7fffff00) return NERF_ERR_ARG;
  if (n_rows == 0) return NERF_OK;                      /* empty batch: nothing to read or write */
  if (!pos || !dir || (!sigma_rgb && !comp_rgb)) return NERF_ERR_ARG;
  if (!nerf_aligned(packed, 16) || (sigma_rgb && !nerf_aligned(sigma_rgb, 16)) || (stash && !nerf_aligned(stash, 16))) return NERF_ERR_ALIGN;
  MlpArgs a;
  fill_mlp_args(a, *m, packed, n_rows);
  if ((rc = make_weight_maps(a, *m, packed))) return rc;
  a.pos = pos; a.dir = dir; a.z = z; a.n_sample = n_sample;
  a.sigma_rgb = sigma_rgb;
  a.comp_rgb = comp_rgb; a.comp_depth = comp_depth; a.comp_acc = comp_acc;
  a.stash = static_cast<uint8_t*>(stash);
  a.prog = m->fwd;
  const int grid = mlp_grid(*dev, a.n_tiles);
  if (stash) mlp_fused_kernel<MODE_FWD, true><<<grid, kMlpThreads, MlpSmem::kTotal, (cudaStream_t)stream>>>(a);
  else mlp_fused_kernel<MODE_FWD, false><<<grid, kMlpThreads, MlpSmem::kTotal, (cudaStream_t)stream>>>(a);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

void launch_adam(float* p, const float* g, float* m, float* v, int64_t n, AdamCoef c, float grad_scale, AdamSchedule sched, cudaStream_t s) {
  const bool vec = (n & 3) == 0 && nerf_aligned(p, 16) && nerf_aligned(g, 16) && nerf_aligned(m, 16) && nerf_aligned(v, 16);
  if (vec) adam_kernel<4><<<blocks_for(n / 4, 256), 256, 0, s>>>(p, g, m, v, n, c, grad_scale, sched);
  else adam_kernel<1><<<blocks_for(n, 256), 256, 0, s>>>(p, g, m, v, n, c, grad_scale, sched);
}

}  // namespace

extern "C" {

int nerf_abi_version(void) { return NERF_ABI_VERSION; }

const char* nerf_error_string(int code) {
  switch (code) {
    case NERF_OK: return "ok";
    case NERF_ERR_ARG: return "invalid argument";
    case NERF_ERR_ALIGN: return "misaligned pointer";
    case NERF_ERR_ARCH: return "device is not sm_100 (no CPU or other-arch fallback exists)";
    case NERF_ERR_LAYOUT: return "unknown model layout";
    case NERF_ERR_SELFTEST: return "self test failed";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}

int nerf_device_check(void) {
  DeviceInfo* dev;
  return device_info(&dev);
}

int nerf_layout_id(int embed_pos_L, int embed_direction_L, int pos_encoding, int density_relu, int num_channels) {
  const int id = make_layout_id(embed_pos_L, embed_direction_L, pos_encoding, density_relu, num_channels);
  return id < 0 ? NERF_ERR_LAYOUT : id;
}

int64_t nerf_param_count(int layout) {
  const ModelLayout* m = get_layout(layout);
  return m ? m->n_params : NERF_ERR_LAYOUT;
}

int64_t nerf_param_offset(int layout, int index) {
  const ModelLayout* m = get_layout(layout);
  if (!m) return NERF_ERR_LAYOUT;
  if (index < 0 || index >= 2 * kNumLinear) return NERF_ERR_ARG;
  return (index & 1) ? m->b_off[index >> 1] : m->w_off[index >> 1];
}

// ------------------------------------------------------------------------------------ K1
int nerf_raygen_image(const double* poses, int pose_index, double fx, double fy, double cx, double cy,
                      int width, int height, int64_t ray_begin, int64_t ray_count, float* rays_o,
                      float* rays_d, void* stream) {
  DeviceInfo* dev;
  int rc = device_info(&dev);
  if (rc != NERF_OK) return rc;
  if (!poses || !rays_o || !rays_d || pose_index < 0 || width <= 0 || height <= 0 || ray_begin < 0 || ray_count < 0 ||
      ray_begin + ray_count > (int64_t)width * height || fx == 0.0 || fy == 0.0)
    return NERF_ERR_ARG;
  if (ray_count == 0) return NERF_OK;
  raygen_image_kernel<<<blocks_for(ray_count, 256), 256, 0, (cudaStream_t)stream>>>(
      poses, pose_index, fx, fy, cx, cy, width, ray_begin, ray_count, rays_o, rays_d);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

int nerf_raygen_batch(const double* poses, int n_images, double fx, double fy, double cx, double cy,
                      int width, int height, const int32_t* img_idx, const int32_t* px, const int32_t* py,
                      const float* images, int64_t n_rays, float* rays_o, float* rays_d, float* target,
                      void* stream) {
  DeviceInfo* dev;
  int rc = device_info(&dev);
  if (rc != NERF_OK) return rc;
  if (!poses || !img_idx || !px || !py || !rays_o || !rays_d || n_images <= 0 || width <= 0 || height <= 0 ||
      n_rays < 0 || fx == 0.0 || fy == 0.0)
    return NERF_ERR_ARG;
  if (n_rays == 0) return NERF_OK;
  raygen_batch_kernel<<<blocks_for(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(
      poses, fx, fy, cx, cy, width, height, img_idx, px, py, images, n_rays, rays_o, rays_d, target);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

int nerf_draw_pixels(uint64_t seed, uint64_t offset, int n_images, int width, int height, int64_t n_rays,
                     int32_t* img_idx, int32_t* px, int32_t* py, void* stream) {
  DeviceInfo* dev;
  int rc = device_info(&dev);
  if (rc != NERF_OK) return rc;
  if (!img_idx || !px || !py || n_images <= 0 || width <= 0 || height <= 0 || n_rays < 0) return NERF_ERR_ARG;
  if (n_rays == 0) return NERF_OK;
  draw_pixels_kernel<<<blocks_for(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(seed, offset, n_images, width, height,
                                                                                n_rays, img_idx, px, py);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

int nerf_train_batch(const double* poses, int n_images, double fx, double fy, double cx, double cy, int width, int height,
                     const float* images, uint64_t seed, uint64_t ray_offset, int64_t n_rays, float* rays_o, float* rays_d,
                     float* target, int n_sample, float near, float far, int perturb, uint64_t z_seed, uint64_t z_offset,
                     float* z, void* stream) {
  DeviceInfo* dev;
  int rc = device_info(&dev);
  if (rc != NERF_OK) return rc;
  if (!poses || !images || !rays_o || !rays_d || !target || !z || n_images <= 0 || width <= 0 || height <= 0 || n_rays < 0 ||
      n_sample <= 0 || fx == 0.0 || fy == 0.0)
    return NERF_ERR_ARG;
  if (n_rays == 0) return NERF_OK;
  const unsigned ray_blocks = blocks_for(n_rays, 256), z_blocks = blocks_for((n_rays * n_sample + 3) / 4, 256);
  train_batch_kernel<<<ray_blocks + z_blocks, 256, 0, (cudaStream_t)stream>>>(nullptr, 1, 0, nullptr, poses, fx, fy, cx, cy, width, height, n_images, images, seed,
                                                                             ray_offset, n_rays, rays_o, rays_d, target, n_sample, near,
                                                                             far, perturb, z_seed, z_offset, z, ray_blocks);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

int nerf_unpack_batch(const float* batch9, int64_t n_rays, float* rays_o, float* rays_d, float* target, int n_sample, float near,
                      float far, int perturb, uint64_t z_seed, uint64_t z_offset, float* z, void* stream) {
  DeviceInfo* dev;
  int rc = device_info(&dev);
  if (rc != NERF_OK) return rc;
  if (!batch9 || !rays_o || !rays_d || !target || !z || n_rays < 0 || n_sample <= 0) return NERF_ERR_ARG;
  if (n_rays == 0) return NERF_OK;
  const unsigned ray_blocks = blocks_for(n_rays, 256), z_blocks = blocks_for((n_rays * n_sample + 3) / 4, 256);
  train_batch_kernel<<<ray_blocks + z_blocks, 256, 0, (cudaStream_t)stream>>>(nullptr, 1, 0, batch9, nullptr, 1.0, 1.0, 0.0, 0.0, 1, 1, 1, nullptr, 0, 0, n_rays,
                                                                             rays_o, rays_d, target, n_sample, near, far, perturb, z_seed,
                                                                             z_offset, z, ray_blocks);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

// ------------------------------------------------------------------------------------ K2
int nerf_sample_z(int64_t n_rays, int n_sample, float near, float far, int perturb, const float* t_rand,
                  uint64_t seed, uint64_t offset, float* z, void* stream) {
  DeviceInfo* dev;
  int rc = device_info(&dev);
  if (rc != NERF_OK) return rc;
  if (!z || n_rays < 0 || n_sample <= 0) return NERF_ERR_ARG;
  const int64_t total = n_rays * n_sample;
  if (total == 0) return NERF_OK;
  sample_z_kernel<<<blocks_for((total + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(total, n_sample, near, far, perturb,
                                                                           t_rand, seed, offset, z);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

// ------------------------------------------------------------------------------------ K3
int nerf_encode(const float* x, int64_t n, int L, float* out, void* stream) {
  DeviceInfo* dev;
  int rc = device_info(&dev);
  if (rc != NERF_OK) return rc;
  if (!x || !out || n < 0 || L < 0 || L > 30) return NERF_ERR_ARG;
  const int64_t total = n * 3 * (2 * L + 1);
  if (total == 0) return NERF_OK;
  encode_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, n, L, out);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

// ------------------------------------------------------------------------------------ K4
int64_t nerf_mlp_packed_bytes(int layout) {
  const ModelLayout* m = get_layout(layout);
  return m ? m->packed_bytes : NERF_ERR_LAYOUT;
}

static int launch_pack(const float* params, int layout, void* packed, unsigned long long* bump, unsigned long long* bump2, void* stream) {
  DeviceInfo* dev;
  int rc = device_info(&dev);
  if (rc != NERF_OK) return rc;
  const ModelLayout* m = get_layout(layout);
  if (!m) return NERF_ERR_LAYOUT;
  if (!params || !packed) return NERF_ERR_ARG;
  if (!nerf_aligned(packed, 16)) return NERF_ERR_ALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  PackSmallArgs p{};
  p.hidden = m->hidden;
  p.wd_src = m->w_off[L_DENS]; p.bd_src = m->b_off[L_DENS]; p.w2_src = m->w_off[L_RGB2]; p.b2_src = m->b_off[L_RGB2];
  p.sm_wd = m->sm_wd; p.sm_bd = m->sm_bd; p.sm_w2 = m->sm_w2; p.sm_b2 = m->sm_b2;
  // one launch: grid.y = forward steps, dgrad steps, then one row of blocks for the fp32 head table
  pack_all_kernel<<<dim3(64, m->fwd.n_steps + m->bwd.n_steps + 1), 256, 0, s>>>(
      params, m->fwd, m->bwd, p, static_cast<uint8_t*>(packed), reinterpret_cast<float*>(static_cast<uint8_t*>(packed) + m->small_off),
      bump, bump2);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

int nerf_mlp_pack_weights(const float* params, int layout, void* packed, void* stream) {
  return launch_pack(params, layout, packed, nullptr, nullptr, stream);
}

int64_t nerf_mlp_stash_bytes(int layout, int64_t n_rows) {
  const ModelLayout* m = get_layout(layout);
  if (!m) return NERF_ERR_LAYOUT;
  if (n_rows < 0) return NERF_ERR_ARG;
  return ((n_rows + 127) / 128) * (int64_t)m->stash_tile_bytes;
}

int64_t nerf_mlp_dstash_bytes(int layout, int64_t n_rows) {
  const ModelLayout* m = get_layout(layout);
  if (!m) return NERF_ERR_LAYOUT;
  if (n_rows < 0) return NERF_ERR_ARG;
  return ((n_rows + 127) / 128) * (int64_t)m->dstash_tile_bytes;
}

int nerf_mlp_forward(const void* packed, int layout, const float* pos, const float* dir, int64_t n_rows,
                     float* sigma_rgb, void* stash, void* stream) {
  return launch_forward(packed, layout, pos, dir, nullptr, n_rows, 0, sigma_rgb, stash, stream);
}

int nerf_mlp_forward_rays(const void* packed, int layout, const float* rays_o, const float* rays_d, const float* z,
                          int64_t n_rays, int n_sample, float* sigma_rgb, void* stash, void* stream) {
  if (!z || n_sample <= 0 || n_rays < 0) return NERF_ERR_ARG;
  return launch_forward(packed, layout, rays_o, rays_d, z, n_rays * n_sample, n_sample, sigma_rgb, stash, stream);
}

int nerf_mlp_render_rays(const void* packed, int layout, const float* rays_o, const float* rays_d, const float* z, int64_t n_rays,
                         int n_sample, float* rgb, float* depth, float* acc, void* stream) {
  if (!z || !rgb || n_rays < 0) return NERF_ERR_ARG;
  if (n_sample != 64 && n_sample != 128) return NERF_ERR_ARG;      /* a 128-row tile must be exactly two rays or one */
  return launch_forward(packed, layout, rays_o, rays_d, z, n_rays * n_sample, n_sample, nullptr, nullptr, stream, rgb, depth, acc);
}

int nerf_mlp_dgrad(const void* packed, int layout, const void* stash, const float* sigma_rgb,
                   const float* d_sigma_rgb, int64_t n_rows, void* dstash, void* stream) {
  DeviceInfo* dev;
  int rc = device_info(&dev);
  if (rc != NERF_OK) return rc;
  const ModelLayout* m = get_layout(layout);
  if (!m) return NERF_ERR_LAYOUT;
  if (!packed || !stash || !sigma_rgb || !d_sigma_rgb || !dstash || n_rows < 0) return NERF_ERR_ARG;
  if (!nerf_aligned(packed, 16) || !nerf_aligned(stash, 16) || !nerf_aligned(dstash, 16) || !nerf_aligned(sigma_rgb, 16) ||
      !nerf_aligned(d_sigma_rgb, 16))
    return NERF_ERR_ALIGN;
  if (n_rows == 0) return NERF_OK;
  MlpArgs a;
  fill_mlp_args(a, *m, packed, n_rows);
  if ((rc = make_weight_maps(a, *m, packed))) return rc;
  a.sigma_rgb = const_cast<float*>(sigma_rgb);
  a.d_sigma_rgb = d_sigma_rgb;
  a.stash = static_cast<uint8_t*>(const_cast<void*>(stash));
  a.dstash = static_cast<uint8_t*>(dstash);
  a.prog = m->bwd;
  const int grid = mlp_grid(*dev, a.n_tiles);
  mlp_fused_kernel<MODE_BWD, false><<<grid, kMlpThreads, MlpSmem::kTotal, (cudaStream_t)stream>>>(a);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

int nerf_mlp_wgrad(int layout, const void* stash, const void* dstash, int64_t n_rows, float* grads, int accumulate,
                   void* stream) {
  DeviceInfo* dev;
  int rc = device_info(&dev);
  if (rc != NERF_OK) return rc;
  const ModelLayout* m = get_layout(layout);
  if (!m) return NERF_ERR_LAYOUT;
  if (!stash || !dstash || !grads || n_rows < 0) return NERF_ERR_ARG;
  if (!nerf_aligned(stash, 16) || !nerf_aligned(dstash, 16)) return NERF_ERR_ALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  if (!accumulate) NERF_CUDA_TRY(cudaMemsetAsync(grads, 0, sizeof(float) * m->n_params, s));
  if (n_rows == 0) return NERF_OK;
  const int n_tiles = (int)((n_rows + 127) / 128);

#if NERF_WGRAD_PAIR
  {
    // CTA pairs (wgrad_pair.cuh).  All (item, tile) work is laid on one line in item order, weighted by the item's bytes per
    // tile, and cut into n_pairs equal pieces: a pair gets a tile range of one item, or the tail of one and the head of the next.
    W2PairArgs w;
    std::memset(&w, 0, sizeof(w));
    w.stash = static_cast<const uint8_t*>(stash); w.dstash = static_cast<const uint8_t*>(dstash); w.grads = grads;
    w.stash_tile_bytes = m->stash_tile_bytes; w.dstash_tile_bytes = m->dstash_tile_bytes;
    w.n_items = m->n_w2;
    for (int i = 0; i < m->n_w2; ++i) w.items[i] = m->w2[i];
    const int n_pairs = w2_partition(*m, n_tiles, dev->sms / 2, w.n_units, w.unit);
    if (n_pairs < 1) return NERF_ERR_ARG;
    w.n_pairs = n_pairs;
    wgrad_pair_kernel<<<2 * n_pairs, kW2Threads, kW2SmemBytes, s>>>(w);
    NERF_LAUNCH_CHECK();
    return NERF_OK;
  }
#else
  if (m->n_wgrad == 0) return NERF_ERR_LAYOUT;      /* the single-CTA kernel is built for 256 channels */
  // weight gradients: split the SMs over the work items in proportion to their cost
  WgradArgs w;
  std::memset(&w, 0, sizeof(w));
  w.stash = static_cast<const uint8_t*>(stash); w.dstash = static_cast<const uint8_t*>(dstash); w.grads = grads;
  w.stash_tile_bytes = m->stash_tile_bytes; w.dstash_tile_bytes = m->dstash_tile_bytes;
  w.n_tiles = n_tiles; w.n_items = m->n_wgrad;
  // Split the SMs over the items so that the LARGEST per-CTA share of an item's cost is as small as possible (the kernel
  // ends with its slowest CTA): start from one CTA per item and keep giving a CTA to the item with the largest cost / CTAs.
  int n_cta[kMaxWgradItems];
  int used = 0;
  for (int i = 0; i < m->n_wgrad; ++i) { w.items[i] = m->wg[i]; n_cta[i] = 1; ++used; }
  while (used < dev->sms) {
    int best = -1;
    double worst = 0.0;
    for (int i = 0; i < m->n_wgrad; ++i) {
      const double per = (double)m->wg[i].cost / n_cta[i];
      if (n_cta[i] < n_tiles && per > worst) { worst = per; best = i; }
    }
    if (best < 0) break;                         /* every item already has one CTA per tile */
    ++n_cta[best]; ++used;
  }
  int ctas = 0;
  for (int i = 0; i < m->n_wgrad; ++i) {
    w.cta_begin[i] = ctas;
    ctas += n_cta[i];
  }
  w.cta_begin[m->n_wgrad] = ctas;
  // head gradients (density_layer, rgb_layer2) on CUDA cores.  Both kernels are HBM-bound streams over the same stashes and
  // fit on an SM together (no shared memory / 21 K registers here, one 197 KB CTA there), so the small one is forked onto a
  // side stream after the gradient memset and joined back: everything stays ordered on the caller's stream.
  HeadGradArgs h{};
  h.stash = w.stash; h.dstash = w.dstash; h.grads = grads;
  h.stash_tile_bytes = m->stash_tile_bytes; h.dstash_tile_bytes = m->dstash_tile_bytes; h.n_tiles = n_tiles;
  h.st_h8 = m->st_h[8]; h.st_r1 = m->st_r1; h.ds_sr = m->ds_sr;
  h.g_wd = m->w_off[L_DENS]; h.g_bd = m->b_off[L_DENS]; h.g_w2 = m->w_off[L_RGB2]; h.g_b2 = m->b_off[L_RGB2];
  // few blocks: every block ends with one atomic per gradient element (644 addresses), which serialise per address in L2
  const int hgrid = n_tiles < 2 * dev->sms ? n_tiles : 2 * dev->sms;
  if (!dev->side) {
    NERF_CUDA_TRY(cudaStreamCreateWithFlags(&dev->side, cudaStreamNonBlocking));
    NERF_CUDA_TRY(cudaEventCreateWithFlags(&dev->ev_fork, cudaEventDisableTiming));
    NERF_CUDA_TRY(cudaEventCreateWithFlags(&dev->ev_join, cudaEventDisableTiming));
  }
  NERF_CUDA_TRY(cudaEventRecord(dev->ev_fork, s));
  NERF_CUDA_TRY(cudaStreamWaitEvent(dev->side, dev->ev_fork, 0));
  head_grad_kernel<<<hgrid, 384, 0, dev->side>>>(h);
  NERF_LAUNCH_CHECK();
  NERF_CUDA_TRY(cudaEventRecord(dev->ev_join, dev->side));
  wgrad_kernel<<<ctas, 192, kWgSmemBytes, s>>>(w);
  NERF_LAUNCH_CHECK();
  NERF_CUDA_TRY(cudaStreamWaitEvent(s, dev->ev_join, 0));
  return NERF_OK;
#endif
}

int nerf_mlp_backward(const void* packed, int layout, const void* stash, const float* sigma_rgb,
                      const float* d_sigma_rgb, int64_t n_rows, void* dstash, float* grads, int accumulate,
                      void* stream) {
  if (!grads) return NERF_ERR_ARG;
  int rc = nerf_mlp_dgrad(packed, layout, stash, sigma_rgb, d_sigma_rgb, n_rows, dstash, stream);
  if (rc != NERF_OK) return rc;
  return nerf_mlp_wgrad(layout, stash, dstash, n_rows, grads, accumulate, stream);
}

// ------------------------------------------------------------------------------------ K5
int nerf_composite_forward(const float* sigma_rgb, const float* z, int64_t n_rays, int n_sample, float* rgb,
                           float* depth, float* acc, void* stream) {
  DeviceInfo* dev;
  int rc = device_info(&dev);
  if (rc != NERF_OK) return rc;
  if (!sigma_rgb || !z || !rgb || n_rays < 0 || n_sample <= 0) return NERF_ERR_ARG;
  if (!nerf_aligned(sigma_rgb, 16)) return NERF_ERR_ALIGN;
  if (n_rays == 0) return NERF_OK;
  const unsigned cf_blocks = blocks_for(n_rays, kCompWarpsPerBlock);
  const float4* sr4 = reinterpret_cast<const float4*>(sigma_rgb);
  cudaStream_t cs = (cudaStream_t)stream;
  switch ((n_sample & 31) == 0 ? n_sample >> 5 : 0) {            // full-chunk ray lengths get the straight-line kernels
    case 1: composite_forward_exact_kernel<1><<<cf_blocks, kCompWarpsPerBlock * 32, 0, cs>>>(sr4, z, n_rays, rgb, depth, acc); break;
    case 2: composite_forward_exact_kernel<2><<<cf_blocks, kCompWarpsPerBlock * 32, 0, cs>>>(sr4, z, n_rays, rgb, depth, acc); break;
    case 4: composite_forward_exact_kernel<4><<<cf_blocks, kCompWarpsPerBlock * 32, 0, cs>>>(sr4, z, n_rays, rgb, depth, acc); break;
    case 6: composite_forward_exact_kernel<6><<<cf_blocks, kCompWarpsPerBlock * 32, 0, cs>>>(sr4, z, n_rays, rgb, depth, acc); break;
    case 8: composite_forward_exact_kernel<8><<<cf_blocks, kCompWarpsPerBlock * 32, 0, cs>>>(sr4, z, n_rays, rgb, depth, acc); break;
    default: composite_forward_kernel<<<cf_blocks, kCompWarpsPerBlock * 32, 0, cs>>>(sr4, z, n_rays, n_sample, rgb, depth, acc);
  }
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

// ------------------------------------------------------------------------------------ whole frame (slice)
static inline int64_t align256(int64_t v) { return (v + 255) & ~(int64_t)255; }

int64_t nerf_render_workspace_bytes(int64_t n_rays, int n_sample) {
  if (n_rays < 0 || n_sample <= 0) return NERF_ERR_ARG;
  return 2 * align256(n_rays * 12) + align256(n_rays * n_sample * 4) + align256(n_rays * n_sample * 16);
}

int nerf_render_image(const void* packed, int layout, const double* poses, int pose_index, double fx, double fy, double cx,
                      double cy, int width, int height, int64_t ray_begin, int64_t ray_count, int n_sample, float near,
                      float far, int perturb, uint64_t seed, uint64_t offset, float* rgb, float* depth, float* acc,
                      void* workspace, void* stream) {
  if (ray_count < 0 || n_sample <= 0) return NERF_ERR_ARG;
  if (ray_count == 0) return NERF_OK;
  if (!workspace) return NERF_ERR_ARG;
  if (!nerf_aligned(workspace, 256)) return NERF_ERR_ALIGN;
  uint8_t* w = static_cast<uint8_t*>(workspace);
  float* o = reinterpret_cast<float*>(w);
  float* d = reinterpret_cast<float*>(w + align256(ray_count * 12));
  float* z = reinterpret_cast<float*>(w + 2 * align256(ray_count * 12));
  float* sr = reinterpret_cast<float*>(w + 2 * align256(ray_count * 12) + align256(ray_count * n_sample * 4));
  int rc;
  if ((rc = nerf_raygen_image(poses, pose_index, fx, fy, cx, cy, width, height, ray_begin, ray_count, o, d, stream))) return rc;
  if ((rc = nerf_sample_z(ray_count, n_sample, near, far, perturb, nullptr, seed, offset, z, stream))) return rc;
  // 64 / 128 samples per ray: the compositing runs in the MLP kernel's epilogue (sigma_rgb never leaves the SM); else two kernels
  if ((n_sample == 64 || n_sample == 128) && rgb) return nerf_mlp_render_rays(packed, layout, o, d, z, ray_count, n_sample, rgb, depth, acc, stream);
  if ((rc = nerf_mlp_forward_rays(packed, layout, o, d, z, ray_count, n_sample, sr, nullptr, stream))) return rc;
  return nerf_composite_forward(sr, z, ray_count, n_sample, rgb, depth, acc, stream);
}

int nerf_composite_backward(const float* sigma_rgb, const float* z, const float* g_rgb, int64_t n_rays, int n_sample,
                            float* d_sigma_rgb, void* stream) {
  DeviceInfo* dev;
  int rc = device_info(&dev);
  if (rc != NERF_OK) return rc;
  if (!sigma_rgb || !z || !g_rgb || !d_sigma_rgb || n_rays < 0 || n_sample <= 0) return NERF_ERR_ARG;
  if (!nerf_aligned(sigma_rgb, 16) || !nerf_aligned(d_sigma_rgb, 16)) return NERF_ERR_ALIGN;
  if (n_rays == 0) return NERF_OK;
  // instances unrolled for 2 / 4 / 8 chunks of 32 samples: a 64-sample ray does not pay the registers of a 256-sample one
  const unsigned cb_blocks = blocks_for(n_rays, kCompWarpsPerBlock);
  if (n_sample <= 64) composite_backward_kernel<2><<<cb_blocks, kCompWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(sigma_rgb), z, g_rgb, n_rays, n_sample, reinterpret_cast<float4*>(d_sigma_rgb));
  else if (n_sample <= 128) composite_backward_kernel<4><<<cb_blocks, kCompWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(sigma_rgb), z, g_rgb, n_rays, n_sample, reinterpret_cast<float4*>(d_sigma_rgb));
  else if (n_sample <= 32 * kCompMaxChunks) composite_backward_kernel<8><<<cb_blocks, kCompWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(sigma_rgb), z, g_rgb, n_rays, n_sample, reinterpret_cast<float4*>(d_sigma_rgb));
  else composite_backward_long_kernel<false><<<cb_blocks, kCompWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(       /* two-pass form: any ray length */
      reinterpret_cast<const float4*>(sigma_rgb), z, g_rgb, n_rays, n_sample, 1.0f, nullptr, nullptr, nullptr, reinterpret_cast<float4*>(d_sigma_rgb));
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

int64_t nerf_composite_train_workspace_bytes(int64_t n_rays) {
  if (n_rays < 0) return NERF_ERR_ARG;
  return (int64_t)sizeof(float) * (blocks_for(n_rays, kCompWarpsPerBlock) + 1);
}

int nerf_composite_train(const float* sigma_rgb, const float* z, const float* target, int64_t n_rays, int n_sample,
                         float grad_scale, float* loss, float* d_sigma_rgb, void* workspace, void* stream) {
  DeviceInfo* dev;
  int rc = device_info(&dev);
  if (rc != NERF_OK) return rc;
  if (!sigma_rgb || !z || !target || !loss || !d_sigma_rgb || !workspace || n_rays <= 0 || n_sample <= 0) return NERF_ERR_ARG;
  if (!nerf_aligned(sigma_rgb, 16) || !nerf_aligned(d_sigma_rgb, 16) || !nerf_aligned(workspace, 4)) return NERF_ERR_ALIGN;
  const unsigned blocks = blocks_for(n_rays, kCompWarpsPerBlock);
  float* partial = static_cast<float*>(workspace);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(partial + blocks);      /* zero before first use; the kernel re-zeroes it */
  if (n_sample <= 64) composite_train_kernel<2><<<blocks, kCompWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(sigma_rgb), z, target, n_rays, n_sample, grad_scale, loss, partial, ticket,
      reinterpret_cast<float4*>(d_sigma_rgb));
  else if (n_sample <= 128) composite_train_kernel<4><<<blocks, kCompWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(sigma_rgb), z, target, n_rays, n_sample, grad_scale, loss, partial, ticket,
      reinterpret_cast<float4*>(d_sigma_rgb));
  else if (n_sample <= 32 * kCompMaxChunks) composite_train_kernel<8><<<blocks, kCompWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(sigma_rgb), z, target, n_rays, n_sample, grad_scale, loss, partial, ticket,
      reinterpret_cast<float4*>(d_sigma_rgb));
  else composite_backward_long_kernel<true><<<blocks, kCompWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(           /* two-pass form: any ray length */
      reinterpret_cast<const float4*>(sigma_rgb), z, target, n_rays, n_sample, grad_scale, loss, partial, ticket,
      reinterpret_cast<float4*>(d_sigma_rgb));
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

int nerf_mse_loss(const float* rgb, const float* target, int64_t n_rays, float grad_scale, float* loss, float* g_rgb,
                  void* stream) {
  DeviceInfo* dev;
  int rc = device_info(&dev);
  if (rc != NERF_OK) return rc;
  if (!rgb || !target || !loss || n_rays <= 0) return NERF_ERR_ARG;
  mse_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(rgb, target, n_rays * 3, grad_scale, loss, g_rgb);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

// ------------------------------------------------------------------------------------ optimizer
int nerf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                   double beta1, double beta2, double eps, int64_t step, float grad_scale, void* stream) {
  DeviceInfo* dev;
  int rc = device_info(&dev);
  if (rc != NERF_OK) return rc;
  if (!params || !grads || !exp_avg || !exp_avg_sq || n < 0 || step < 1) return NERF_ERR_ARG;
  if (n == 0) return NERF_OK;
  launch_adam(params, grads, exp_avg, exp_avg_sq, n, adam_coef(lr, beta1, beta2, eps, (double)step), grad_scale, AdamSchedule{},
              (cudaStream_t)stream);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

// ------------------------------------------------------------------------------------ whole train step
int nerf_train_step(const NerfTrainStep* c, int phases, void* stream) {
  DeviceInfo* dev;
  int rc = device_info(&dev);
  if (rc != NERF_OK) return rc;
  if (!c || c->struct_bytes != (int64_t)sizeof(NerfTrainStep) || phases < 1 || phases > NERF_STEP_ALL) return NERF_ERR_ARG;
  const ModelLayout* m = get_layout(c->layout);
  if (!m) return NERF_ERR_LAYOUT;
  if (c->n_rays <= 0 || c->n_sample <= 0 || c->world < 1 || c->rank < 0 || c->rank >= c->world || c->lr_decay_steps < 1 ||
      !c->step_state || !c->rays_o || !c->rays_d || !c->target || !c->z || !c->sigma_rgb || !c->d_sigma_rgb || !c->loss ||
      !c->composite_ws || !c->stash || !c->dstash || !c->packed || !c->params || !c->grads || !c->exp_avg || !c->exp_avg_sq)
    return NERF_ERR_ARG;
  if (!c->batch9 && (!c->poses || !c->images || c->n_images <= 0 || c->width <= 0 || c->height <= 0 || c->fx == 0.0 || c->fy == 0.0))
    return NERF_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t rows = c->n_rays * c->n_sample;
  const unsigned ray_blocks = blocks_for(c->n_rays, 256), z_blocks = blocks_for((rows + 3) / 4, 256);
  if (phases & NERF_STEP_FORWARD) {
  train_batch_kernel<<<ray_blocks + z_blocks, 256, 0, s>>>(c->step_state, (unsigned)c->world, (unsigned)c->rank, c->batch9, c->poses, c->fx,
                                                          c->fy, c->cx, c->cy, c->width, c->height, c->n_images, c->images, c->seed, 0,
                                                          c->n_rays, c->rays_o, c->rays_d, c->target, c->n_sample, c->near, c->far,
                                                          c->perturb, c->z_seed, 0, c->z, ray_blocks);
  NERF_LAUNCH_CHECK();
  if ((rc = nerf_mlp_forward_rays(c->packed, c->layout, c->rays_o, c->rays_d, c->z, c->n_rays, c->n_sample, c->sigma_rgb, c->stash, stream)))
    return rc;
  if ((rc = nerf_composite_train(c->sigma_rgb, c->z, c->target, c->n_rays, c->n_sample, 1.0f, c->loss, c->d_sigma_rgb, c->composite_ws, stream)))
    return rc;
  }
  if (phases & NERF_STEP_GRADS) {
    if ((rc = nerf_mlp_dgrad(c->packed, c->layout, c->stash, c->sigma_rgb, c->d_sigma_rgb, rows, c->dstash, stream))) return rc;
    if ((rc = nerf_mlp_wgrad(c->layout, c->stash, c->dstash, rows, c->grads, 0, stream))) return rc;
  }
  if (!(phases & NERF_STEP_UPDATE)) return NERF_OK;
  const AdamSchedule sched{c->step_state, c->lr0, c->lr_decay, c->beta1, c->beta2, c->eps, (long long)c->lr_decay_steps};
  if (c->world > 1 && (c->multicast_grads || c->peer_grads[0])) {
    // data parallel: the gradient sum over ranks is read through NVLink inside the optimiser kernel (nerf_adam_step_reduce)
    if (c->world > 8 || (m->n_params & 3) || !nerf_aligned(c->params, 16) || !nerf_aligned(c->exp_avg, 16) ||
        !nerf_aligned(c->exp_avg_sq, 16) || (c->multicast_grads && !nerf_aligned(c->multicast_grads, 16)))
      return NERF_ERR_ALIGN;
    PeerGrads peers{};
    for (int r = 0; r < c->world && !c->multicast_grads; ++r) {
      if (!c->peer_grads[r] || !nerf_aligned(c->peer_grads[r], 16)) return NERF_ERR_ALIGN;
      peers.ptr[r] = c->peer_grads[r];
    }
    PeerFlags flags{};
    const bool fused_barrier = c->exchange_epoch != nullptr;
    for (int r = 0; r < c->world && fused_barrier; ++r) {
      if (!c->peer_flags[r] || !nerf_aligned(c->peer_flags[r], 4)) return NERF_ERR_ALIGN;
      flags.ptr[r] = c->peer_flags[r];
    }
    adam_reduce_kernel<<<blocks_for(m->n_params / 4, 256), 256, 0, s>>>(c->params, peers, c->world, c->multicast_grads, c->exp_avg,
                                                                        c->exp_avg_sq, m->n_params / 4, AdamCoef{}, c->grad_scale, sched,
                                                                        flags, c->rank, fused_barrier ? c->exchange_epoch : nullptr);
  } else {
    launch_adam(c->params, c->grads, c->exp_avg, c->exp_avg_sq, m->n_params, AdamCoef{}, c->grad_scale, sched, s);
  }
  NERF_LAUNCH_CHECK();
  return launch_pack(c->params, c->layout, c->packed, c->step_state, (c->world > 1) ? c->exchange_epoch : nullptr, stream);
}

int nerf_adam_step_reduce(float* params, const float* const* peer_grads_host, int n_ranks, const float* multicast_grads,
                          float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1, double beta2, double eps, int64_t step,
                          float grad_scale, unsigned int* const* peer_flags_host, int rank, const unsigned long long* exchange_epoch,
                          void* stream) {
  DeviceInfo* dev;
  int rc = device_info(&dev);
  if (rc != NERF_OK) return rc;
  if (!params || !exp_avg || !exp_avg_sq || n < 0 || step < 1 || n_ranks < 1 || n_ranks > 8 || (n & 3)) return NERF_ERR_ARG;
  if (!multicast_grads && !peer_grads_host) return NERF_ERR_ARG;
  PeerGrads peers{};
  for (int r = 0; r < n_ranks && peer_grads_host; ++r) {
    if (!peer_grads_host[r] || !nerf_aligned(peer_grads_host[r], 16)) return NERF_ERR_ALIGN;
    peers.ptr[r] = peer_grads_host[r];
  }
  if (!nerf_aligned(params, 16) || !nerf_aligned(exp_avg, 16) || !nerf_aligned(exp_avg_sq, 16) ||
      (multicast_grads && !nerf_aligned(multicast_grads, 16)))
    return NERF_ERR_ALIGN;
  if (n == 0) return NERF_OK;
  PeerFlags flags{};
  if (exchange_epoch != nullptr) {
    if (!peer_flags_host || rank < 0 || rank >= n_ranks) return NERF_ERR_ARG;
    for (int r = 0; r < n_ranks; ++r) {
      if (!peer_flags_host[r] || !nerf_aligned(peer_flags_host[r], 4)) return NERF_ERR_ALIGN;
      flags.ptr[r] = peer_flags_host[r];
    }
  }
  adam_reduce_kernel<<<blocks_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(params, peers, n_ranks, multicast_grads, exp_avg, exp_avg_sq,
                                                                             n / 4, adam_coef(lr, beta1, beta2, eps, (double)step),
                                                                             grad_scale, AdamSchedule{}, flags, rank, exchange_epoch);
  NERF_LAUNCH_CHECK();
  return NERF_OK;
}

// ------------------------------------------------------------------------------------ self test
int nerf_selftest_umma(int variant, float* result_host) {
  DeviceInfo* dev;
  int rc = device_info(&dev);
  if (rc != NERF_OK) return rc;
  if (variant < 0 || variant > 2 || !result_host) return NERF_ERR_ARG;
  float* d = nullptr;                        // diagnostics only: the one entry point that allocates
  NERF_CUDA_TRY(cudaMalloc(&d, 4 * sizeof(float)));
  cudaMemset(d, 0, 4 * sizeof(float));
  if (variant == 2) selftest_umma_pair_kernel<<<2, 128, 32768>>>(d);
  else selftest_umma_kernel<<<1, 128, 32768>>>(variant, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(result_host, d, 4 * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return (int)e;
  return (result_host[0] == 0.0f && result_host[1] > 0.0f) ? NERF_OK : NERF_ERR_SELFTEST;
}

#ifdef NERF_DIAGNOSTICS
// ------------------------------------------------------------------------------------ diagnostics
int nerf_debug_trace(void* device_buffer) {
  g_trace.store(static_cast<long long*>(device_buffer));
  return NERF_OK;
}

// diagnostics: UMMA issue-rate microbenchmark (selftest.cuh); cycles_host[0] = cycles, [1] = MMAs
int nerf_debug_umma_rate(int variant, int n_mma, long long* cycles_host) {
  DeviceInfo* dev;
  int rc = device_info(&dev);
  if (rc != NERF_OK) return rc;
  if (variant < 0 || variant > 80 || n_mma < 2 || !cycles_host) return NERF_ERR_ARG;
  long long* d = nullptr;
  NERF_CUDA_TRY(cudaMalloc(&d, 2 * sizeof(long long)));
  cudaMemset(d, 0, 2 * sizeof(long long));
  cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 131072);
  if (variant >= 5) {                                  // 5..9: CTA-pair variants 0..4; 10..25: 5 + loop-ingredient flags
    cudaFuncSetAttribute(umma_pair_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 65536);
    umma_pair_rate_kernel<<<2, 128, 65536 + 65536>>>(variant - 5, n_mma, d);
  } else
  umma_rate_kernel<<<1, 128, 65536 + 131072>>>(variant, n_mma, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(cycles_host, d, 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e == cudaSuccess ? NERF_OK : (int)e;
}

// diagnostics: store-rate microbenchmark (selftest.cuh); out_host[0] = max cycles over CTAs, [1] = bytes per CTA, [2] = elapsed ns
int nerf_debug_store_rate(int warps, int iters, int mode, void* device_buffer, long long buffer_bytes, long long* out_host) {
  DeviceInfo* dev;
  int rc = device_info(&dev);
  if (rc != NERF_OK) return rc;
  if (warps < 1 || warps > 32 || iters < 1 || !device_buffer || !out_host) return NERF_ERR_ARG;
  if ((long long)dev->sms * iters * warps * 512 > buffer_bytes) return NERF_ERR_ARG;
  long long* d = nullptr;
  NERF_CUDA_TRY(cudaMalloc(&d, 2 * sizeof(long long)));
  cudaMemset(d, 0, 2 * sizeof(long long));
  cudaFuncSetAttribute(store_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  store_rate_kernel<<<dev->sms, warps * 32, 200 * 1024>>>(static_cast<uint4*>(device_buffer), iters, mode, d);   // warm-up
  cudaMemset(d, 0, 2 * sizeof(long long));
  cudaEventRecord(e0);
  store_rate_kernel<<<dev->sms, warps * 32, 200 * 1024>>>(static_cast<uint4*>(device_buffer), iters, mode, d);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (e == cudaSuccess) e = cudaMemcpy(out_host, d, 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  out_host[2] = (long long)(ms * 1e6);
  cudaFree(d); cudaEventDestroy(e0); cudaEventDestroy(e1);
  return e == cudaSuccess ? NERF_OK : (int)e;
}

#endif  // NERF_DIAGNOSTICS

}  // extern "C"
