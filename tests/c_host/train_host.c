/* A torch-free C host for libnerf_b200.so: the reference's training loop body (Phase2/train.py:447-468) driven through
 * the C ABI alone (include/nerf_b200.h) -- cudaMalloc'ed buffers, one nerf_train_step call per step, replayed as a CUDA
 * graph from the second step on.  Test infrastructure (tests/test_c_host.py builds and runs it and compares the losses with
 * the Python drop-in's); it is also the example INTEGRATION.md points a C / C++ / cgo / JNI maintainer to.
 *
 *   train_host <scene.bin> <steps>
 * scene.bin: int32 {n_images, width, height, n_rays, n_sample, layout, seed, perturb}; double {fx, fy, cx, cy};
 *            double poses[n_images][16]; float images[n_images][height][width][3]; float params[nerf_param_count(layout)]
 * prints one line per step: "<step> <loss>"
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nerf_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define NK(x) do { int r_ = (x); if (r_ != 0) { fprintf(stderr, "%s: %d (%s)\n", #x, r_, nerf_error_string(r_)); return 3; } } while (0)

static void* dmalloc(size_t bytes, int zero) {
  void* p = NULL;
  if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) { fprintf(stderr, "cudaMalloc(%zu) failed\n", bytes); exit(2); }
  if (zero) cudaMemset(p, 0, bytes);
  return p;
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s scene.bin steps\n", argv[0]); return 1; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 1; }
  const int steps = atoi(argv[2]);
  int32_t h[8];
  double k[4];
  if (fread(h, sizeof(int32_t), 8, f) != 8 || fread(k, sizeof(double), 4, f) != 4) return 1;
  const int n_images = h[0], width = h[1], height = h[2], n_rays = h[3], n_sample = h[4], layout = h[5], seed = h[6], perturb = h[7];
  const int64_t n_params = nerf_param_count(layout);
  if (n_params <= 0) { fprintf(stderr, "bad layout %d\n", layout); return 1; }
  const size_t n_pose = (size_t)n_images * 16, n_img = (size_t)n_images * height * width * 3;
  double* poses_h = (double*)malloc(n_pose * sizeof(double));
  float* images_h = (float*)malloc(n_img * sizeof(float));
  float* params_h = (float*)malloc((size_t)n_params * sizeof(float));
  if (fread(poses_h, sizeof(double), n_pose, f) != n_pose || fread(images_h, sizeof(float), n_img, f) != n_img ||
      fread(params_h, sizeof(float), (size_t)n_params, f) != (size_t)n_params) { fprintf(stderr, "short scene file\n"); return 1; }
  fclose(f);

  NK(nerf_device_check());
  const int64_t rows = (int64_t)n_rays * n_sample;
  double* poses = (double*)dmalloc(n_pose * sizeof(double), 0);
  float* images = (float*)dmalloc(n_img * sizeof(float), 0);
  float* params = (float*)dmalloc((size_t)n_params * 4, 0);
  CK(cudaMemcpy(poses, poses_h, n_pose * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(images, images_h, n_img * sizeof(float), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(params, params_h, (size_t)n_params * 4, cudaMemcpyHostToDevice));

  NerfTrainStep c;
  memset(&c, 0, sizeof(c));
  c.struct_bytes = sizeof(c);
  c.layout = layout; c.n_sample = n_sample; c.perturb = perturb; c.world = 1; c.rank = 0;
  c.n_images = n_images; c.width = width; c.height = height; c.n_rays = n_rays;
  c.fx = k[0]; c.fy = k[1]; c.cx = k[2]; c.cy = k[3];
  c.near = 2.0f; c.far = 6.0f;
  c.seed = (uint64_t)seed; c.z_seed = (uint64_t)seed ^ 0x5bd1e995ull;
  c.lr0 = 5e-4; c.lr_decay = 0.1; c.lr_decay_steps = 50000;                   /* train.py:401-404 */
  c.beta1 = 0.9; c.beta2 = 0.999; c.eps = 1e-8; c.grad_scale = 1.0f;
  c.poses = poses; c.images = images; c.batch9 = NULL;
  c.rays_o = (float*)dmalloc((size_t)n_rays * 12, 0); c.rays_d = (float*)dmalloc((size_t)n_rays * 12, 0);
  c.target = (float*)dmalloc((size_t)n_rays * 12, 0); c.z = (float*)dmalloc((size_t)rows * 4, 0);
  c.sigma_rgb = (float*)dmalloc((size_t)rows * 16, 0); c.d_sigma_rgb = (float*)dmalloc((size_t)rows * 16, 0);
  c.loss = (float*)dmalloc(4, 1);
  c.composite_ws = dmalloc((size_t)nerf_composite_train_workspace_bytes(n_rays) + 4, 1);
  c.stash = dmalloc((size_t)nerf_mlp_stash_bytes(layout, rows), 0);
  c.dstash = dmalloc((size_t)nerf_mlp_dstash_bytes(layout, rows), 0);
  c.packed = dmalloc((size_t)nerf_mlp_packed_bytes(layout), 0);
  c.params = params;
  c.grads = (float*)dmalloc((size_t)n_params * 4, 1);
  c.exp_avg = (float*)dmalloc((size_t)n_params * 4, 1);
  c.exp_avg_sq = (float*)dmalloc((size_t)n_params * 4, 1);
  c.step_state = (unsigned long long*)dmalloc(8, 1);

  cudaStream_t stream;
  CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  NK(nerf_mlp_pack_weights(params, layout, c.packed, stream));

  cudaGraphExec_t graph_exec = NULL;
  float* loss_h = NULL;
  CK(cudaMallocHost((void**)&loss_h, 4));
  for (int step = 0; step < steps; ++step) {
    if (graph_exec) {
      CK(cudaGraphLaunch(graph_exec, stream));
    } else {
      NK(nerf_train_step(&c, NERF_STEP_ALL, stream));                          /* first step: plain launches (lazy init) */
    }
    CK(cudaMemcpyAsync(loss_h, c.loss, 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    printf("%d %.9g\n", step, (double)*loss_h);
    if (step == 0) {                                                           /* then: capture once, replay */
      cudaGraph_t graph;
      CK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
      NK(nerf_train_step(&c, NERF_STEP_ALL, stream));
      CK(cudaStreamEndCapture(stream, &graph));
      CK(cudaGraphInstantiate(&graph_exec, graph, 0));
      CK(cudaGraphDestroy(graph));
    }
  }
  unsigned long long it = 0;
  CK(cudaMemcpy(&it, c.step_state, 8, cudaMemcpyDeviceToHost));
  printf("iterations %llu graph %d\n", it, graph_exec != NULL);
  return it == (unsigned long long)steps ? 0 : 4;
}
