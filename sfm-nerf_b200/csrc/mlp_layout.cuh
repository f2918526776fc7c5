// mlp_layout.cuh -- parameter layout of the live NeRFmodel (NeRFModel.py:102-140) and the
// "layer programs" the fused tcgen05 kernels interpret (host + device).
//
// On-chip activation tiles and every HBM stash block use ONE blocked layout
//     [k_core = channel/8][row 0..127][8 channels]   (16 B per (k_core,row), 2 KB per k_core)
// which is simultaneously
//   * the no-swizzle K-major UMMA canonical layout (LBO = 2048 B, SBO = 128 B) when the channel
//     axis is the contraction axis (forward / dgrad A operands), and
//   * the no-swizzle MN-major canonical layout (LBO = 128 B, SBO = row_count*16 B) when the row
//     (sample) axis is the contraction axis (wgrad operands),
// so tiles move between SMEM and HBM with plain 1-D bulk copies and no re-tiling.
// Packed weights use the same form with rows = output features, one image per K=16 step, split
// into the halves the two CTAs of a pair hold:
//     [k_step][half 0..1][k_core 0..1][n_local 0..N/2-1][8]   (N*32 B per k-step, N*16 B per CTA).
#pragma once
#include <stdint.h>

namespace nerf {

constexpr int kMaxSteps = 12;
constexpr int kNumLinear = 13;          // fc1, fc_layers x4, skip, fc_layers2 x3, density, feature, rgb1, rgb2
constexpr int kKcoreBytes = 2048;       // 128 rows x 16 B
constexpr int kActBytes = 32 * kKcoreBytes;   // 256 channels = 64 KB
constexpr int kEncBytes = 8 * kKcoreBytes;    // 64 channels  = 16 KB
constexpr int kRingSlots = 16;                // slabs in flight (the two tile slots take turns, each streams its own pass of a layer's slabs)
constexpr int kRingSlotBytes = 128 * 32;      // one CTA's half (N/2 = 128 rows) of a K=16 step of an N=256 layer = 4 KB
constexpr int kCommitGroup = 2;               // ring slots are released (one tcgen05.commit) per this many k-steps (the odd slot's barrier)
constexpr int kLoadBatch = 2;                 // slabs a loader / relay warp handles per trip (one per lane): finer trips = shorter supply latency
constexpr int kWaitGroup = 4;                 // k-steps whose full barriers an MMA warp polls back to back before issuing them

// epilogue flags
enum : int32_t {
  EPI_WRITE_DIR = 2,      // fwd: also place the direction encoding into ENC (after the skip layer)
  EPI_DENSITY = 4,        // fwd: density head from fp32 registers (NeRFModel.py:192)
  EPI_RGB = 8,            // fwd: rgb head + sigmoid (NeRFModel.py:199-200), writes sigma_rgb
  EPI_BWD_SIGMA = 16,     // bwd: add d_sigma_pre * w_density before masking
  EPI_LAST = 32           // no further step consumes ACT
};

struct Step {
  int32_t n;            // UMMA N
  int32_t act_ksteps;   // K=16 steps read from ACT (k_cores 0..)
  int32_t enc_ksteps;   // K=16 steps read from ENC
  int32_t w_off;        // byte offset of this step's weight image inside the packed buffer
  int32_t bias_off;     // float offset inside the packed "small" table, -1 = none
  int32_t flags;
  int32_t stash_off;    // byte offset of the OUTPUT block inside a tile's stash (fwd) / dstash (bwd)
  int32_t mask_off;     // byte offset (fwd stash) of the ReLU bit mask: written (fwd) / applied to the output (bwd)
  // pack-time description of the B operand (rows n, contraction index kk):
  int32_t p_w;          // float offset of the Linear weight in the flat parameter vector
  int32_t p_in;         // its in_features (row stride)
  int32_t p_transpose;  // 0: B[n][kk] = W[n][col(kk)];  1: B[n][kk] = W[kk][n]  (dgrad)
  int32_t act_cols;     // kk < 16*act_ksteps maps to source column kk (valid if < act_cols)
  int32_t enc_col0;     // kk >= 16*act_ksteps maps to source column enc_col0 + (kk - 16*act_ksteps)
  int32_t enc_cols;     //   valid if (kk - 16*act_ksteps) < enc_cols
  int32_t p_rows;       // valid rows n
  int32_t p_bias;       // float offset of the bias in the flat vector (-1: none); it rides on the MMA:
  int32_t bias_kstep;   //   1: one extra K=16 step reads ENC k-step 3, whose channel 63 is the constant 1.0
  int32_t bias_in_enc;  //   1: the ENC k-steps of this step already cover channel 63 (PE layers)
};

struct Program {
  int32_t n_steps;
  int32_t image_bytes;     // total bytes of the steps' weight images
  int32_t total_ksteps;    // sum over steps of (act_ksteps + enc_ksteps)
  int32_t pad_;
  Step s[kMaxSteps];
};

// One wgrad work item: dW[m][n] = sum_rows dY[row][m] * X[row][n]
struct WgradItem {
  int32_t dy_off;       // byte offset of the dY block inside a dstash tile
  int32_t m_kcores;     // 32 (M = 256, two UMMA halves) or 16 (M = 128)
  int32_t x_off;        // byte offset of the input block inside a (fwd) stash tile
  int32_t n;            // UMMA N (multiple of 16, <= 256)
  int32_t n_valid;      // columns actually written
  int32_t w_off;        // float offset of the Linear weight in the flat vector
  int32_t in_f;         // its row stride
  int32_t col0;         // first column written
  int32_t b_off;        // float offset of the bias (-1: this item does not own the bias)
  int32_t cost;         // relative cost per tile (for CTA partitioning)
};
constexpr int kMaxWgradItems = 16;

// layout ids: 0 = live, positional encoding (10, 4); 1 = live, pos_encoding=False; 2 / 3 = the same parameter layouts with a
// ReLU density head instead of softplus -- what the reference's shipped (legacy) checkpoints evaluate once their weights
// are remapped onto the live layer table (sfm-nerf_b200/legacy.py; SURVEY A.4, commented class NeRFModel.py:7-99).
constexpr int kNumLayouts = 4;

struct ModelLayout {
  int32_t layout;
  int32_t density_relu;              // 1: density = relu(.) (legacy checkpoints), 0: softplus (NeRFModel.py:192)
  int32_t pos_in, dir_in;            // 63/27 or 3/3
  int32_t pos_L, dir_L;              // 10/4 or 0/0
  int32_t n_params;
  int32_t w_off[kNumLinear], b_off[kNumLinear], out_f[kNumLinear], in_f[kNumLinear];
  // packed buffer: [fwd image][bwd image][small fp32 table]
  Program fwd, bwd;
  int32_t small_off;                 // byte offset of the fp32 table
  int32_t small_floats;
  int32_t packed_bytes;
  // small fp32 table (float offsets): the two heads that run on CUDA cores (layer biases ride on the MMA)
  int32_t sm_wd, sm_bd, sm_w2, sm_b2;
  // per-tile stash (bytes): PE, DE, H0..H9 (h0..h8, feature), R1, then 11 ReLU bit masks
  // ([row][8 x u32] = 4 KB each; bit c of a row = output channel c of that layer was > 0)
  int32_t st_pe, st_de, st_h[10], st_r1, st_mask, stash_tile_bytes;
  // per-tile dstash: dY0..dY9 (64 KB each), dY10 (32 KB), dsr fp32 [128][4]
  int32_t ds_y[11], ds_sr, dstash_tile_bytes;
  int32_t n_wgrad;
  WgradItem wg[kMaxWgradItems];
};

enum { L_FC1 = 0, L_FC0 = 1, L_SKIP = 5, L_FC2_0 = 6, L_DENS = 9, L_FEAT = 10, L_RGB1 = 11, L_RGB2 = 12 };

inline bool build_model_layout(int layout, ModelLayout& m) {
  if (layout < 0 || layout >= kNumLayouts) return false;
  m = ModelLayout{};
  m.layout = layout;
  m.density_relu = layout >= 2 ? 1 : 0;
  const bool pe = (layout & 1) == 0;
  m.pos_L = pe ? 10 : 0; m.dir_L = pe ? 4 : 0;
  m.pos_in = pe ? 63 : 3; m.dir_in = pe ? 27 : 3;
  const int C = 256;
  const int outs[kNumLinear] = {C, C, C, C, C, C, C, C, C, 1, C, 128, 3};
  const int ins[kNumLinear] = {m.pos_in, C, C, C, C, C + m.pos_in, C, C, C, C, C, C + m.dir_in, 128};
  int off = 0;
  for (int i = 0; i < kNumLinear; ++i) {
    m.out_f[i] = outs[i]; m.in_f[i] = ins[i];
    m.w_off[i] = off; off += outs[i] * ins[i];
    m.b_off[i] = off; off += outs[i];
  }
  m.n_params = off;

  const int pos_ks = (m.pos_in + 15) / 16;     // 4 (PE) or 1
  const int dir_ks = (m.dir_in + 15) / 16;     // 2 (PE) or 1

  // stash
  int so = 0;
  m.st_pe = so; so += 8 * kKcoreBytes;
  m.st_de = so; so += 4 * kKcoreBytes;
  for (int i = 0; i < 10; ++i) { m.st_h[i] = so; so += kActBytes; }
  m.st_r1 = so; so += 16 * kKcoreBytes;
  m.st_mask = so; so += 11 * 4096;
  m.stash_tile_bytes = so;
  int dso = 0;
  for (int i = 0; i < 10; ++i) { m.ds_y[i] = dso; dso += kActBytes; }
  m.ds_y[10] = dso; dso += 16 * kKcoreBytes;
  m.ds_sr = dso; dso += 128 * 16;
  m.dstash_tile_bytes = dso;

  auto add = [](Program& p, Step st) {
    st.w_off = p.image_bytes;
    p.image_bytes += (st.act_ksteps + st.enc_ksteps + st.bias_kstep) * st.n * 32;
    p.total_ksteps += st.act_ksteps + st.enc_ksteps + st.bias_kstep;
    p.s[p.n_steps++] = st;
  };
  // ---- forward program ------------------------------------------------------------------
  int bias_slot = 0;
  auto fstep = [&](int lin, int n, int act_ks, int enc_ks, int enc_col0, int enc_cols, int flags,
                   int stash_off) {
    Step st{};
    st.n = n; st.act_ksteps = act_ks; st.enc_ksteps = enc_ks;
    st.bias_off = 256 * bias_slot++;
    st.flags = flags; st.stash_off = stash_off; st.mask_off = -1;
    st.p_w = m.w_off[lin]; st.p_in = m.in_f[lin]; st.p_transpose = 0;
    st.act_cols = act_ks ? 256 : 0; st.enc_col0 = enc_col0; st.enc_cols = enc_cols;
    st.p_rows = m.out_f[lin];
    st.p_bias = m.b_off[lin];
    st.bias_in_enc = (enc_ks == 4 && enc_col0 + 63 >= st.p_in) ? 1 : 0;   // PE tile (64 ch): channel 63 is free
    st.bias_kstep = st.bias_in_enc ? 0 : 1;
    return st;
  };
  Program& f = m.fwd;
  add(f, fstep(L_FC1, 256, 0, pos_ks, 0, m.pos_in, 0, m.st_h[0]));
  for (int i = 0; i < 4; ++i) add(f, fstep(L_FC0 + i, 256, 16, 0, 0, 0, 0, m.st_h[1 + i]));
  add(f, fstep(L_SKIP, 256, 16, pos_ks, 256, m.pos_in, EPI_WRITE_DIR, m.st_h[5]));
  for (int i = 0; i < 3; ++i)
    add(f, fstep(L_FC2_0 + i, 256, 16, 0, 0, 0, (i == 2 ? EPI_DENSITY : 0), m.st_h[6 + i]));
  add(f, fstep(L_FEAT, 256, 16, 0, 0, 0, 0, m.st_h[9]));
  add(f, fstep(L_RGB1, 128, 16, dir_ks, 256, m.dir_in, EPI_RGB | EPI_LAST, m.st_r1));
  for (int i = 0; i < f.n_steps; ++i) f.s[i].mask_off = m.st_mask + i * 4096;

  // ---- backward (dgrad) program: dX = dY . W, masked by relu'(x) ---------------------------
  auto bstep = [&](int lin, int act_ks, int flags, int dstash_off, int mask_off) {
    Step st{};
    st.n = 256; st.act_ksteps = act_ks; st.enc_ksteps = 0;
    st.bias_off = -1; st.flags = flags; st.stash_off = dstash_off; st.mask_off = mask_off;
    st.p_w = m.w_off[lin]; st.p_in = m.in_f[lin]; st.p_transpose = 1;
    st.act_cols = m.out_f[lin]; st.enc_col0 = 0; st.enc_cols = 0;
    st.p_rows = 256;                    // only the first 256 input columns carry gradient
    st.p_bias = -1; st.bias_kstep = 0; st.bias_in_enc = 0;
    return st;
  };
  Program& b = m.bwd;
  // dY10 (rgb1 pre-activation grad, 128 ch) is produced by the prologue.  Step j outputs the
  // masked gradient w.r.t. the pre-activation of forward step (9-j): dY9 (feature) ... dY0 (fc1).
  add(b, bstep(L_RGB1, 8, 0, m.ds_y[9], m.st_mask + 9 * 4096));
  add(b, bstep(L_FEAT, 16, EPI_BWD_SIGMA, m.ds_y[8], m.st_mask + 8 * 4096));
  add(b, bstep(L_FC2_0 + 2, 16, 0, m.ds_y[7], m.st_mask + 7 * 4096));
  add(b, bstep(L_FC2_0 + 1, 16, 0, m.ds_y[6], m.st_mask + 6 * 4096));
  add(b, bstep(L_FC2_0 + 0, 16, 0, m.ds_y[5], m.st_mask + 5 * 4096));
  add(b, bstep(L_SKIP, 16, 0, m.ds_y[4], m.st_mask + 4 * 4096));
  add(b, bstep(L_FC0 + 3, 16, 0, m.ds_y[3], m.st_mask + 3 * 4096));
  add(b, bstep(L_FC0 + 2, 16, 0, m.ds_y[2], m.st_mask + 2 * 4096));
  add(b, bstep(L_FC0 + 1, 16, 0, m.ds_y[1], m.st_mask + 1 * 4096));
  add(b, bstep(L_FC0 + 0, 16, EPI_LAST, m.ds_y[0], m.st_mask + 0 * 4096));

  // packed buffer
  const int fwd_bytes = f.image_bytes, bwd_bytes = b.image_bytes;
  for (int i = 0; i < b.n_steps; ++i) b.s[i].w_off += fwd_bytes;
  m.sm_wd = 0; m.sm_bd = m.sm_wd + 256; m.sm_w2 = m.sm_bd + 4; m.sm_b2 = m.sm_w2 + 384;
  m.small_floats = m.sm_b2 + 4;
  m.small_off = fwd_bytes + bwd_bytes;
  m.packed_bytes = m.small_off + m.small_floats * 4;

  // ---- wgrad work items -------------------------------------------------------------------
  auto wadd = [&](int lin, int dy_off, int m_kcores, int x_off, int n, int n_valid, int col0,
                  bool owns_bias) {
    WgradItem w{};
    w.dy_off = dy_off; w.m_kcores = m_kcores; w.x_off = x_off; w.n = n; w.n_valid = n_valid;
    w.w_off = m.w_off[lin]; w.in_f = m.in_f[lin]; w.col0 = col0;
    w.b_off = owns_bias ? m.b_off[lin] : -1;
    // cost per 128-row tile: the bytes of both operand blocks (the kernel is HBM-bound; k-cores of 2 KB)
    w.cost = m_kcores + n / 8;
    m.wg[m.n_wgrad++] = w;
  };
  const int pos_n = pos_ks * 16, dir_n = dir_ks * 16;
  wadd(L_FC1, m.ds_y[0], 32, m.st_pe, pos_n, m.pos_in, 0, true);
  for (int i = 0; i < 4; ++i) wadd(L_FC0 + i, m.ds_y[1 + i], 32, m.st_h[i], 256, 256, 0, true);
  wadd(L_SKIP, m.ds_y[5], 32, m.st_h[4], 256, 256, 0, true);
  wadd(L_SKIP, m.ds_y[5], 32, m.st_pe, pos_n, m.pos_in, 256, false);
  for (int i = 0; i < 3; ++i) wadd(L_FC2_0 + i, m.ds_y[6 + i], 32, m.st_h[5 + i], 256, 256, 0, true);
  wadd(L_FEAT, m.ds_y[9], 32, m.st_h[8], 256, 256, 0, true);
  wadd(L_RGB1, m.ds_y[10], 16, m.st_h[9], 256, 256, 0, true);
  wadd(L_RGB1, m.ds_y[10], 16, m.st_de, dir_n, m.dir_in, 256, false);
  return true;
}

}  // namespace nerf
