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
constexpr int kSlabBytes = 128 * 32;          // one CTA's half (N/2 = 128 rows) of a K=16 step of an N=256 layer = 4 KB
constexpr int kRingSlotBytes = 2 * kSlabBytes; // a ring slot holds a SLAB PAIR: two consecutive k-steps of one pass (the second half stays
                                              // unused for the last k-step of a pass with an odd number of them)
constexpr int kRingSlots = 8;                 // slab pairs in flight (the two tile slots take turns, each streams its own pass of a layer's slabs)

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

// ---- CTA-pair wgrad (wgrad_pair.cuh): one work item = everything one pair accumulates from the operand blocks it stages ----
enum : int32_t { W2_OUT_WEIGHT = 0, W2_OUT_BIAS = 1, W2_OUT_HEAD = 2 };
struct W2Seg {            // one bulk copy: bytes from (dstash | stash) tile + off[rank] -- each CTA of the pair loads ITS half
  int32_t off[2];
  int32_t bytes;
  int8_t from_dstash[2];
  int8_t pad_[2];
};
struct W2Load { int32_t n_segs; W2Seg seg[3]; };                  // the segments land back to back in one 40 KB ring slot
struct W2Mma {            // D[256 x n] (+)= A[rows x 256]^T . B[rows x n], 8 k-steps of 16 rows per 128-row tile
  int32_t a_load, a_kc;   // A: load index of the tile + k-core offset inside its slot (16 k-cores = this CTA's 128 M rows)
  int32_t b_load, b_kc;   // B: likewise (n/16 k-cores per CTA); b_load < 0: the constant-ones k-core (bias gradients)
  int32_t n, d_col;
};
struct W2Out {            // what a block of TMEM columns is: rows = this CTA's 128 M rows
  int32_t kind, d_col, n_valid, dst, ld, col0;   // WEIGHT: grads[dst + row*ld + col0 + j] += D[row][j], j < n_valid
  int32_t row_base[2];                           // BIAS:   grads[dst + row] += D[row][0]
  int32_t rank_mask;                             // HEAD:   grads[dst + (c-col0)*ld + row] += D[row][c] + D[row][c+8], col0 <= c < col0+n_valid
};
struct W2Item {
  int32_t n_loads, n_mmas, n_outs;
  int32_t depth, stage_bytes, load_off[3];   // ring geometry: `depth` stages of stage_bytes, load l of a tile at load_off[l] inside its stage
  int32_t cot_load, cot_kc;   // where the head cotangents are staged (-1: not in this item)
  int32_t g_bd, g_b2;         // float offsets of the head biases whose gradients come from the staged cotangents (-1: none)
  int32_t cost;               // bytes per tile and CTA (for the split of the pairs over the items)
  W2Load load[3];
  W2Mma mma[5];
  W2Out out[6];
};
constexpr int kMaxW2Items = 10;
constexpr int kW2RingBytes = 221184;          // 216 KB of operand stages per CTA (wgrad_pair.cuh)
constexpr int kW2MaxPairs = 96;
constexpr int kW2MaxUnits = 2;
struct W2Unit { int32_t item, t0, t1; };      // one pair works on tiles [t0, t1) of `item`, then flushes its accumulators

// layout ids: 0 = live, positional encoding (10, 4); 1 = live, pos_encoding=False; 2 / 3 = the same parameter layouts with a
// ReLU density head instead of softplus -- what the reference's shipped (legacy) checkpoints evaluate once their weights
// are remapped onto the live layer table (sfm-nerf_b200/legacy.py; SURVEY A.4, commented class NeRFModel.py:7-99).
constexpr int kNumLayouts = 4;
// Generalised ids (nerf_layout_id): bit 24 set, bits [8,16) = embed_pos_L, [16,24) = embed_direction_L (each 0..10: the encoding
// must fit the 64-channel tile next to the constant-1 channel), bit 0 = pos_encoding off, bit 1 = ReLU density, bit 2 = hidden
// width 128 instead of 256.
constexpr int kLayoutCustomBit = 1 << 24;
constexpr int kLayoutHidden128Bit = 4;
constexpr int kMaxEncodingL = 10;
inline int make_layout_id(int pos_L, int dir_L, int pos_encoding, int density_relu, int hidden) {
  if (pos_L < 0 || dir_L < 0 || pos_L > kMaxEncodingL || dir_L > kMaxEncodingL || (hidden != 256 && hidden != 128)) return -1;
  const int flags = (density_relu ? 2 : 0) | (hidden == 128 ? kLayoutHidden128Bit : 0);
  if (!pos_encoding) return hidden == 256 ? 1 + flags : (kLayoutCustomBit | 1 | flags);
  if (pos_L == 10 && dir_L == 4 && hidden == 256) return flags;
  return kLayoutCustomBit | (pos_L << 8) | (dir_L << 16) | flags;
}

struct ModelLayout {
  int32_t layout;
  int32_t density_relu;              // 1: density = relu(.) (legacy checkpoints), 0: softplus (NeRFModel.py:192)
  int32_t hidden;                    // num_channels: 256 or 128
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
  int32_t ds_y[11], ds_sr, ds_cot, dstash_tile_bytes;     // ds_cot: bf16 head cotangents [hi | lo][128 rows][8] (wgrad_pair.cuh)
  int32_t n_wgrad;
  WgradItem wg[kMaxWgradItems];
  int32_t n_w2;
  W2Item w2[kMaxW2Items];
};

enum { L_FC1 = 0, L_FC0 = 1, L_SKIP = 5, L_FC2_0 = 6, L_DENS = 9, L_FEAT = 10, L_RGB1 = 11, L_RGB2 = 12 };

inline bool build_model_layout(int layout, ModelLayout& m) {
  int pos_L = 10, dir_L = 4;
  if (layout >= 0 && (layout & kLayoutCustomBit)) {
    pos_L = (layout >> 8) & 0xff; dir_L = (layout >> 16) & 0xff;
    if ((layout & ~(kLayoutCustomBit | 0xffff00 | 7)) != 0 || pos_L > kMaxEncodingL || dir_L > kMaxEncodingL) return false;
  } else if (layout < 0 || layout >= kNumLayouts) {
    return false;
  }
  m = ModelLayout{};
  m.layout = layout;
  m.density_relu = (layout & 2) ? 1 : 0;
  const bool pe = (layout & 1) == 0;
  m.pos_L = pe ? pos_L : 0; m.dir_L = pe ? dir_L : 0;
  m.pos_in = 3 * (2 * m.pos_L + 1); m.dir_in = 3 * (2 * m.dir_L + 1);       // NeRFModel.py:112-114 (3 when the encoding is off)
  const int C = (layout >= 0 && (layout & kLayoutCustomBit) && (layout & kLayoutHidden128Bit)) ? 128 : 256;   // num_channels
  m.hidden = C;
  const int block_bytes = (C / 8) * kKcoreBytes;                            // one C-channel activation / gradient block of a tile
  const int c_ks = C / 16;                                                  // K = 16 steps over a C-channel activation tile
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
  m.st_de = so; so += (dir_ks > 2 ? 8 : 4) * kKcoreBytes;                  // direction-encoding tile: 32 channels (L <= 4) or 64
  for (int i = 0; i < 10; ++i) { m.st_h[i] = so; so += block_bytes; }
  m.st_r1 = so; so += 16 * kKcoreBytes;
  m.st_mask = so; so += 11 * 4096;
  m.stash_tile_bytes = so;
  int dso = 0;
  for (int i = 0; i < 10; ++i) { m.ds_y[i] = dso; dso += block_bytes; }
  m.ds_y[10] = dso; dso += 16 * kKcoreBytes;
  m.ds_sr = dso; dso += 128 * 16;
  m.ds_cot = dso; dso += 2 * kKcoreBytes;
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
    st.act_cols = act_ks ? C : 0; st.enc_col0 = enc_col0; st.enc_cols = enc_cols;
    st.p_rows = m.out_f[lin];
    st.p_bias = m.b_off[lin];
    st.bias_in_enc = (enc_ks == 4 && enc_col0 + 63 >= st.p_in) ? 1 : 0;   // PE tile (64 ch): channel 63 is free
    st.bias_kstep = st.bias_in_enc ? 0 : 1;
    return st;
  };
  Program& f = m.fwd;
  add(f, fstep(L_FC1, C, 0, pos_ks, 0, m.pos_in, 0, m.st_h[0]));
  for (int i = 0; i < 4; ++i) add(f, fstep(L_FC0 + i, C, c_ks, 0, 0, 0, 0, m.st_h[1 + i]));
  add(f, fstep(L_SKIP, C, c_ks, pos_ks, C, m.pos_in, EPI_WRITE_DIR, m.st_h[5]));
  for (int i = 0; i < 3; ++i)
    add(f, fstep(L_FC2_0 + i, C, c_ks, 0, 0, 0, (i == 2 ? EPI_DENSITY : 0), m.st_h[6 + i]));
  add(f, fstep(L_FEAT, C, c_ks, 0, 0, 0, 0, m.st_h[9]));
  add(f, fstep(L_RGB1, 128, c_ks, dir_ks, C, m.dir_in, EPI_RGB | EPI_LAST, m.st_r1));
  for (int i = 0; i < f.n_steps; ++i) f.s[i].mask_off = m.st_mask + i * 4096;

  // ---- backward (dgrad) program: dX = dY . W, masked by relu'(x) ---------------------------
  auto bstep = [&](int lin, int act_ks, int flags, int dstash_off, int mask_off) {
    Step st{};
    st.n = C; st.act_ksteps = act_ks; st.enc_ksteps = 0;
    st.bias_off = -1; st.flags = flags; st.stash_off = dstash_off; st.mask_off = mask_off;
    st.p_w = m.w_off[lin]; st.p_in = m.in_f[lin]; st.p_transpose = 1;
    st.act_cols = m.out_f[lin]; st.enc_col0 = 0; st.enc_cols = 0;
    st.p_rows = C;                      // only the first C input columns (the activations, not the encodings) carry gradient
    st.p_bias = -1; st.bias_kstep = 0; st.bias_in_enc = 0;
    return st;
  };
  Program& b = m.bwd;
  // dY10 (rgb1 pre-activation grad, 128 ch) is produced by the prologue.  Step j outputs the
  // masked gradient w.r.t. the pre-activation of forward step (9-j): dY9 (feature) ... dY0 (fc1).
  add(b, bstep(L_RGB1, 8, 0, m.ds_y[9], m.st_mask + 9 * 4096));
  add(b, bstep(L_FEAT, c_ks, EPI_BWD_SIGMA, m.ds_y[8], m.st_mask + 8 * 4096));
  add(b, bstep(L_FC2_0 + 2, c_ks, 0, m.ds_y[7], m.st_mask + 7 * 4096));
  add(b, bstep(L_FC2_0 + 1, c_ks, 0, m.ds_y[6], m.st_mask + 6 * 4096));
  add(b, bstep(L_FC2_0 + 0, c_ks, 0, m.ds_y[5], m.st_mask + 5 * 4096));
  add(b, bstep(L_SKIP, c_ks, 0, m.ds_y[4], m.st_mask + 4 * 4096));
  add(b, bstep(L_FC0 + 3, c_ks, 0, m.ds_y[3], m.st_mask + 3 * 4096));
  add(b, bstep(L_FC0 + 2, c_ks, 0, m.ds_y[2], m.st_mask + 2 * 4096));
  add(b, bstep(L_FC0 + 1, c_ks, 0, m.ds_y[1], m.st_mask + 1 * 4096));
  add(b, bstep(L_FC0 + 0, c_ks, EPI_LAST, m.ds_y[0], m.st_mask + 0 * 4096));

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
  if (C == 256) {                                      // the round-1 single-CTA kernel (wgrad.cuh) is built for 256 channels only
  wadd(L_FC1, m.ds_y[0], 32, m.st_pe, pos_n, m.pos_in, 0, true);
  for (int i = 0; i < 4; ++i) wadd(L_FC0 + i, m.ds_y[1 + i], 32, m.st_h[i], 256, 256, 0, true);
  wadd(L_SKIP, m.ds_y[5], 32, m.st_h[4], 256, 256, 0, true);
  wadd(L_SKIP, m.ds_y[5], 32, m.st_pe, pos_n, m.pos_in, 256, false);
  for (int i = 0; i < 3; ++i) wadd(L_FC2_0 + i, m.ds_y[6 + i], 32, m.st_h[5 + i], 256, 256, 0, true);
  wadd(L_FEAT, m.ds_y[9], 32, m.st_h[8], 256, 256, 0, true);
  wadd(L_RGB1, m.ds_y[10], 16, m.st_h[9], 256, 256, 0, true);
  wadd(L_RGB1, m.ds_y[10], 16, m.st_de, dir_n, m.dir_in, 256, false);
  }

  // ---- CTA-pair wgrad items ---------------------------------------------------------------
  const int H = 16 * kKcoreBytes;                     // half of a 256-channel block: this CTA's 16 k-cores
  auto seg = [](int off0, int off1, int bytes, bool ds0, bool ds1) {
    W2Seg sg{}; sg.off[0] = off0; sg.off[1] = off1; sg.bytes = bytes; sg.from_dstash[0] = ds0; sg.from_dstash[1] = ds1; return sg;
  };
  auto half_seg = [&](int off, int bytes_total, bool ds) { return seg(off, off + bytes_total / 2, bytes_total / 2, ds, ds); };
  auto mma = [](int a_load, int a_kc, int b_load, int b_kc, int n, int d_col) { W2Mma q{a_load, a_kc, b_load, b_kc, n, d_col}; return q; };
  auto out = [](int kind, int d_col, int n_valid, int dst, int ld, int col0, int rb0, int rb1, int mask) {
    W2Out o{}; o.kind = kind; o.d_col = d_col; o.n_valid = n_valid; o.dst = dst; o.ld = ld; o.col0 = col0;
    o.row_base[0] = rb0; o.row_base[1] = rb1; o.rank_mask = mask; return o;
  };
  auto finish = [&](W2Item& it) {
    it.cost = 0;
    for (int l = 0; l < it.n_loads; ++l) {
      it.load_off[l] = it.cost;
      for (int g = 0; g < it.load[l].n_segs; ++g) it.cost += it.load[l].seg[g].bytes;
    }
    it.stage_bytes = it.cost;
    it.depth = kW2RingBytes / it.stage_bytes;
    if (it.depth > 3) it.depth = 3;
    m.w2[m.n_w2++] = it;
  };
  if (C == 256) {
  auto plain = [&](int lin, int dy_off, int x_off) {               // dY_l x h_(l-1), 256 x 256
    W2Item it{}; it.cot_load = -1; it.cot_kc = 0; it.g_bd = -1; it.g_b2 = -1;
    it.n_loads = 2;
    it.load[0].n_segs = 1; it.load[0].seg[0] = half_seg(dy_off, 2 * H, true);
    it.load[1].n_segs = 1; it.load[1].seg[0] = half_seg(x_off, 2 * H, false);
    it.n_mmas = 2; it.mma[0] = mma(0, 0, 1, 0, 256, 0); it.mma[1] = mma(0, 0, -1, 0, 16, 256);
    it.n_outs = 2;
    it.out[0] = out(W2_OUT_WEIGHT, 0, 256, m.w_off[lin], m.in_f[lin], 0, 0, 128, 3);
    it.out[1] = out(W2_OUT_BIAS, 256, 1, m.b_off[lin], 0, 0, 0, 128, 3);
    finish(it);
  };
  {   // skip_layer: dY5 x [h4 | PE], and fc1: dY0 x PE from the same staged PE half
    W2Item it{}; it.cot_load = -1; it.g_bd = -1; it.g_b2 = -1;
    const int pe_bytes = (pos_n / 8) * kKcoreBytes;                // PE channels used by the MMAs (64 or 16), split over the pair
    it.n_loads = 3;
    it.load[0].n_segs = 1; it.load[0].seg[0] = half_seg(m.ds_y[5], 2 * H, true);
    it.load[1].n_segs = 1; it.load[1].seg[0] = half_seg(m.ds_y[0], 2 * H, true);
    it.load[2].n_segs = 2; it.load[2].seg[0] = half_seg(m.st_h[4], 2 * H, false); it.load[2].seg[1] = half_seg(m.st_pe, pe_bytes, false);
    it.n_mmas = 5;
    it.mma[0] = mma(0, 0, 2, 0, 256, 0);            // dY5 x h4
    it.mma[1] = mma(0, 0, 2, 16, pos_n, 256);       // dY5 x PE
    it.mma[2] = mma(1, 0, 2, 16, pos_n, 320);       // dY0 x PE
    it.mma[3] = mma(0, 0, -1, 0, 16, 384);          // bias of skip_layer
    it.mma[4] = mma(1, 0, -1, 0, 16, 400);          // bias of fc1
    it.n_outs = 5;
    it.out[0] = out(W2_OUT_WEIGHT, 0, 256, m.w_off[L_SKIP], m.in_f[L_SKIP], 0, 0, 128, 3);
    it.out[1] = out(W2_OUT_WEIGHT, 256, m.pos_in, m.w_off[L_SKIP], m.in_f[L_SKIP], 256, 0, 128, 3);
    it.out[2] = out(W2_OUT_WEIGHT, 320, m.pos_in, m.w_off[L_FC1], m.in_f[L_FC1], 0, 0, 128, 3);
    it.out[3] = out(W2_OUT_BIAS, 384, 1, m.b_off[L_SKIP], 0, 0, 0, 128, 3);
    it.out[4] = out(W2_OUT_BIAS, 400, 1, m.b_off[L_FC1], 0, 0, 0, 128, 3);
    finish(it);
  }
  for (int i = 0; i < 4; ++i) plain(L_FC0 + i, m.ds_y[1 + i], m.st_h[i]);
  for (int i = 0; i < 3; ++i) plain(L_FC2_0 + i, m.ds_y[6 + i], m.st_h[5 + i]);
  {   // feature_layer: dY9 x h8, and density_layer's head: h8 (as the M side) x cotangents
    W2Item it{}; it.g_b2 = -1;
    it.n_loads = 2;
    it.load[0].n_segs = 1; it.load[0].seg[0] = half_seg(m.ds_y[9], 2 * H, true);
    it.load[1].n_segs = 2; it.load[1].seg[0] = half_seg(m.st_h[8], 2 * H, false); it.load[1].seg[1] = half_seg(m.ds_cot, 2 * kKcoreBytes, true);
    it.cot_load = 1; it.cot_kc = 16; it.g_bd = m.b_off[L_DENS];
    it.n_mmas = 3;
    it.mma[0] = mma(0, 0, 1, 0, 256, 0);
    it.mma[1] = mma(1, 0, 1, 16, 16, 256);          // h8 x cot
    it.mma[2] = mma(0, 0, -1, 0, 16, 272);
    it.n_outs = 3;
    it.out[0] = out(W2_OUT_WEIGHT, 0, 256, m.w_off[L_FEAT], m.in_f[L_FEAT], 0, 0, 128, 3);
    it.out[1] = out(W2_OUT_HEAD, 256, 1, m.w_off[L_DENS], 0, 0, 0, 128, 3);
    it.out[2] = out(W2_OUT_BIAS, 272, 1, m.b_off[L_FEAT], 0, 0, 0, 128, 3);
    finish(it);
  }
  {   // rgb_layer1: leader rows = dY10 (128 ch) x [h9 | DE]; peer rows = r1 (128 ch) x cotangents = rgb_layer2's head
    W2Item it{}; it.g_bd = -1;
    const int de_bytes = (dir_n / 8) * kKcoreBytes;
    it.n_loads = 2;
    it.load[0].n_segs = 1; it.load[0].seg[0] = seg(m.ds_y[10], m.st_r1, H, true, false);
    it.load[1].n_segs = 3; it.load[1].seg[0] = half_seg(m.st_h[9], 2 * H, false); it.load[1].seg[1] = half_seg(m.st_de, de_bytes, false);
    it.load[1].seg[2] = half_seg(m.ds_cot, 2 * kKcoreBytes, true);
    const int de_kc = dir_n / 16;                                  // DE k-cores per CTA
    it.cot_load = 1; it.cot_kc = 16 + de_kc; it.g_b2 = m.b_off[L_RGB2];
    it.n_mmas = 4;
    it.mma[0] = mma(0, 0, 1, 0, 256, 0);
    it.mma[1] = mma(0, 0, 1, 16, dir_n, 256);
    const int d_cot = 256 + dir_n, d_bias = d_cot + 16;            // TMEM columns behind the direction block
    it.mma[2] = mma(0, 0, 1, 16 + de_kc, 16, d_cot);
    it.mma[3] = mma(0, 0, -1, 0, 16, d_bias);
    it.n_outs = 4;
    it.out[0] = out(W2_OUT_WEIGHT, 0, 256, m.w_off[L_RGB1], m.in_f[L_RGB1], 0, 0, 0, 1);
    it.out[1] = out(W2_OUT_WEIGHT, 256, m.dir_in, m.w_off[L_RGB1], m.in_f[L_RGB1], 256, 0, 0, 1);
    it.out[2] = out(W2_OUT_HEAD, d_cot, 3, m.w_off[L_RGB2], 128, 1, 0, 0, 2);
    it.out[3] = out(W2_OUT_BIAS, d_bias, 1, m.b_off[L_RGB1], 0, 0, 0, 0, 1);
    finish(it);
  }
  } else {
    // ---- hidden width 128: a layer's dY has 128 channels = ONE CTA's rows of the M = 256 pair, so every item carries TWO layers:
    // the leader's rows are dY of layer a, the peer's rows dY of layer b, and B = [X_a | X_b] (each CTA loads the block of its own
    // layer).  Of the four blocks of D the two diagonal ones are the gradients; the off-diagonal ones are never read out.
    const int Bk = 16 * kKcoreBytes;                   // a 128-channel block of a tile (16 k-cores)
    auto two = [&](int lin_a, int dy_a, int x_a, int lin_b, int dy_b, int x_b, bool dens_head_on_b) {
      W2Item it{}; it.cot_load = -1; it.cot_kc = 0; it.g_bd = -1; it.g_b2 = -1;
      it.n_loads = 2;
      it.load[0].n_segs = 1; it.load[0].seg[0] = seg(dy_a, dy_b, Bk, true, true);
      it.load[1].n_segs = 1; it.load[1].seg[0] = seg(x_a, x_b, Bk, false, false);
      it.n_mmas = 2; it.mma[0] = mma(0, 0, 1, 0, 256, 0); it.mma[1] = mma(0, 0, -1, 0, 16, 256);
      const int d_head = 272;
      if (dens_head_on_b) {                            // X_b = h8 on the peer: h8 (as the M side) x cotangents = density_layer's gradient
        it.load[1].n_segs = 2; it.load[1].seg[1] = half_seg(m.ds_cot, 2 * kKcoreBytes, true);
        it.cot_load = 1; it.cot_kc = 16; it.g_bd = m.b_off[L_DENS];
        it.mma[2] = mma(1, 0, 1, 16, 16, d_head); it.n_mmas = 3;
      }
      it.n_outs = 0;
      it.out[it.n_outs++] = out(W2_OUT_WEIGHT, 0, 128, m.w_off[lin_a], m.in_f[lin_a], 0, 0, 0, 1);
      it.out[it.n_outs++] = out(W2_OUT_WEIGHT, 128, 128, m.w_off[lin_b], m.in_f[lin_b], 0, 0, 0, 2);
      it.out[it.n_outs++] = out(W2_OUT_BIAS, 256, 1, m.b_off[lin_a], 0, 0, 0, 0, 1);
      it.out[it.n_outs++] = out(W2_OUT_BIAS, 256, 1, m.b_off[lin_b], 0, 0, 0, 0, 2);
      if (dens_head_on_b) it.out[it.n_outs++] = out(W2_OUT_HEAD, d_head, 1, m.w_off[L_DENS], 0, 0, 0, 0, 2);
      finish(it);
    };
    {   // leader: skip_layer (dY5), peer: fc1 (dY0); B = [h4 | PE], each CTA half of both parts
      W2Item it{}; it.cot_load = -1; it.g_bd = -1; it.g_b2 = -1;
      const int pe_bytes = (pos_n / 8) * kKcoreBytes;
      it.n_loads = 2;
      it.load[0].n_segs = 1; it.load[0].seg[0] = seg(m.ds_y[5], m.ds_y[0], Bk, true, true);
      it.load[1].n_segs = 2; it.load[1].seg[0] = half_seg(m.st_h[4], Bk, false); it.load[1].seg[1] = half_seg(m.st_pe, pe_bytes, false);
      it.n_mmas = 3;
      it.mma[0] = mma(0, 0, 1, 0, 128, 0);            // [dY5 ; dY0] x h4   (the peer's block is not read out)
      it.mma[1] = mma(0, 0, 1, 8, pos_n, 128);        // [dY5 ; dY0] x PE
      it.mma[2] = mma(0, 0, -1, 0, 16, 192);
      it.n_outs = 5;
      it.out[0] = out(W2_OUT_WEIGHT, 0, 128, m.w_off[L_SKIP], m.in_f[L_SKIP], 0, 0, 0, 1);
      it.out[1] = out(W2_OUT_WEIGHT, 128, m.pos_in, m.w_off[L_SKIP], m.in_f[L_SKIP], 128, 0, 0, 1);
      it.out[2] = out(W2_OUT_WEIGHT, 128, m.pos_in, m.w_off[L_FC1], m.in_f[L_FC1], 0, 0, 0, 2);
      it.out[3] = out(W2_OUT_BIAS, 192, 1, m.b_off[L_SKIP], 0, 0, 0, 0, 1);
      it.out[4] = out(W2_OUT_BIAS, 192, 1, m.b_off[L_FC1], 0, 0, 0, 0, 2);
      finish(it);
    }
    two(L_FC0 + 0, m.ds_y[1], m.st_h[0], L_FC0 + 1, m.ds_y[2], m.st_h[1], false);
    two(L_FC0 + 2, m.ds_y[3], m.st_h[2], L_FC0 + 3, m.ds_y[4], m.st_h[3], false);
    two(L_FC2_0 + 0, m.ds_y[6], m.st_h[5], L_FC2_0 + 1, m.ds_y[7], m.st_h[6], false);
    two(L_FC2_0 + 2, m.ds_y[8], m.st_h[7], L_FEAT, m.ds_y[9], m.st_h[8], true);
    {   // leader: rgb_layer1 (dY10), peer: r1 rows for rgb_layer2's head; B = [h9 | DE | cot], each CTA half of every part
      W2Item it{}; it.g_bd = -1;
      const int de_bytes = (dir_n / 8) * kKcoreBytes;
      it.n_loads = 2;
      it.load[0].n_segs = 1; it.load[0].seg[0] = seg(m.ds_y[10], m.st_r1, Bk, true, false);
      it.load[1].n_segs = 3; it.load[1].seg[0] = half_seg(m.st_h[9], Bk, false); it.load[1].seg[1] = half_seg(m.st_de, de_bytes, false);
      it.load[1].seg[2] = half_seg(m.ds_cot, 2 * kKcoreBytes, true);
      const int de_kc = dir_n / 16;
      it.cot_load = 1; it.cot_kc = 8 + de_kc; it.g_b2 = m.b_off[L_RGB2];
      const int d_de = 128, d_cot = 128 + dir_n, d_bias = d_cot + 16;
      it.n_mmas = 4;
      it.mma[0] = mma(0, 0, 1, 0, 128, 0);
      it.mma[1] = mma(0, 0, 1, 8, dir_n, d_de);
      it.mma[2] = mma(0, 0, 1, 8 + de_kc, 16, d_cot);
      it.mma[3] = mma(0, 0, -1, 0, 16, d_bias);
      it.n_outs = 4;
      it.out[0] = out(W2_OUT_WEIGHT, 0, 128, m.w_off[L_RGB1], m.in_f[L_RGB1], 0, 0, 0, 1);
      it.out[1] = out(W2_OUT_WEIGHT, d_de, m.dir_in, m.w_off[L_RGB1], m.in_f[L_RGB1], 128, 0, 0, 1);
      it.out[2] = out(W2_OUT_HEAD, d_cot, 3, m.w_off[L_RGB2], 128, 1, 0, 0, 2);
      it.out[3] = out(W2_OUT_BIAS, d_bias, 1, m.b_off[L_RGB1], 0, 0, 0, 0, 1);
      finish(it);
    }
  }
  return true;
}

// Split of the pair-wgrad work over `n_pairs` CTA pairs.  All (item, tile) work is laid on one line in item order, weighted by
// the item's bytes per tile, and cut into n_pairs equal pieces; piece boundaries are rounded to whole tiles.  A pair gets a tile
// range of one item, or the tail of one item and the head of the next (<= kW2MaxUnits units).  Returns the number of pairs used
// (tiny inputs: one pair per item), or -1 if a piece would need more units than that.
inline int w2_partition(const ModelLayout& m, int n_tiles, int n_pairs, int32_t* n_units, W2Unit (*unit)[kW2MaxUnits]) {
  if (n_pairs > kW2MaxPairs) n_pairs = kW2MaxPairs;
  if (n_tiles < 16 || n_pairs < m.n_w2) {
    for (int i = 0; i < m.n_w2; ++i) { n_units[i] = 1; unit[i][0] = W2Unit{i, 0, n_tiles}; }
    return m.n_w2;
  }
  double total = 0.0;
  for (int i = 0; i < m.n_w2; ++i) total += (double)m.w2[i].cost * n_tiles;
  int prev_item = 0, prev_tile = 0;
  for (int p = 0; p < n_pairs; ++p) {
    // position of the END of piece p on the line -> (item, tile)
    int item = m.n_w2, tile = 0;
    if (p + 1 < n_pairs) {
      const double end = total * (p + 1) / n_pairs;
      double start = 0.0;
      for (item = 0; item < m.n_w2; ++item) {
        const double span = (double)m.w2[item].cost * n_tiles;
        if (end < start + span) break;
        start += span;
      }
      if (item < m.n_w2) {
        tile = (int)((end - start) / (double)m.w2[item].cost + 0.5);
        if (tile >= n_tiles) { ++item; tile = 0; }
      }
    }
    if (item < prev_item || (item == prev_item && tile < prev_tile)) { item = prev_item; tile = prev_tile; }   // monotone
    int nu = 0;
    int it = prev_item, t = prev_tile;
    while (it < item || (it == item && t < tile)) {
      const int t_end = (it == item) ? tile : n_tiles;
      if (t_end > t) {
        if (nu == kW2MaxUnits) return -1;
        unit[p][nu++] = W2Unit{it, t, t_end};
      }
      if (it == item) break;
      ++it; t = 0;
    }
    n_units[p] = nu;
    prev_item = item; prev_tile = tile;
  }
  return n_pairs;
}

}  // namespace nerf
