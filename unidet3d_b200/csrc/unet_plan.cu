// Stage plan of the sparse-conv U-Net: the whole eval-mode backbone (reference unidet3d/spconv_unet.py:117-240 -- the
// recursion of residual blocks, strided conv, sub-U-Net, inverse conv, concat and tail blocks) issued by ONE C call.
//
// Host code only: it walks the plan and calls ud3d_gemm_fwd for each of the 4 + 11 (L - 1) ... convolutions with the
// same fusion as the Python executor (unidet3d_b200/spconv_unet.py:_forward_level): operand-form feature maps, the
// consumer's folded BatchNorm+ReLU applied by the producer's epilogue, residual adds in the epilogue, the
// [identity | decoder] concat never materialised (two column halves of one buffer).  What it removes is the host cost
// of the ~50 launches (argument marshalling, output allocation: ~23 us each through Python + ctypes, 1.3 ms per batch --
// more than half of the GPU time of the backbone): with the plan the host issues the backbone in ~0.15 ms and stays
// ahead of the GPU, so two batches in flight are GPU-bound.
//
// Buffers: a fixed set per level inside the caller's workspace (ping-pong pairs for the block outputs), so consecutive
// convs reuse the same L2-resident lines.
#include "common.cuh"

using namespace ud3d;

namespace {

struct Bump {
  uint8_t* base;
  size_t top = 0;
  float* take(size_t rows, size_t cols) {
    float* p = (float*)(base + top);
    top += (rows * cols * 4 + 255) & ~(size_t)255;
    return p;
  }
};

struct LevelBufs {
  float *cat_raw, *cat_act, *y_act, *raw[2], *act[2], *a_down, *r, *d_raw, *d_act, *u_act;
};

// same order in the size query and in the run
LevelBufs carve(Bump& b, const ud3d_unet_plan* plan, const ud3d_unet_tables* lv, int l) {
  LevelBufs f = {};
  const size_t n = (size_t)lv[l].n, c = (size_t)plan->level[l].c;
  const bool has_sub = l + 1 < plan->n_levels;
  f.y_act = b.take(n, c);
  f.raw[0] = b.take(n, c); f.raw[1] = b.take(n, c);
  f.act[0] = b.take(n, c); f.act[1] = b.take(n, c);
  if (has_sub) {
    const size_t n1 = (size_t)lv[l + 1].n, c1 = (size_t)plan->level[l + 1].c;
    f.cat_raw = b.take(n, 2 * c); f.cat_act = b.take(n, 2 * c);
    f.a_down = b.take(n, c); f.r = b.take(n, c);
    f.d_raw = b.take(n1, c1); f.d_act = b.take(n1, c1); f.u_act = b.take(n1, c1);
  }
  return f;
}

struct Act {            // an operand-form output of a conv: relu(result * scale + shift), split
  float* buf; int ld; const float* scale; const float* shift;
};

struct Runner {
  const ud3d_unet_plan* plan;
  const ud3d_unet_tables* lv;
  float* const* level_out;
  void* stream;
  Bump bump;

  // out[o,:] = sum_k in_act[table[k][o],:] W_k (+ residual), operand-form input
  int conv(const float* in_act, int ld_in, int c_in, const void* w, int K, int c_out, const int32_t* table, const uint32_t* mask,
           const int32_t* perm, int n_out, const float* residual, int ld_res, float* raw, int ld_raw, bool want_raw, const Act* acts,
           int n_acts) {
    ud3d_gemm_args a = {};
    a.in = in_act; a.ld_in = ld_in; a.c_in = c_in;
    a.table = table; a.tile_mask = mask; a.K = K; a.n_out = n_out;
    a.w_packed = w;
    a.out = raw; a.ld_out = ld_raw; a.c_out = c_out;
    a.residual = residual; a.ld_res = ld_res;
    a.in_split = 1; a.no_raw = want_raw ? 0 : 1;
    for (int i = 0; i < n_acts; ++i) {
      a.out_act[i] = acts[i].buf; a.ld_act[i] = acts[i].ld; a.act_scale[i] = acts[i].scale; a.act_shift[i] = acts[i].shift;
    }
    a.row_perm = perm;
    return ud3d_gemm_fwd(&a, stream);
  }

  // ResidualBlock with equal channel counts (spconv_unet.py:74-91): x + SubM3(BN.SubM3(BN.x)); x_act = operand form of
  // relu(bn0(x))
  int block(const ud3d_unet_block& bp, int l, const float* x_raw, int ld_x, const float* x_act, int ld_xa, float* y_act, float* out_raw,
            int ld_out, bool want_raw, const Act* acts, int n_acts) {
    const ud3d_unet_tables& t = lv[l];
    const int c = plan->level[l].c;
    Act mid = {y_act, c, bp.bn1_scale, bp.bn1_shift};
    int rc = conv(x_act, ld_xa, c, bp.w0, 27, c, t.subm, t.subm_mask, t.row_perm, t.n, nullptr, 0, out_raw, ld_out, false, &mid, 1);
    if (rc) return rc;
    return conv(y_act, c, c, bp.w1, 27, c, t.subm, t.subm_mask, t.row_perm, t.n, x_raw, ld_x, out_raw, ld_out, want_raw, acts, n_acts);
  }

  // x_raw / x_act: the level's input (fp32 and operand form under blocks[0].bn0).  out_raw: the level's fp32 result (may
  // be NULL below the top level: then it is not stored).  out_act (+ out_scale / out_shift = the parent's deconv
  // BatchNorm): operand-form result for the parent's inverse conv, NULL at the top level.
  int level(int l, const float* x_raw, const float* x_act, float* out_raw, float* out_act, const float* out_scale, const float* out_shift) {
    const ud3d_unet_level& P = plan->level[l];
    const ud3d_unet_tables& t = lv[l];
    const int c = P.c, reps = plan->block_reps, n = t.n;
    const bool has_sub = l + 1 < plan->n_levels;
    LevelBufs f = carve(bump, plan, lv, l);
    Act fin = {out_act, c, out_scale, out_shift};
    const int n_fin = out_act ? 1 : 0;
    const bool want_final_raw = out_raw != nullptr;
    int rc;
    const float* raw = x_raw;
    int ld_raw = c;
    const float* act = x_act;
    int ld_act = c;
    int pp = 0;
    if (!has_sub) {
      for (int i = 0; i < reps; ++i) {
        const bool last = i == reps - 1;
        Act nxt = {f.act[pp], c, last ? nullptr : P.blocks[i + 1].bn0_scale, last ? nullptr : P.blocks[i + 1].bn0_shift};
        float* o = (last && out_raw) ? out_raw : f.raw[pp];
        rc = block(P.blocks[i], l, raw, ld_raw, act, ld_act, f.y_act, o, c, last ? want_final_raw : true, last ? &fin : &nxt,
                   last ? n_fin : 1);
        if (rc) return rc;
        raw = o; act = nxt.buf; pp ^= 1;
      }
      return UD3D_OK;
    }
    const ud3d_unet_level& S = plan->level[l + 1];
    const ud3d_unet_tables& t1 = lv[l + 1];
    const int c1 = S.c;
    const float* t_sc = P.tail[0].bn0_scale;
    const float* t_sh = P.tail[0].bn0_shift;        // [2c]: BatchNorm of the concatenated map
    for (int i = 0; i < reps; ++i) {
      const bool last = i == reps - 1;
      if (last) {
        // the encoder-side output: fp32 into the identity half of the concat buffer, operand form twice (under the
        // strided conv's BatchNorm and under the tail block's BatchNorm, first c channels)
        Act two[2] = {{f.a_down, c, P.down_scale, P.down_shift}, {f.cat_act, 2 * c, t_sc, t_sh}};
        rc = block(P.blocks[i], l, raw, ld_raw, act, ld_act, f.y_act, f.cat_raw, 2 * c, true, two, 2);
      } else {
        Act nxt = {f.act[pp], c, P.blocks[i + 1].bn0_scale, P.blocks[i + 1].bn0_shift};
        rc = block(P.blocks[i], l, raw, ld_raw, act, ld_act, f.y_act, f.raw[pp], c, true, &nxt, 1);
        raw = f.raw[pp]; act = f.act[pp]; pp ^= 1;
      }
      if (rc) return rc;
    }
    // SparseConv3d(k=2, s=2) -> sub-U-Net -> SparseInverseConv3d(k=2) into the decoder half of the concat buffer
    Act sub_in = {f.d_act, c1, S.blocks[0].bn0_scale, S.blocks[0].bn0_shift};
    rc = conv(f.a_down, c, c, P.down_w, 8, c1, t.child, t.child_mask, nullptr, t1.n, nullptr, 0, f.d_raw, c1, true, &sub_in, 1);
    if (rc) return rc;
    rc = level(l + 1, f.d_raw, f.d_act, level_out ? level_out[l + 1] : nullptr, f.u_act, P.up_scale, P.up_shift);
    if (rc) return rc;
    Act dec = {f.cat_act + c, 2 * c, t_sc + c, t_sh + c};
    rc = conv(f.u_act, c1, c1, P.up_w, 8, c, t.up, t.up_mask, nullptr, n, nullptr, 0, f.cat_raw + c, 2 * c, true, &dec, 1);
    if (rc) return rc;
    // tail block 0 (2c -> c): SubM1(cat) + SubM3(BN.SubM3(BN.cat))   (spconv_unet.py:36-38)
    {
      ud3d_gemm_args a = {};
      a.in = f.cat_raw; a.ld_in = 2 * c; a.c_in = 2 * c; a.K = 1; a.n_out = n;
      a.w_packed = P.tail[0].wi; a.out = f.r; a.ld_out = c; a.c_out = c;
      rc = ud3d_gemm_fwd(&a, stream);
      if (rc) return rc;
    }
    Act mid = {f.y_act, c, P.tail[0].bn1_scale, P.tail[0].bn1_shift};
    rc = conv(f.cat_act, 2 * c, 2 * c, P.tail[0].w0, 27, c, t.subm, t.subm_mask, t.row_perm, n, nullptr, 0, f.raw[0], c, false, &mid, 1);
    if (rc) return rc;
    {
      const bool last = reps == 1;
      Act nxt = {f.act[0], c, last ? nullptr : P.tail[1].bn0_scale, last ? nullptr : P.tail[1].bn0_shift};
      float* o = (last && out_raw) ? out_raw : f.raw[0];
      rc = conv(f.y_act, c, c, P.tail[0].w1, 27, c, t.subm, t.subm_mask, t.row_perm, n, f.r, c, o, c, last ? want_final_raw : true,
                last ? &fin : &nxt, last ? n_fin : 1);
      if (rc) return rc;
      raw = o; act = f.act[0]; pp = 1;
    }
    for (int i = 1; i < reps; ++i) {
      const bool last = i == reps - 1;
      Act nxt = {f.act[pp], c, last ? nullptr : P.tail[i + 1].bn0_scale, last ? nullptr : P.tail[i + 1].bn0_shift};
      float* o = (last && out_raw) ? out_raw : f.raw[pp];
      rc = block(P.tail[i], l, raw, c, act, c, f.y_act, o, c, last ? want_final_raw : true, last ? &fin : &nxt, last ? n_fin : 1);
      if (rc) return rc;
      raw = o; act = nxt.buf; pp ^= 1;
    }
    return UD3D_OK;
  }
};

int check_plan(const ud3d_unet_plan* plan, const ud3d_unet_tables* lv, const char* who) {
  UD3D_CHECK_ARG(plan && lv, "%s: NULL argument", who);
  UD3D_CHECK_ARG(plan->n_levels >= 1 && plan->n_levels <= UD3D_UNET_MAX_LEVELS && plan->block_reps >= 1 &&
                     plan->block_reps <= UD3D_UNET_MAX_REPS,
                 "%s: need 1..%d levels and 1..%d blocks per stage", who, UD3D_UNET_MAX_LEVELS, UD3D_UNET_MAX_REPS);
  for (int l = 0; l < plan->n_levels; ++l) {
    UD3D_CHECK_ARG(plan->level[l].c > 0 && plan->level[l].c % 32 == 0, "%s: level %d: channel count must be a multiple of 32", who, l);
    UD3D_CHECK_ARG(lv[l].n > 0, "%s: level %d has no voxels", who, l);
  }
  return UD3D_OK;
}

}  // namespace

extern "C" {

size_t ud3d_unet_workspace_bytes(const ud3d_unet_plan* plan, const ud3d_unet_tables* lv) {
  if (!plan || !lv || plan->n_levels < 1 || plan->n_levels > UD3D_UNET_MAX_LEVELS) return 0;
  Bump b{nullptr};
  for (int l = 0; l < plan->n_levels; ++l) carve(b, plan, lv, l);
  return b.top + 256;
}

int ud3d_unet_forward(const ud3d_unet_plan* plan, const ud3d_unet_tables* lv, const float* x_raw, const float* x_act, float* out_raw,
                      float* const* level_out, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_plan(plan, lv, "ud3d_unet_forward");
  if (rc) return rc;
  UD3D_CHECK_ARG(x_raw && x_act && out_raw && ws, "ud3d_unet_forward: NULL argument");
  UD3D_CHECK_ARG(((uintptr_t)ws & 255) == 0, "ud3d_unet_forward: workspace must be 256-byte aligned");
  if (ws_bytes < ud3d_unet_workspace_bytes(plan, lv)) {
    set_error("ud3d_unet_forward: workspace too small");
    return UD3D_EWORKSPACE;
  }
  for (int l = 0; l < plan->n_levels; ++l) {
    const ud3d_unet_tables& t = lv[l];
    UD3D_CHECK_ARG(t.subm && t.subm_mask, "ud3d_unet_forward: level %d: SubM3 table / tile mask missing", l);
    if (l + 1 < plan->n_levels)
      UD3D_CHECK_ARG(t.child && t.child_mask && t.up && t.up_mask, "ud3d_unet_forward: level %d: strided-conv tables missing", l);
  }
  Runner r{plan, lv, level_out, stream, Bump{(uint8_t*)ws}};
  return r.level(0, x_raw, x_act, out_raw, nullptr, nullptr, nullptr);
}

}  // extern "C"
