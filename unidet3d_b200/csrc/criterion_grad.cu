// Gradients of the criterion: d det_loss / d logits and d det_loss / d boxes of one
// (decoder layer, scene), given the match matrix and the sums ud3d_criterion_layer produced for it.
//   reference: unidet3d/criterion.py:86-142 under torch.autograd -- weighted cross-entropy (F.cross_entropy with the
//   class-weight vector [1, ..., 1, non_object_weight], mean = sum(w_i nll_i) / sum(w_i)) and the mean DIoU loss of the
//   matched pairs (axis_aligned_iou_loss.py:14-53 / rotated_iou_loss.py:14-82).
// One warp per query: lanes stride over the C + 1 logits (softmax - one-hot, scaled) and over the G ground-truth columns
// of the match matrix (forward-mode derivative of the pair's DIoU, box_loss.cuh); fixed-order shuffle reductions, no
// atomics: deterministic.  T <= ~4096 rows x <= 85 logits and a few hundred matched pairs: latency-bound.
#include "common.cuh"
#include "box_loss.cuh"

#include <float.h>

namespace ud3d {

__global__ void __launch_bounds__(128) crit_grad_kernel(ud3d_criterion_grad_args a) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (q >= a.T) return;
  const int C = a.C1 - 1;
  const int dim = a.box_dim;
  // ---- matched ground truths of this query: target label (the largest matched index wins, criterion.py:96) and
  //      the sum of the pair gradients
  int last = -1;
  float gb[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int g = lane; g < a.G; g += 32) {
    if (!a.match[(size_t)q * a.G + g]) continue;
    last = g;
    float gr[7];
    bl::pair_loss_grad<float>(a.boxes + (size_t)q * dim, a.gt_boxes + (size_t)g * dim, dim, gr);
    for (int k = 0; k < dim; ++k) gb[k] += gr[k];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
#pragma unroll
    for (int k = 0; k < 7; ++k) gb[k] += __shfl_xor_sync(0xffffffffu, gb[k], o);
  }
  const float n_pairs = a.sums[3];
  const float bscale = n_pairs > 0.f ? a.scales[1] / n_pairs : 0.f;
  if (lane < dim) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 7; ++k)
      if (k == lane) v = gb[k];
    a.d_boxes[(size_t)q * dim + lane] = bscale * v;
  }
  // ---- weighted cross-entropy
  const int target = last >= 0 ? (int)a.gt_labels[last] : C;
  const float w = target == C ? a.non_object_weight : 1.f;
  const float* row = a.logits + (size_t)q * a.ld_logits;
  float m = -FLT_MAX;
  for (int c = lane; c < a.C1; c += 32) m = fmaxf(m, row[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int c = lane; c < a.C1; c += 32) s += expf(row[c] - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float cscale = a.scales[0] * w / a.sums[1];
  float* out = a.d_logits + (size_t)q * a.ld_dlogits;
  for (int c = lane; c < a.C1; c += 32) out[c] = cscale * (expf(row[c] - m) / s - (c == target ? 1.f : 0.f));
}

// Backward of one scene's head outputs (encoder.py:165-201): the per-dataset class column gather and the box decode.
//   d_logits[t, :] = 0;  d_logits[t, cols[j]] = d_cls[t, j]         (cols are distinct: a gather, encoder.py:191-194)
//   d_raw[t, :]    = J_decode(raw[t])^T d_box[t]                    (PredBBox exp + _bbox_pred_to_bbox, encoder.py:109-111,241-283)
// One thread per query row; a NULL d_cls / d_box means "no gradient" (zeros are written).
__global__ void __launch_bounds__(128) head_bwd_kernel(const float* __restrict__ raw, int ld_raw, const float* __restrict__ d_box,
                                                       int with_angle, const float* __restrict__ d_cls, const int32_t* __restrict__ cols,
                                                       int n_cols, int T, float* __restrict__ d_raw, int ld_draw,
                                                       float* __restrict__ d_logits, int ld_dlogits, int n_union) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  float* dl = d_logits + (size_t)t * ld_dlogits;
  for (int c = 0; c < n_union; ++c) dl[c] = 0.f;
  if (d_cls)
    for (int j = 0; j < n_cols; ++j) dl[cols[j]] = d_cls[(size_t)t * n_cols + j];
  float out[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (d_box) {
    const int dim = with_angle ? 7 : 6;
    float r[8], g[7];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = raw[(size_t)t * ld_raw + i];
    for (int j = 0; j < dim; ++j) g[j] = d_box[(size_t)t * dim + j];
    bl::bbox_decode_backward<float>(r, with_angle != 0, g, out);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) d_raw[(size_t)t * ld_draw + i] = out[i];
}

}  // namespace ud3d

using namespace ud3d;

extern "C" int ud3d_head_backward(const float* raw, int ld_raw, const float* d_box, int with_angle, const float* d_cls, const int32_t* cols,
                                  int n_cols, int T, float* d_raw, int ld_draw, float* d_logits, int ld_dlogits, int n_union,
                                  void* stream) {
  UD3D_CHECK_ARG(raw && d_raw && d_logits && ld_raw >= 8 && ld_draw >= 8 && T >= 0 && n_union > 0 && ld_dlogits >= n_union,
                 "ud3d_head_backward: bad argument");
  UD3D_CHECK_ARG(!d_cls || (cols && n_cols > 0 && n_cols <= n_union), "ud3d_head_backward: d_cls needs the column list");
  if (T == 0) return UD3D_OK;
  head_bwd_kernel<<<cdiv(T, 128), 128, 0, (cudaStream_t)stream>>>(raw, ld_raw, d_box, with_angle, d_cls, cols, n_cols, T, d_raw, ld_draw,
                                                                d_logits, ld_dlogits, n_union);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

extern "C" int ud3d_criterion_layer_grad(const ud3d_criterion_grad_args* a, void* stream) {
  UD3D_CHECK_ARG(a && a->logits && a->boxes && a->sums && a->scales && a->d_logits && a->d_boxes, "ud3d_criterion_layer_grad: NULL argument");
  UD3D_CHECK_ARG(a->T > 0 && a->C1 >= 2 && a->ld_logits >= a->C1 && a->ld_dlogits >= a->C1 && a->G >= 0, "ud3d_criterion_layer_grad: bad sizes");
  UD3D_CHECK_ARG(a->box_dim == 6 || a->box_dim == 7, "ud3d_criterion_layer_grad: box_dim must be 6 or 7");
  UD3D_CHECK_ARG(a->G == 0 || (a->gt_boxes && a->gt_labels && a->match), "ud3d_criterion_layer_grad: NULL GT argument");
  crit_grad_kernel<<<cdiv(a->T, 4), 128, 0, (cudaStream_t)stream>>>(*a);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}
