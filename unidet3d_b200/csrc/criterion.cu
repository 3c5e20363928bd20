// Training-side targets, matcher and loss values: forward values only.
//   reference: unidet3d/unidet3d.py:220-275 (get_bboxes_by_masks), :371-409 (get_targets),
//              unidet3d/criterion.py:44-320 (criterion, costs, UniMatcher),
//              unidet3d/axis_aligned_iou_loss.py:14-53, unidet3d/rotated_iou_loss.py:14-82 (DIoU).
// Everything here is latency-bound integer / small-float work (T <= 3000 queries x G <= ~100 boxes per scene and
// layer): one CTA per GT column for the k-th-smallest selection, one thread per query for matching + loss terms,
// fixed-order block reductions (deterministic).
#include "common.cuh"
#include "boxes.cuh"

#include <float.h>

namespace ud3d {

constexpr float kCostInf = 1e8f;      // UniMatcher.inf / get_targets float_max
constexpr int kMaxTopk = 15;          // topk + 1 <= 16 candidates kept per thread

// ---------------------------------------------------------------- DIoU losses
// axis-aligned: boxes (centre, size); `penalty_t` = the GT box the centre-distance penalty is taken against
__device__ __forceinline__ float diou_aligned(const float* p, const float* t, const float* penalty_t) {
  float p1[3], p2[3], t1[3], t2[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    p1[a] = p[a] - p[a + 3] / 2; p2[a] = p[a] + p[a + 3] / 2;
    t1[a] = t[a] - t[a + 3] / 2; t2[a] = t[a] + t[a + 3] / 2;
  }
  const float a1 = (p2[0] - p1[0]) * (p2[1] - p1[1]) * (p2[2] - p1[2]);
  const float a2 = (t2[0] - t1[0]) * (t2[1] - t1[1]) * (t2[2] - t1[2]);
  float ov = 1.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) ov *= fmaxf(fminf(p2[a], t2[a]) - fmaxf(p1[a], t1[a]), 0.f);
  const float uni = fmaxf(a1 + a2 - ov, 1e-6f);
  const float iou_loss = 1.f - ov / uni;
  float r2 = 0.f, c2 = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float q1 = penalty_t[a] - penalty_t[a + 3] / 2, q2 = penalty_t[a] + penalty_t[a + 3] / 2;
    const float pc = (p1[a] + p2[a]) / 2, tc = (q1 + q2) / 2;
    r2 += (pc - tc) * (pc - tc);
    const float lo = fminf(p1[a], q1), hi = fmaxf(p2[a], q2);
    c2 += (lo - hi) * (lo - hi);
  }
  return iou_loss + r2 / c2;
}

// rotated: boxes (x, y, z, w, h, l, alpha)
__device__ float diou_rotated(const float* p, const float* t) {
  const float inter = box_overlap_rot(p, t, 1e-6f);
  const float zmax1 = p[2] + p[5] * 0.5f, zmin1 = p[2] - p[5] * 0.5f;
  const float zmax2 = t[2] + t[5] * 0.5f, zmin2 = t[2] - t[5] * 0.5f;
  const float z_ov = fmaxf(fminf(zmax1, zmax2) - fmaxf(zmin1, zmin2), 0.f);
  const float inter3 = inter * z_ov;
  const float union3 = p[3] * p[4] * p[5] + t[3] * t[4] * t[5] - inter3;
  P2 c1[5], c2[5];
  box_corners(p, c1);
  box_corners(t, c2);
  float x_max = -FLT_MAX, x_min = FLT_MAX, y_max = -FLT_MAX, y_min = FLT_MAX;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    x_max = fmaxf(x_max, fmaxf(c1[k].x, c2[k].x)); x_min = fminf(x_min, fminf(c1[k].x, c2[k].x));
    y_max = fmaxf(y_max, fmaxf(c1[k].y, c2[k].y)); y_min = fminf(y_min, fminf(c1[k].y, c2[k].y));
  }
  const float z_max = fmaxf(zmax1, zmax2), z_min = fminf(zmin1, zmin2);
  // r2 over (x, y, w) of the BEV boxes (rotated_iou_loss.py:24-25,61)
  const float r2 = (p[0] - t[0]) * (p[0] - t[0]) + (p[1] - t[1]) * (p[1] - t[1]) + (p[3] - t[3]) * (p[3] - t[3]);
  const float cc = (x_min - x_max) * (x_min - x_max) + (y_min - y_max) * (y_min - y_max) + (z_min - z_max) * (z_min - z_max);
  return 1.f - (inter3 / union3 - r2 / cc);
}

struct CritView {
  const float* logits; int ld; int T; int C1;
  const float* boxes; int dim;
  const float* gt; const int64_t* labels; int G;
  const uint8_t* qmask;
  float w_cls, w_box;
  const float* lse_max;   // [T] row max
  const float* lse_sum;   // [T] sum exp(x - max)
};

// matching cost of (query q, GT g); identical code path in the threshold and the match kernels
__device__ __forceinline__ float match_cost(const CritView& v, int q, int g) {
  if (!v.qmask[(size_t)g * v.T + q]) return kCostInf;
  const int lab = (int)v.labels[g];
  const float prob = expf(v.logits[(size_t)q * v.ld + lab] - v.lse_max[q]) / v.lse_sum[q];
  const float* pb = v.boxes + (size_t)q * v.dim;
  const float* gb = v.gt + (size_t)g * v.dim;
  const float box = v.dim == 7 ? diou_rotated(pb, gb) : diou_aligned(pb, gb, v.gt /* GT 0: axis_aligned_iou_loss.py:51 */);
  return -prob * v.w_cls + box * v.w_box;
}

__global__ void __launch_bounds__(256) crit_lse_kernel(const float* __restrict__ logits, int ld, int T, int C1,
                                                       float* __restrict__ row_max, float* __restrict__ row_sum) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= T) return;
  const float* r = logits + (size_t)q * ld;
  float m = -FLT_MAX;
  for (int c = 0; c < C1; ++c) m = fmaxf(m, r[c]);
  float s = 0.f;
  for (int c = 0; c < C1; ++c) s += expf(r[c] - m);
  row_max[q] = m;
  row_sum[q] = s;
}

// k-th smallest (k = kth, 1-based, with multiplicity) of f(i), i < n, over one CTA of 256 threads: every thread keeps
// its kth smallest values sorted ascending, then kth rounds of "global minimum of the heads".  Result in all threads.
template <class F>
__device__ float block_kth_smallest(int n, int kth, F f) {
  __shared__ float s_val[256];
  __shared__ int s_own[256];
  float best[kMaxTopk + 1];
#pragma unroll
  for (int j = 0; j <= kMaxTopk; ++j) best[j] = FLT_MAX;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float v = f(i);
    float last = FLT_MAX;
#pragma unroll
    for (int j = 0; j <= kMaxTopk; ++j)
      if (j == kth - 1) last = best[j];
    if (v < last) {
      // replace the largest kept value, then one pass of adjacent swaps restores the ascending order
#pragma unroll
      for (int j = 0; j <= kMaxTopk; ++j)
        if (j == kth - 1) best[j] = v;
#pragma unroll
      for (int j = kMaxTopk; j >= 1; --j) {
        if (j < kth && best[j] < best[j - 1]) {
          const float t = best[j];
          best[j] = best[j - 1];
          best[j - 1] = t;
        }
      }
    }
  }
  int head = 0;
  float result = FLT_MAX;
  for (int round = 0; round < kth; ++round) {
    float mine = FLT_MAX;
#pragma unroll
    for (int j = 0; j <= kMaxTopk; ++j)
      if (j == head) mine = best[j];
    if (head >= kth) mine = FLT_MAX;
    s_val[threadIdx.x] = mine;
    s_own[threadIdx.x] = threadIdx.x;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) {
        const float a = s_val[threadIdx.x], b = s_val[threadIdx.x + o];
        if (b < a) { s_val[threadIdx.x] = b; s_own[threadIdx.x] = s_own[threadIdx.x + o]; }
      }
      __syncthreads();
    }
    result = s_val[0];
    if (s_own[0] == (int)threadIdx.x) ++head;
    __syncthreads();
  }
  return result;
}

__global__ void __launch_bounds__(256) crit_threshold_kernel(CritView v, int topk, float* __restrict__ thr) {
  const int g = blockIdx.x;
  const float r = block_kth_smallest(v.T, topk + 1, [&](int q) { return match_cost(v, q, g); });
  if (threadIdx.x == 0) thr[g] = r;
}

__device__ __forceinline__ void block_sum4(float (&acc)[4], float* out4) {
  __shared__ float s_acc[8][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float x = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) s_acc[threadIdx.x >> 5][j] = x;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float x = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) x += s_acc[w][threadIdx.x];
    out4[threadIdx.x] = x;
  }
}

// one thread per query: matches, target label, CE and box-loss terms; per-CTA partial sums
__global__ void __launch_bounds__(256) crit_match_loss_kernel(CritView v, const float* __restrict__ thr, float non_object_weight,
                                                              uint8_t* __restrict__ match, float* __restrict__ partial) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (q < v.T) {
    const int C = v.C1 - 1;
    int target = C;
    for (int g = 0; g < v.G; ++g) {
      const float c = match_cost(v, q, g);
      const bool m = c < thr[g];
      match[(size_t)q * v.G + g] = m ? 1 : 0;
      if (m) {
        target = (int)v.labels[g];      // the largest matched GT index wins (criterion.py:96)
        const float* pb = v.boxes + (size_t)q * v.dim;
        const float* gb = v.gt + (size_t)g * v.dim;
        acc[2] += v.dim == 7 ? diou_rotated(pb, gb) : diou_aligned(pb, gb, gb);
        acc[3] += 1.f;
      }
    }
    const float w = target == C ? non_object_weight : 1.f;
    const float nll = (v.lse_max[q] + logf(v.lse_sum[q])) - v.logits[(size_t)q * v.ld + target];
    acc[0] = w * nll;
    acc[1] = w;
  }
  block_sum4(acc, partial + (size_t)blockIdx.x * 4);
}

__global__ void crit_final_kernel(const float* __restrict__ partial, int n_blocks, float* __restrict__ sums) {
  if (threadIdx.x < 4) {
    float x = 0.f;
    for (int b = 0; b < n_blocks; ++b) x += partial[(size_t)b * 4 + threadIdx.x];
    sums[threadIdx.x] = x;
  }
}

// ---------------------------------------------------------------- get_bboxes_by_masks
__device__ __forceinline__ int crit_f2ord(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float crit_ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void inst_box_init_kernel(int* aabb, int n_inst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_inst * 6) aabb[i] = (i % 6) < 3 ? crit_f2ord(INFINITY) : crit_f2ord(-INFINITY);
}
__global__ void inst_box_accum_kernel(const float* __restrict__ pts, int ld, const int64_t* __restrict__ inst, int n, int n_inst,
                                      int* aabb) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const long long id = inst[i];
    if (id < 0 || id >= n_inst) continue;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int o = crit_f2ord(pts[(size_t)i * ld + a]);
      atomicMin(aabb + id * 6 + a, o);
      atomicMax(aabb + id * 6 + 3 + a, o);
    }
  }
}
__global__ void inst_box_final_kernel(const int* __restrict__ aabb, int n_inst, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_inst * 3) return;
  const int b = i / 3, a = i % 3;
  const float lo = crit_ord2f(aabb[b * 6 + a]), hi = crit_ord2f(aabb[b * 6 + 3 + a]);
  out[b * 6 + a] = (hi + lo) / 2;
  out[b * 6 + 3 + a] = hi - lo;
}

// ---------------------------------------------------------------- get_targets
__device__ __forceinline__ float center_dist2(const float* __restrict__ centers, const float* __restrict__ gt, int dim, int s, int g) {
  float d = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float x = gt[(size_t)g * dim + a] - centers[(size_t)s * 3 + a];
    d += x * x;
  }
  return d;
}
__global__ void __launch_bounds__(256) target_threshold_kernel(const float* __restrict__ centers, int S, const float* __restrict__ gt,
                                                               int dim, int kth, float* __restrict__ thr) {
  const int g = blockIdx.x;
  const float r = block_kth_smallest(S, kth, [&](int s) { return center_dist2(centers, gt, dim, s, g); });
  if (threadIdx.x == 0) thr[g] = r;
}
__global__ void target_assign_kernel(const float* __restrict__ centers, int S, const float* __restrict__ gt, int dim, int G,
                                     const float* __restrict__ thr, uint8_t* __restrict__ masks) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  float best = kCostInf;
  int arg = -1;
  for (int g = 0; g < G; ++g) {
    const float d = center_dist2(centers, gt, dim, s, g);
    if (d < thr[g] && d < best) { best = d; arg = g; }
  }
  for (int g = 0; g < G; ++g) masks[(size_t)g * S + s] = g == arg ? 1 : 0;
}

}  // namespace ud3d

using namespace ud3d;

extern "C" {

int ud3d_boxes_by_instance(const float* points, int ld_pts, const int64_t* inst, int n, int n_inst, float* out, void* ws,
                           size_t ws_bytes, void* stream) {
  UD3D_CHECK_ARG(n >= 0 && n_inst >= 0 && ld_pts >= 3, "ud3d_boxes_by_instance: bad sizes");
  if (n_inst == 0) return UD3D_OK;
  UD3D_CHECK_ARG(points && inst && out && ws, "ud3d_boxes_by_instance: NULL argument");
  UD3D_CHECK_ARG(ws_bytes >= (size_t)n_inst * 6 * 4, "ud3d_boxes_by_instance: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int* aabb = (int*)ws;
  inst_box_init_kernel<<<cdiv(n_inst * 6, 256), 256, 0, st>>>(aabb, n_inst);
  UD3D_LAUNCH_CHECK();
  if (n > 0) {
    int blocks = cdiv(n, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    inst_box_accum_kernel<<<blocks, 256, 0, st>>>(points, ld_pts, inst, n, n_inst, aabb);
    UD3D_LAUNCH_CHECK();
  }
  inst_box_final_kernel<<<cdiv(n_inst * 3, 256), 256, 0, st>>>(aabb, n_inst, out);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int ud3d_targets_by_distance(const float* centers, int S, const float* gt_boxes, int box_dim, int G, int topk, uint8_t* masks,
                             void* ws, size_t ws_bytes, void* stream) {
  UD3D_CHECK_ARG(S >= 0 && G >= 0 && box_dim >= 3, "ud3d_targets_by_distance: bad sizes");
  UD3D_CHECK_ARG(topk >= 0 && topk <= kMaxTopk, "ud3d_targets_by_distance: topk must be in [0, 15]");
  if (S == 0 || G == 0) return UD3D_OK;
  UD3D_CHECK_ARG(centers && gt_boxes && masks && ws, "ud3d_targets_by_distance: NULL argument");
  UD3D_CHECK_ARG(ws_bytes >= (size_t)G * 4, "ud3d_targets_by_distance: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* thr = (float*)ws;
  const int kth = topk + 1 < S ? topk + 1 : S;        // torch.topk(d, min(topk + 1, S)).values[-1]
  target_threshold_kernel<<<G, 256, 0, st>>>(centers, S, gt_boxes, box_dim, kth, thr);
  UD3D_LAUNCH_CHECK();
  target_assign_kernel<<<cdiv(S, 256), 256, 0, st>>>(centers, S, gt_boxes, box_dim, G, thr, masks);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

size_t ud3d_criterion_workspace_bytes(int T, int G) {
  const size_t t = (size_t)(T > 0 ? T : 1), g = (size_t)(G > 0 ? G : 1);
  return align_up(t * 4, 256) * 2 + align_up(g * 4, 256) + align_up((size_t)cdiv(t, 256) * 16, 256);
}

int ud3d_criterion_layer(const ud3d_criterion_args* a, void* ws, size_t ws_bytes, void* stream) {
  UD3D_CHECK_ARG(a && a->logits && a->boxes && a->sums, "ud3d_criterion_layer: NULL argument");
  UD3D_CHECK_ARG(a->T > 0 && a->C1 >= 2 && a->ld_logits >= a->C1 && a->G >= 0, "ud3d_criterion_layer: bad sizes");
  UD3D_CHECK_ARG(a->box_dim == 6 || a->box_dim == 7, "ud3d_criterion_layer: box_dim must be 6 or 7");
  UD3D_CHECK_ARG(a->topk >= 0 && a->topk <= kMaxTopk, "ud3d_criterion_layer: topk must be in [0, 15]");
  UD3D_CHECK_ARG(a->G == 0 || (a->gt_boxes && a->gt_labels && a->query_masks && a->match), "ud3d_criterion_layer: NULL GT argument");
  UD3D_CHECK_ARG(a->G == 0 || a->T >= a->topk + 1, "ud3d_criterion_layer: needs T >= topk + 1 (torch.topk raises)");
  UD3D_CHECK_ARG(ws && ws_bytes >= ud3d_criterion_workspace_bytes(a->T, a->G), "ud3d_criterion_layer: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* p = (uint8_t*)ws;
  float* row_max = (float*)p; p += align_up((size_t)a->T * 4, 256);
  float* row_sum = (float*)p; p += align_up((size_t)a->T * 4, 256);
  float* thr = (float*)p; p += align_up((size_t)(a->G > 0 ? a->G : 1) * 4, 256);
  float* partial = (float*)p;
  const int qblocks = cdiv(a->T, 256);
  crit_lse_kernel<<<qblocks, 256, 0, st>>>(a->logits, a->ld_logits, a->T, a->C1, row_max, row_sum);
  UD3D_LAUNCH_CHECK();
  CritView v;
  v.logits = a->logits; v.ld = a->ld_logits; v.T = a->T; v.C1 = a->C1;
  v.boxes = a->boxes; v.dim = a->box_dim;
  v.gt = a->gt_boxes; v.labels = a->gt_labels; v.G = a->G;
  v.qmask = a->query_masks;
  v.w_cls = a->w_cls; v.w_box = a->w_box;
  v.lse_max = row_max; v.lse_sum = row_sum;
  if (a->G > 0) {
    crit_threshold_kernel<<<a->G, 256, 0, st>>>(v, a->topk, thr);
    UD3D_LAUNCH_CHECK();
  }
  crit_match_loss_kernel<<<qblocks, 256, 0, st>>>(v, thr, a->non_object_weight, a->match, partial);
  UD3D_LAUNCH_CHECK();
  crit_final_kernel<<<1, 32, 0, st>>>(partial, qblocks, a->sums);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

}  // extern "C"
