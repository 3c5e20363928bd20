// Detection evaluator on the GPU: the reference's indoor_eval (unidet3d/indoor_eval.py:56-160,
// eval_det_cls + average_precision 'area') for ALL classes and IoU thresholds of a result set at once, so that boxes
// never leave the device between predict() and the metric.
//   1. eval_match : one thread per detection -- 3-D IoU (mmdet3d BaseInstance3DBoxes.overlaps: height overlap x BEV
//                   overlap; interval arithmetic for yaw == 0 pairs, the rotated-rectangle clipper of boxes.cuh
//                   otherwise) against the ground truth of its image and class; first maximum wins (strict '>').
//   2. eval_claim : the reference walks the detections of a class in descending score order and lets the FIRST
//                   detection whose best ground-truth box is still free take it.  Order-free formulation: a detection
//                   is a true positive iff it has the smallest rank among the detections that point at the same box
//                   with IoU above the threshold -> one atomicMin per (threshold, detection).  Deterministic.
//   3. eval_ap    : one CTA per (class, threshold): inclusive scan of the TP flags (recall / precision, fp = rank + 1 -
//                   tp), reverse max-scan (the monotone precision envelope), area under the curve, in double like numpy.
#include "boxes.cuh"
#include "common.cuh"

namespace ud3d {

__device__ __forceinline__ float iou3d_pair(const float* a, const float* b) {
  const float top = fminf(a[2] + a[5] * 0.5f, b[2] + b[5] * 0.5f), bot = fmaxf(a[2] - a[5] * 0.5f, b[2] - b[5] * 0.5f);
  const float oh = fmaxf(top - bot, 0.f);
  float ba[7], bb[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) { ba[k] = a[k]; bb[k] = b[k]; }
  ba[3] = fmaxf(ba[3], 1e-4f); ba[4] = fmaxf(ba[4], 1e-4f);
  bb[3] = fmaxf(bb[3], 1e-4f); bb[4] = fmaxf(bb[4], 1e-4f);
  float inter;
  if (a[6] == 0.f && b[6] == 0.f) {
    const float wx = fmaxf(fminf(ba[0] + ba[3] * 0.5f, bb[0] + bb[3] * 0.5f) - fmaxf(ba[0] - ba[3] * 0.5f, bb[0] - bb[3] * 0.5f), 0.f);
    const float wy = fmaxf(fminf(ba[1] + ba[4] * 0.5f, bb[1] + bb[4] * 0.5f) - fmaxf(ba[1] - ba[4] * 0.5f, bb[1] - bb[4] * 0.5f), 0.f);
    inter = wx * wy;
  } else {
    inter = box_overlap_rot(ba, bb, 1e-5f);
  }
  const float o3 = inter * oh;
  const float v1 = a[3] * a[4] * a[5], v2 = b[3] * b[4] * b[5];
  return o3 / fmaxf(v1 + v2 - o3, 1e-8f);
}

__global__ void eval_count_gt_kernel(const int32_t* __restrict__ gt_labels, int G, int n_cls, int32_t* __restrict__ npos) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < G; j += gridDim.x * blockDim.x) {
    const int c = gt_labels[j];
    if (c >= 0 && c < n_cls) atomicAdd(npos + c, 1);
  }
}

__global__ void eval_match_kernel(const float* __restrict__ det_boxes, const int32_t* __restrict__ det_labels,
                                  const int32_t* __restrict__ det_img, int D, const float* __restrict__ gt_boxes,
                                  const int32_t* __restrict__ gt_labels, const int32_t* __restrict__ gt_img_offsets, int n_img,
                                  float* __restrict__ iou_max, int32_t* __restrict__ jmax) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const int img = det_img[d], lab = det_labels[d];
  float best = -INFINITY;
  int bj = -1;
  if (img >= 0 && img < n_img) {
    float box[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) box[k] = det_boxes[(size_t)d * 7 + k];
    for (int j = gt_img_offsets[img]; j < gt_img_offsets[img + 1]; ++j) {
      if (gt_labels[j] != lab) continue;
      const float v = iou3d_pair(box, gt_boxes + (size_t)j * 7);
      if (v > best) { best = v; bj = j; }
    }
  }
  iou_max[d] = best;
  jmax[d] = bj;
}

__global__ void eval_claim_kernel(const int32_t* __restrict__ order, int D, const float* __restrict__ iou_max,
                                  const int32_t* __restrict__ jmax, int G, const float* __restrict__ thr, int n_thr,
                                  int32_t* __restrict__ best_rank) {
  const long long total = (long long)D * n_thr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i % D), t = (int)(i / D);
    const int d = order[r];
    if (jmax[d] >= 0 && iou_max[d] > thr[t]) atomicMin(best_rank + (size_t)t * G + jmax[d], r);
  }
}

// block-wide inclusive scans over 1024 threads (sum of ints / max of doubles), carry handled by the caller
__device__ __forceinline__ int block_scan_sum(int v, int* warp_tot) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  if (lane == 31) warp_tot[w] = v;
  __syncthreads();
  if (w == 0) {
    int t = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += u;
    }
    warp_tot[lane] = t;
  }
  __syncthreads();
  const int r = v + (w > 0 ? warp_tot[w - 1] : 0);
  __syncthreads();
  return r;
}
__device__ __forceinline__ double block_scan_max(double v, double* warp_tot) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v = fmax(v, t);
  }
  if (lane == 31) warp_tot[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double u = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t = fmax(t, u);
    }
    warp_tot[lane] = t;
  }
  __syncthreads();
  const double r = w > 0 ? fmax(v, warp_tot[w - 1]) : v;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(1024) eval_ap_kernel(const int32_t* __restrict__ order, const int32_t* __restrict__ class_offsets,
                                                       int n_cls, int D, const float* __restrict__ iou_max,
                                                       const int32_t* __restrict__ jmax, int G, const float* __restrict__ thr,
                                                       const int32_t* __restrict__ best_rank, const int32_t* __restrict__ npos,
                                                       int32_t* __restrict__ tp_cum, float* __restrict__ ap, double* __restrict__ rec) {
  __shared__ int s_int[32];
  __shared__ double s_dbl[32];
  __shared__ double s_acc;
  const int c = blockIdx.x, t = blockIdx.y;
  const int beg = class_offsets[c], end = class_offsets[c + 1];
  const int n = end - beg;
  const double np_ = (double)npos[c];
  int32_t* cum = tp_cum + (size_t)t * D;
  const float th = thr[t];
  const int32_t* best = best_rank + (size_t)t * G;
  // forward: inclusive count of true positives in score order
  int carry = 0;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    int flag = 0;
    if (i < n) {
      const int r = beg + i, d = order[r];
      flag = (jmax[d] >= 0 && iou_max[d] > th && best[jmax[d]] == r) ? 1 : 0;
    }
    const int inc = block_scan_sum(flag, s_int) + carry;
    if (i < n) cum[beg + i] = inc;
    __syncthreads();
    if (threadIdx.x == 1023) s_int[0] = inc;
    __syncthreads();
    carry = s_int[0];
    __syncthreads();
  }
  // backward: precision envelope (suffix maximum) and the area: sum over i of (rec_i - rec_{i-1}) * max_{k >= i} prec_k
  if (threadIdx.x == 0) s_acc = 0.0;
  __syncthreads();
  double carry_max = 0.0;
  double local = 0.0;
  for (int base = 0; base < n; base += 1024) {
    const int k = base + threadIdx.x;           // k-th element from the end
    const int i = n - 1 - k;
    double prec = 0.0;
    int tpc = 0, prev = 0;
    if (k < n) {
      tpc = cum[beg + i];
      prev = i > 0 ? cum[beg + i - 1] : 0;
      prec = (double)tpc / fmax((double)(i + 1), 2.220446049250313e-16);
    }
    const double env = fmax(block_scan_max(prec, s_dbl), carry_max);
    if (k < n && tpc != prev) local += ((double)tpc / np_ - (double)prev / np_) * env;
    __syncthreads();
    if (threadIdx.x == 1023) s_dbl[0] = env;
    __syncthreads();
    carry_max = s_dbl[0];
    __syncthreads();
  }
  // block reduction of the area
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) s_dbl[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < 32; ++w) a += s_dbl[w];
    const int tp_total = n > 0 ? cum[end - 1] : 0;
    // npos == 0: recall = 0 / 0 = nan in the reference, and so is the area (np.nanmean then skips the class)
    ap[c * gridDim.y + t] = (n > 0 && npos[c] == 0) ? __int_as_float(0x7fc00000) : (float)a;
    rec[c * gridDim.y + t] = n == 0 ? 0.0 : (double)tp_total / np_;
  }
}

}  // namespace ud3d

using namespace ud3d;

extern "C" {

size_t ud3d_eval_workspace_bytes(int D, int G, int n_thr) {
  const size_t d = (size_t)(D > 0 ? D : 1), g = (size_t)(G > 0 ? G : 1), t = (size_t)(n_thr > 0 ? n_thr : 1);
  return align_up(d * 4, 256) * 2 + align_up(t * g * 4, 256) + align_up(t * d * 4, 256) + align_up(t * 4, 256);
}

int ud3d_eval_detections(const float* det_boxes, const int32_t* det_labels, const int32_t* det_img, int D, const int32_t* order,
                         const int32_t* class_offsets, int n_cls, const float* gt_boxes, const int32_t* gt_labels,
                         const int32_t* gt_img_offsets, int G, int n_img, const float* thr_host, int n_thr, float* ap, double* rec,
                         int32_t* npos, void* ws, size_t ws_bytes, void* stream) {
  UD3D_CHECK_ARG(class_offsets && gt_img_offsets && thr_host && ap && rec && npos && ws, "ud3d_eval_detections: NULL argument");
  UD3D_CHECK_ARG(D >= 0 && G >= 0 && n_cls > 0 && n_img >= 0 && n_thr > 0 && n_thr <= 16, "ud3d_eval_detections: bad sizes (1..16 thresholds)");
  UD3D_CHECK_ARG(D == 0 || (det_boxes && det_labels && det_img && order), "ud3d_eval_detections: NULL detections");
  UD3D_CHECK_ARG(G == 0 || (gt_boxes && gt_labels), "ud3d_eval_detections: NULL ground truth");
  if (ws_bytes < ud3d_eval_workspace_bytes(D, G, n_thr)) {
    set_error("ud3d_eval_detections: workspace too small");
    return UD3D_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t d = (size_t)(D > 0 ? D : 1), g = (size_t)(G > 0 ? G : 1);
  uint8_t* p = (uint8_t*)ws;
  float* iou_max = (float*)p; p += align_up(d * 4, 256);
  int32_t* jmax = (int32_t*)p; p += align_up(d * 4, 256);
  int32_t* best_rank = (int32_t*)p; p += align_up((size_t)n_thr * g * 4, 256);
  int32_t* tp_cum = (int32_t*)p; p += align_up((size_t)n_thr * d * 4, 256);
  float* thr = (float*)p;
  UD3D_CUDA(cudaMemcpyAsync(thr, thr_host, (size_t)n_thr * 4, cudaMemcpyHostToDevice, st));
  UD3D_CUDA(cudaMemsetAsync(best_rank, 0x7f, (size_t)n_thr * g * 4, st));
  UD3D_CUDA(cudaMemsetAsync(npos, 0, (size_t)n_cls * 4, st));
  if (G > 0) {
    eval_count_gt_kernel<<<cdiv(G, 256), 256, 0, st>>>(gt_labels, G, n_cls, npos);
    UD3D_LAUNCH_CHECK();
  }
  if (D > 0) {
    eval_match_kernel<<<cdiv(D, 128), 128, 0, st>>>(det_boxes, det_labels, det_img, D, gt_boxes, gt_labels, gt_img_offsets, n_img, iou_max, jmax);
    UD3D_LAUNCH_CHECK();
    long long total = (long long)D * n_thr;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    eval_claim_kernel<<<blocks, 256, 0, st>>>(order, D, iou_max, jmax, G, thr, n_thr, best_rank);
    UD3D_LAUNCH_CHECK();
  }
  eval_ap_kernel<<<dim3(n_cls, n_thr), 1024, 0, st>>>(order, class_offsets, n_cls, D, iou_max, jmax, G, thr, best_rank, npos, tp_cum, ap, rec);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

}  // extern "C"
