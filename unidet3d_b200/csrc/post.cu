// Post-processing: class softmax + global top-k, batched multi-class 3D NMS (rotated BEV / axis-aligned
// BEV / axis-aligned 3-D IoU, one launch for all classes, on-device sweep, no host sync), and
// superpoint box trimming (point-in-box vote + masked AABB) without the reference's
// [n_pts, n_boxes, 6] temporary.  Reference: unidet3d/unidet3d.py:475-677.
#include "common.cuh"
#include "boxes.cuh"

#include <float.h>

namespace ud3d {

// ---------------------------------------------------------------- softmax scores (drop last column)
__global__ void softmax_scores_kernel(const float* __restrict__ logits, int ld, int T, int C1, float* __restrict__ scores) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const float* l = logits + (size_t)t * ld;
  float m = -INFINITY;
  for (int c = 0; c < C1; ++c) m = fmaxf(m, l[c]);
  float s = 0.f;
  for (int c = 0; c < C1; ++c) s += expf(l[c] - m);
  const int C = C1 - 1;
  for (int c = 0; c < C; ++c) scores[(size_t)t * C + c] = expf(l[c] - m) / s;
}

// ---------------------------------------------------------------- top-k (k <= 1024), single CTA
// 64-bit composite key = (score bits << 32) | ~index : descending key order == descending score,
// ties broken towards the lower flat index; 8 radix-select passes of 8 bits, then a bitonic sort of
// the k survivors in shared memory.
__device__ __forceinline__ unsigned long long topk_key(const float* scores, int i) {
  return ((unsigned long long)__float_as_uint(scores[i]) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
}

__global__ void __launch_bounds__(1024) topk_select_kernel(const float* __restrict__ scores, int n, int C, int k,
                                                           float* __restrict__ out_scores, int32_t* __restrict__ out_labels,
                                                           int32_t* __restrict__ out_query) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long s_prefix, s_mask;
  __shared__ int s_remaining;
  __shared__ unsigned long long cand[1024];
  __shared__ int s_cnt;
  const int tid = threadIdx.x;
  if (tid == 0) {
    s_prefix = 0ull;
    s_mask = 0ull;
    s_remaining = k;
    s_cnt = 0;
  }
  __syncthreads();
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    const unsigned long long prefix = s_prefix, mask = s_mask;
    for (int i = tid; i < n; i += 1024) {
      unsigned long long key = topk_key(scores, i);
      if ((key & mask) == prefix) atomicAdd(&hist[(unsigned)(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid < 32) {
      // suffix scan over the 256 bins by one warp: lane l owns bins [8l, 8l+8)
      const int rem = s_remaining;
      unsigned int loc[8], tot = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        loc[j] = hist[tid * 8 + j];
        tot += loc[j];
      }
      // inclusive suffix sum over lanes (higher lanes = larger digits)
      unsigned int suf = tot;
      for (int o = 1; o < 32; o <<= 1) {
        unsigned int y = __shfl_down_sync(0xffffffffu, suf, o);
        if (tid + o < 32) suf += y;
      }
      const unsigned int above = suf - tot;           // elements in bins owned by higher lanes
      const bool mine = above < (unsigned)rem && suf >= (unsigned)rem;
      if (mine) {
        int d = 7;
        unsigned int a = above;
        for (; d > 0; --d) {
          if (a + loc[d] >= (unsigned)rem) break;
          a += loc[d];
        }
        const bool exact = a + loc[d] == (unsigned)rem;   // the whole bin is selected: no finer digit needed
        s_remaining = exact ? 0 : rem - (int)a;
        s_prefix = prefix | ((unsigned long long)(tid * 8 + d) << shift);
        s_mask = mask | (0xFFull << shift);
      }
    }
    __syncthreads();
    if (s_remaining == 0) break;
  }
  const unsigned long long kth = s_prefix;   // exact key of the k-th largest element (keys are unique)
  for (int i = tid; i < n; i += 1024) {
    unsigned long long key = topk_key(scores, i);
    if (key >= kth) {
      int pos = atomicAdd(&s_cnt, 1);
      if (pos < 1024) cand[pos] = key;
    }
  }
  __syncthreads();
  const int cnt = min(s_cnt, 1024);
  if (tid >= cnt) cand[tid] = 0ull;
  __syncthreads();
  // bitonic sort, descending
  for (int size = 2; size <= 1024; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      int partner = tid ^ stride;
      if (partner > tid) {
        bool desc = (tid & size) == 0;
        unsigned long long x = cand[tid], y = cand[partner];
        if ((x < y) == desc) {
          cand[tid] = y;
          cand[partner] = x;
        }
      }
      __syncthreads();
    }
  }
  if (tid < k) {
    unsigned long long key = cand[tid];
    int idx = (int)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
    out_scores[tid] = __uint_as_float((uint32_t)(key >> 32));
    out_labels[tid] = idx % C;
    out_query[tid] = idx / C;
  }
}

// returns true when box j must be suppressed by kept box i
__device__ __forceinline__ bool nms_suppress(const float* bi, const float* bj, int mode, float thr) {
  const float EPS = 1e-8f;
  if (mode == 0) {
    float sa = bi[3] * bi[4], sb = bj[3] * bj[4];
    float so = box_overlap_rot(bi, bj);
    return so / fmaxf(sa + sb - so, EPS) > thr;
  } else if (mode == 1) {
    float left = fmaxf(bi[0] - bi[3] / 2, bj[0] - bj[3] / 2), right = fminf(bi[0] + bi[3] / 2, bj[0] + bj[3] / 2);
    float top = fmaxf(bi[1] - bi[4] / 2, bj[1] - bj[4] / 2), bottom = fminf(bi[1] + bi[4] / 2, bj[1] + bj[4] / 2);
    float w = fmaxf(right - left, 0.f), hgt = fmaxf(bottom - top, 0.f);
    float inter = w * hgt;
    return inter / fmaxf(bi[3] * bi[4] + bj[3] * bj[4] - inter, EPS) > thr;
  } else {
    // mmdet3d aligned_3d_nms on corners of (centre,size) boxes (criterion.py:180-198 _bbox_to_loss)
    float a1[3], a2[3], b1[3], b2[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      a1[d] = bi[d] - bi[3 + d] / 2; a2[d] = bi[d] + bi[3 + d] / 2;
      b1[d] = bj[d] - bj[3 + d] / 2; b2[d] = bj[d] + bj[3 + d] / 2;
    }
    float va = (a2[0] - a1[0]) * (a2[1] - a1[1]) * (a2[2] - a1[2]);
    float vb = (b2[0] - b1[0]) * (b2[1] - b1[1]) * (b2[2] - b1[2]);
    float inter = fmaxf(0.f, fminf(a2[0], b2[0]) - fmaxf(a1[0], b1[0])) * fmaxf(0.f, fminf(a2[1], b2[1]) - fmaxf(a1[1], b1[1])) *
                  fmaxf(0.f, fminf(a2[2], b2[2]) - fmaxf(a1[2], b1[2]));
    float iou = inter / (va + vb - inter);
    return !(iou <= thr);   // survivors are `iou <= thr`
  }
}

// ---------------------------------------------------------------- NMS stage 1: class-major order
// input is in descending score order; pos = (#valid with smaller label) + (#valid earlier with same label)
struct NmsWs {
  int32_t* sorted_idx;      // [n]
  int32_t* n_valid;         // [1]
  float* boxes7;            // [n,7] in sorted order (yaw 0 when box_dim == 6)
  int32_t* sorted_label;    // [n]
  unsigned long long* mask; // [n, 16]
};
static inline size_t nms_ws_bytes(int n) {
  return align_up((size_t)n * 4, 256) + 256 + align_up((size_t)n * 28, 256) + align_up((size_t)n * 4, 256) +
         align_up((size_t)n * 16 * 8, 256);
}
static inline NmsWs nms_ws_view(void* ws, int n) {
  NmsWs v;
  char* p = (char*)ws;
  v.sorted_idx = (int32_t*)p; p += align_up((size_t)n * 4, 256);
  v.n_valid = (int32_t*)p; p += 256;
  v.boxes7 = (float*)p; p += align_up((size_t)n * 28, 256);
  v.sorted_label = (int32_t*)p; p += align_up((size_t)n * 4, 256);
  v.mask = (unsigned long long*)p;
  return v;
}

constexpr int kNmsMaxLabel = 2048;   // labels must be in [0, kNmsMaxLabel)

__global__ void __launch_bounds__(1024) nms_order_kernel(const float* __restrict__ boxes, int box_dim,
                                                         const float* __restrict__ scores,
                                                         const int32_t* __restrict__ labels, int n, float score_thr,
                                                         NmsWs w) {
  // stable counting sort by class of the score-sorted candidates: class histogram -> exclusive scan -> rank of each
  // element among the earlier elements of its class (warp match + a running per-class counter, warps in order)
  __shared__ int s_hist[kNmsMaxLabel];
  __shared__ int s_run[kNmsMaxLabel];
  __shared__ int s_part[32];
  const int i = threadIdx.x, lane = i & 31, wp = i >> 5;
  for (int c = i; c < kNmsMaxLabel; c += 1024) {
    s_hist[c] = 0;
    s_run[c] = 0;
  }
  __syncthreads();
  const int lab = i < n ? labels[i] : -1;
  const bool valid = i < n && scores[i] > score_thr && lab >= 0 && lab < kNmsMaxLabel;
  if (valid) atomicAdd(&s_hist[lab], 1);
  __syncthreads();
  // exclusive scan of the histogram (2 entries per thread)
  int h0 = s_hist[2 * i], h1 = s_hist[2 * i + 1];
  int x = h0 + h1;
  int incl = x;
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) s_part[wp] = incl;
  __syncthreads();
  if (wp == 0) {
    int t = s_part[lane];
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += y;
    }
    s_part[lane] = t;
  }
  __syncthreads();
  const int base = incl - x + (wp > 0 ? s_part[wp - 1] : 0);
  const int nv = s_part[31];
  __syncthreads();
  s_hist[2 * i] = base;            // class start offsets
  s_hist[2 * i + 1] = base + h0;
  __syncthreads();
  // rank within class, warps processed in order
  const unsigned same = __match_any_sync(0xffffffffu, valid ? lab : -1 - lane);
  const int intra = __popc(same & ((1u << lane) - 1u));
  int rank = 0;
  for (int ww = 0; ww < 32; ++ww) {
    if (wp == ww) {
      const int r0 = valid ? s_run[lab] : 0;          // every lane reads before any lane of the warp updates
      __syncwarp();
      if (valid) {
        rank = r0 + intra;
        if (intra == 0) s_run[lab] = r0 + __popc(same);   // the first lane of each label group advances the counter
      }
    }
    __syncthreads();
  }
  if (valid) {
    const int pos = s_hist[lab] + rank;
    w.sorted_idx[pos] = i;
    w.sorted_label[pos] = lab;
    const float* b = boxes + (size_t)i * box_dim;
    float* o = w.boxes7 + (size_t)pos * 7;
#pragma unroll
    for (int d = 0; d < 6; ++d) o[d] = b[d];
    o[6] = box_dim == 7 ? b[6] : 0.f;
  }
  if (i == 0) *w.n_valid = nv;
}

// stage 2: suppression bit matrix (64 x 64 blocks, same-class pairs, j > i)
__global__ void __launch_bounds__(64) nms_mask_kernel(NmsWs w, int n, int mode, float thr) {
  const int nv = *w.n_valid;
  const int rb = blockIdx.y, cb = blockIdx.x;
  if (rb * 64 >= nv || cb * 64 >= nv || cb < rb) return;
  __shared__ float sb[64 * 7];
  __shared__ int sl[64];
  const int csize = min(nv - cb * 64, 64);
  if ((int)threadIdx.x < csize) {
    for (int d = 0; d < 7; ++d) sb[threadIdx.x * 7 + d] = w.boxes7[(size_t)(cb * 64 + threadIdx.x) * 7 + d];
    sl[threadIdx.x] = w.sorted_label[cb * 64 + threadIdx.x];
  }
  __syncthreads();
  const int i = rb * 64 + threadIdx.x;
  if (i >= nv) return;
  float bi[7];
  for (int d = 0; d < 7; ++d) bi[d] = w.boxes7[(size_t)i * 7 + d];
  const int li = w.sorted_label[i];
  unsigned long long bits = 0ull;
  const int start = (rb == cb) ? threadIdx.x + 1 : 0;
  for (int j = start; j < csize; ++j) {
    if (sl[j] != li) continue;
    if (nms_suppress(bi, sb + j * 7, mode, thr)) bits |= 1ull << j;
  }
  w.mask[(size_t)i * 16 + cb] = bits;
}

// stage 3: greedy sweep.  Boxes only suppress boxes of their own class, and the class-major order makes every class a
// contiguous segment, so the segments are swept independently: warp w takes segments w, w+32, ... (bit matrix staged
// in shared memory); kept flags are then compacted in order (ascending class, descending score).
__global__ void __launch_bounds__(1024) nms_sweep_kernel(NmsWs w, int32_t* __restrict__ keep_out, int32_t* __restrict__ n_keep) {
  extern __shared__ unsigned long long smask[];
  __shared__ unsigned char s_keep[1024];
  __shared__ int s_segstart[1025];
  __shared__ int s_nseg;
  __shared__ int s_part[32];
  const int nv = *w.n_valid;
  const int nblk = (nv + 63) / 64;
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  for (int t = tid; t < nv * 16; t += blockDim.x) {
    int i = t >> 4, c = t & 15;
    smask[t] = (c < nblk && c >= (i >> 6)) ? w.mask[t] : 0ull;
  }
  s_keep[tid] = 0;
  if (tid == 0) s_nseg = 0;
  __syncthreads();
  // segment starts: positions where the label changes (ordered list via warp-aggregated append is not needed:
  // flag + ordered scan)
  const bool is_start = tid < nv && (tid == 0 || w.sorted_label[tid] != w.sorted_label[tid - 1]);
  {
    int f = is_start ? 1 : 0, incl = f;
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) s_part[wp] = incl;
    __syncthreads();
    if (wp == 0) {
      int t = s_part[lane];
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      s_part[lane] = t;
    }
    __syncthreads();
    const int pos = incl - f + (wp > 0 ? s_part[wp - 1] : 0);
    if (is_start) s_segstart[pos] = tid;
    if (tid == 0) {
      s_nseg = s_part[31];
    }
    __syncthreads();
    if (tid == 0) s_segstart[s_nseg] = nv;
    __syncthreads();
  }
  const int nseg = s_nseg;
  for (int sg = wp; sg < nseg; sg += 32) {
    const int st = s_segstart[sg], en = s_segstart[sg + 1];
    const int w0 = st >> 6;
    unsigned long long removed = 0ull;     // lane j owns 64-bit word w0 + j of the removed set
    for (int i = st; i < en; ++i) {
      const unsigned long long r = __shfl_sync(0xffffffffu, removed, (i >> 6) - w0);
      if (!((r >> (i & 63)) & 1ull)) {
        if (lane == 0) s_keep[i] = 1;
        if (w0 + lane < 16) removed |= smask[i * 16 + w0 + lane];
      }
    }
  }
  __syncthreads();
  // ordered compaction of the kept flags
  {
    int f = (tid < nv && s_keep[tid]) ? 1 : 0, incl = f;
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) s_part[wp] = incl;
    __syncthreads();
    if (wp == 0) {
      int t = s_part[lane];
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      s_part[lane] = t;
    }
    __syncthreads();
    const int pos = incl - f + (wp > 0 ? s_part[wp - 1] : 0);
    if (f) keep_out[pos] = w.sorted_idx[tid];
    if (tid == 0) *n_keep = s_part[31];
  }
}

// ---------------------------------------------------------------- superpoint trimming
// reference: trim_bboxes_by_superpoints (unidet3d.py:540-593).  Three phases, all data-parallel:
//   A  vote:   every (point, box) pair is tested once; inside -> red.add cnt[box][sp(point)]
//   B  decide: one warp per (box, superpoint) pair with cnt > 0 walks that superpoint's points
//              (CSR built once per scene) and applies  in = (inside && frac >= low) || frac > up,
//              warp-reduces the AABB and issues 6 atomics per pair
//   C  finalise: centre / size.
// Pairs with cnt == 0 have frac = 0 < low_thr, so none of their points can be selected: skipping
// them is exact.  Nothing of size n_pts x n_boxes is ever materialised (the reference builds a
// [n_pts, n_boxes, 6] fp32 tensor = 2.4 GB at 100k x 1000).
struct TrimWs {
  int* sp_size;     // [n_sp]
  int* sp_start;    // [n_sp + 1]
  int* cursor;      // [n_sp]
  int* sp_points;   // [n_pts]
  int* cnt;         // [m, n_sp]
  int* aabb;        // [m, 6] ordered-int encoded min xyz / max xyz
};
static inline size_t trim_ws_bytes(int n_sp, int n_pts, int m) {
  return align_up((size_t)n_sp * 4, 256) + align_up((size_t)(n_sp + 1) * 4, 256) + align_up((size_t)n_sp * 4, 256) +
         align_up((size_t)n_pts * 4, 256) + align_up((size_t)m * n_sp * 4, 256) + align_up((size_t)m * 24, 256);
}
static inline TrimWs trim_ws_view(void* ws, int n_sp, int n_pts, int m) {
  TrimWs v;
  char* p = (char*)ws;
  v.sp_size = (int*)p; p += align_up((size_t)n_sp * 4, 256);
  v.sp_start = (int*)p; p += align_up((size_t)(n_sp + 1) * 4, 256);
  v.cursor = (int*)p; p += align_up((size_t)n_sp * 4, 256);
  v.sp_points = (int*)p; p += align_up((size_t)n_pts * 4, 256);
  v.cnt = (int*)p; p += align_up((size_t)m * n_sp * 4, 256);
  v.aabb = (int*)p;
  return v;
}

__global__ void sp_size_kernel(const int64_t* __restrict__ sp, int n_pts, int n_sp, int* __restrict__ size) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_pts; p += gridDim.x * blockDim.x) {
    long long s = sp[p];
    if (s >= 0 && s < n_sp) atomicAdd(size + s, 1);
  }
}
// single CTA exclusive scan of sp_size -> sp_start (+ cursor copy)
__global__ void __launch_bounds__(1024) sp_scan_kernel(const int* __restrict__ size, int n_sp, int* __restrict__ start,
                                                       int* __restrict__ cursor) {
  __shared__ int sh[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < n_sp; base += 1024) {
    int i = base + threadIdx.x;
    int v = i < n_sp ? size[i] : 0;
    int x = v;
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) sh[w] = x;
    __syncthreads();
    if (w == 0) {
      int t = sh[lane];
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      sh[lane] = t;
    }
    __syncthreads();
    int incl = x + (w > 0 ? sh[w - 1] : 0) + carry;
    if (i < n_sp) {
      start[i] = incl - v;
      cursor[i] = incl - v;
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) start[n_sp] = carry;
}
__global__ void sp_fill_kernel(const int64_t* __restrict__ sp, int n_pts, int n_sp, int* cursor, int* __restrict__ sp_points) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_pts; p += gridDim.x * blockDim.x) {
    long long s = sp[p];
    if (s >= 0 && s < n_sp) sp_points[atomicAdd(cursor + s, 1)] = p;
  }
}

struct BoxP {
  float b[6];
  float c, s;   // cos(-yaw), sin(-yaw)
};
__device__ __forceinline__ BoxP load_box(const float* __restrict__ boxes, int box_dim, const int32_t* __restrict__ box_index,
                                         int m) {
  const float* src = boxes + (size_t)(box_index ? box_index[m] : m) * box_dim;
  BoxP q;
#pragma unroll
  for (int d = 0; d < 6; ++d) q.b[d] = src[d];
  float yaw = box_dim == 7 ? src[6] : 0.f;
  q.c = cosf(-yaw);
  q.s = sinf(-yaw);
  return q;
}
__device__ __forceinline__ bool point_in_box(float px, float py, float pz, const BoxP& q) {
  // get_face_distances (unidet3d.py:652-677): shift rotated by -yaw about z, then six face distances
  const float* b = q.b;
  float sx = px - b[0], sy = py - b[1], sz = pz - b[2];
  float rx = __fsub_rn(__fmul_rn(sx, q.c), __fmul_rn(sy, q.s));
  float ry = __fadd_rn(__fmul_rn(sx, q.s), __fmul_rn(sy, q.c));
  float cx = b[0] + rx, cy = b[1] + ry, cz = b[2] + sz;
  float d0 = cx - b[0] + b[3] / 2, d1 = b[0] + b[3] / 2 - cx;
  float d2 = cy - b[1] + b[4] / 2, d3 = b[1] + b[4] / 2 - cy;
  float d4 = cz - b[2] + b[5] / 2, d5 = b[2] + b[5] / 2 - cz;
  return fminf(fminf(fminf(d0, d1), fminf(d2, d3)), fminf(d4, d5)) > 0.f;
}

constexpr int kTrimBoxGroup = 32;
// phase A: grid (point chunks, box groups)
__global__ void __launch_bounds__(256) trim_vote_kernel(const float* __restrict__ pts, int ld_pts, const int64_t* __restrict__ sp,
                                                        int n_pts, int n_sp, const float* __restrict__ boxes, int box_dim,
                                                        const int32_t* __restrict__ box_index, int m,
                                                        const int32_t* __restrict__ m_dev, int* __restrict__ cnt) {
  const int mm = m_dev ? min(m, *m_dev) : m;
  const int b0 = blockIdx.y * kTrimBoxGroup;
  if (b0 >= mm) return;
  const int nb = min(kTrimBoxGroup, mm - b0);
  __shared__ BoxP sbox[kTrimBoxGroup];
  if ((int)threadIdx.x < nb) sbox[threadIdx.x] = load_box(boxes, box_dim, box_index, b0 + threadIdx.x);
  __syncthreads();
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_pts; p += gridDim.x * blockDim.x) {
    const float* q = pts + (size_t)p * ld_pts;
    const float x = q[0], y = q[1], z = q[2];
    const long long id = sp[p];
    if (id < 0 || id >= n_sp) continue;
    for (int j = 0; j < nb; ++j)
      if (point_in_box(x, y, z, sbox[j])) atomicAdd(cnt + (size_t)(b0 + j) * n_sp + id, 1);
  }
}

__device__ __forceinline__ int f2ord(float f) {   // monotone float -> int
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void trim_init_kernel(int* aabb, int m) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m * 6) aabb[i] = (i % 6) < 3 ? f2ord(INFINITY) : f2ord(-INFINITY);
}

// phase B: one warp per (box, superpoint) entry of cnt
__global__ void __launch_bounds__(256) trim_decide_kernel(const float* __restrict__ pts, int ld_pts, int n_sp,
                                                          const float* __restrict__ boxes, int box_dim,
                                                          const int32_t* __restrict__ box_index, int m,
                                                          const int32_t* __restrict__ m_dev, const int* __restrict__ cnt,
                                                          const int* __restrict__ sp_size, const int* __restrict__ sp_start,
                                                          const int* __restrict__ sp_points, float low_thr, float up_thr,
                                                          int* aabb) {
  const int mm = m_dev ? min(m, *m_dev) : m;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)mm * n_sp;
  for (long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < total;
       e += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int c = cnt[e];
    if (c == 0) continue;                       // frac = 0 < low_thr: no point of this superpoint is selected
    const int b = (int)(e / n_sp), s = (int)(e % n_sp);
    const float frac = __fdiv_rn((float)c, fmaxf((float)sp_size[s], 1.f));
    const bool del = frac < low_thr, add = frac > up_thr;
    if (del && !add) continue;
    const BoxP q = load_box(boxes, box_dim, box_index, b);
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    const int beg = sp_start[s], end = sp_start[s + 1];
    for (int i = beg + lane; i < end; i += 32) {
      const float* pp = pts + (size_t)sp_points[i] * ld_pts;
      const float x = pp[0], y = pp[1], z = pp[2];
      bool in = add || point_in_box(x, y, z, q);
      if (in) {
        mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
        mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
      }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      for (int o = 16; o > 0; o >>= 1) {
        mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
        mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
      }
    }
    if (lane < 3) atomicMin(aabb + b * 6 + lane, f2ord(mn[lane]));
    else if (lane < 6) atomicMax(aabb + b * 6 + lane, f2ord(mx[lane - 3]));
  }
}

__global__ void trim_final_kernel(const int* __restrict__ aabb, int m, const int32_t* __restrict__ m_dev, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int mm = m_dev ? min(m, *m_dev) : m;
  if (i >= mm * 3) return;
  int b = i / 3, d = i % 3;
  float a = ord2f(aabb[b * 6 + d]), z = ord2f(aabb[b * 6 + 3 + d]);
  out[(size_t)b * 6 + d] = (z + a) / 2.f;
  out[(size_t)b * 6 + 3 + d] = z - a;
}

__global__ void gather_rows_kernel(const float* __restrict__ src, int dim, const int32_t* __restrict__ idx, int n,
                                   float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * dim) return;
  int r = i / dim, c = i % dim;
  out[i] = src[(size_t)idx[r] * dim + c];
}

}  // namespace ud3d

using namespace ud3d;

extern "C" {

int ud3d_topk_scores(const float* logits, int T, int C_plus1, int k, float* scores_out, int32_t* labels_out,
                     int32_t* query_out, void* ws, size_t ws_bytes, void* stream) {
  UD3D_CHECK_ARG(logits && scores_out && labels_out && query_out && ws, "ud3d_topk_scores: NULL argument");
  const int C = C_plus1 - 1;
  UD3D_CHECK_ARG(T > 0 && C > 0 && k > 0 && k <= 1024, "ud3d_topk_scores: need T > 0, C > 0, 0 < k <= 1024");
  UD3D_CHECK_ARG((long long)T * C >= k, "ud3d_topk_scores: k=%d out of range for %lld scores (torch.topk raises here too)", k,
                 (long long)T * C);
  UD3D_CHECK_ARG((long long)T * C < (1ll << 31), "ud3d_topk_scores: too many scores");
  if (ws_bytes < (size_t)T * C * 4) {
    set_error("ud3d_topk_scores: workspace too small");
    return UD3D_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  softmax_scores_kernel<<<cdiv(T, 128), 128, 0, st>>>(logits, C_plus1, T, C_plus1, (float*)ws);
  UD3D_LAUNCH_CHECK();
  topk_select_kernel<<<1, 1024, 0, st>>>((const float*)ws, T * C, C, k, scores_out, labels_out, query_out);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

size_t ud3d_nms_workspace_bytes(int n) { return nms_ws_bytes(n > 0 ? n : 1); }

int ud3d_nms_multiclass(const float* boxes, int box_dim, const float* scores, const int32_t* labels, int n, int mode,
                        float iou_thr, float score_thr, int32_t* keep_out, int32_t* n_keep, void* ws, size_t ws_bytes,
                        void* stream) {
  UD3D_CHECK_ARG(boxes && scores && labels && keep_out && n_keep && ws, "ud3d_nms_multiclass: NULL argument");
  UD3D_CHECK_ARG(n >= 0 && n <= 1024 && (box_dim == 6 || box_dim == 7) && mode >= 0 && mode <= 2,
                 "ud3d_nms_multiclass: need n <= 1024, box_dim 6|7, mode 0..2");
  if (ws_bytes < nms_ws_bytes(n > 0 ? n : 1)) {
    set_error("ud3d_nms_multiclass: workspace too small");
    return UD3D_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    UD3D_CUDA(cudaMemsetAsync(n_keep, 0, 4, st));
    return UD3D_OK;
  }
  NmsWs w = nms_ws_view(ws, n);
  nms_order_kernel<<<1, 1024, 0, st>>>(boxes, box_dim, scores, labels, n, score_thr, w);
  UD3D_LAUNCH_CHECK();
  int nb = cdiv(n, 64);
  nms_mask_kernel<<<dim3(nb, nb), 64, 0, st>>>(w, n, mode, iou_thr);
  UD3D_LAUNCH_CHECK();
  size_t smem = (size_t)n * 16 * 8;
  DeviceCtx* ctx = device_ctx();
  if (!ctx) return UD3D_ECUDA;
  {
    CtxGuard guard(ctx);
    if (ctx_needs_config(ctx, (const void*)nms_sweep_kernel, 1024 * 16 * 8))
      UD3D_CUDA(cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 * 16 * 8));
  }
  nms_sweep_kernel<<<1, 1024, smem, st>>>(w, keep_out, n_keep);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

size_t ud3d_trim_workspace_bytes(int n_sp, int n_pts, int m) {
  return trim_ws_bytes(n_sp > 0 ? n_sp : 1, n_pts > 0 ? n_pts : 1, m > 0 ? m : 1);
}

int ud3d_trim_boxes(const float* points, int ld_pts, const int64_t* sp, int n_pts, int n_sp, const float* boxes,
                    int box_dim, const int32_t* box_index, int m, const int32_t* m_dev, float low_thr, float up_thr,
                    float* out, void* ws, size_t ws_bytes, void* stream) {
  UD3D_CHECK_ARG(points && sp && boxes && out && ws, "ud3d_trim_boxes: NULL argument");
  UD3D_CHECK_ARG(ld_pts >= 3 && n_pts >= 0 && n_sp > 0 && (box_dim == 6 || box_dim == 7) && m >= 0, "ud3d_trim_boxes: bad sizes");
  UD3D_CHECK_ARG((long long)m * n_sp < (1ll << 31), "ud3d_trim_boxes: m * n_sp too large");
  if (ws_bytes < ud3d_trim_workspace_bytes(n_sp, n_pts, m)) {
    set_error("ud3d_trim_boxes: workspace too small");
    return UD3D_EWORKSPACE;
  }
  if (m == 0) return UD3D_OK;
  cudaStream_t st = (cudaStream_t)stream;
  TrimWs w = trim_ws_view(ws, n_sp, n_pts > 0 ? n_pts : 1, m);
  UD3D_CUDA(cudaMemsetAsync(w.sp_size, 0, (size_t)n_sp * 4, st));
  UD3D_CUDA(cudaMemsetAsync(w.cnt, 0, (size_t)m * n_sp * 4, st));
  trim_init_kernel<<<cdiv(m * 6, 256), 256, 0, st>>>(w.aabb, m);
  UD3D_LAUNCH_CHECK();
  if (n_pts > 0) {
    int pblocks = cdiv(n_pts, 256);
    if (pblocks > 148 * 8) pblocks = 148 * 8;
    sp_size_kernel<<<pblocks, 256, 0, st>>>(sp, n_pts, n_sp, w.sp_size);
    UD3D_LAUNCH_CHECK();
    sp_scan_kernel<<<1, 1024, 0, st>>>(w.sp_size, n_sp, w.sp_start, w.cursor);
    UD3D_LAUNCH_CHECK();
    sp_fill_kernel<<<pblocks, 256, 0, st>>>(sp, n_pts, n_sp, w.cursor, w.sp_points);
    UD3D_LAUNCH_CHECK();
    int vb = cdiv(n_pts, 256);
    if (vb > 296) vb = 296;
    trim_vote_kernel<<<dim3(vb, cdiv(m, kTrimBoxGroup)), 256, 0, st>>>(points, ld_pts, sp, n_pts, n_sp, boxes, box_dim, box_index, m,
                                                                   m_dev, w.cnt);
    UD3D_LAUNCH_CHECK();
    long long warps = (long long)m * n_sp;
    long long db = (warps * 32 + 255) / 256;
    if (db > 148 * 16) db = 148 * 16;
    trim_decide_kernel<<<(int)db, 256, 0, st>>>(points, ld_pts, n_sp, boxes, box_dim, box_index, m, m_dev, w.cnt, w.sp_size,
                                                w.sp_start, w.sp_points, low_thr, up_thr, w.aabb);
    UD3D_LAUNCH_CHECK();
  }
  trim_final_kernel<<<cdiv(m * 3, 256), 256, 0, st>>>(w.aabb, m, m_dev, out);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

size_t ud3d_postprocess_workspace_bytes(const ud3d_post_args* a) {
  if (!a) return 0;
  size_t b = align_up((size_t)(a->T > 0 ? a->T : 1) * (a->C1 > 1 ? a->C1 - 1 : 1) * 4, 256);   // scores
  b += align_up((size_t)a->k * 4, 256);                                                        // query
  b += align_up(ud3d_nms_workspace_bytes(a->k), 256);
  if (a->use_trim) b += align_up(ud3d_trim_workspace_bytes(a->n_sp, a->n_pts, a->k), 256);
  return b;
}

int ud3d_postprocess_scene(const ud3d_post_args* a, void* ws, size_t ws_bytes, void* stream) {
  UD3D_CHECK_ARG(a && ws, "ud3d_postprocess_scene: NULL argument");
  UD3D_CHECK_ARG(a->logits && a->boxes && a->scores && a->labels && a->cand && a->keep && a->n_keep, "ud3d_postprocess_scene: NULL buffer");
  UD3D_CHECK_ARG(!a->use_trim || (a->points && a->sp && a->trimmed), "ud3d_postprocess_scene: trim needs points / sp / trimmed");
  UD3D_CHECK_ARG(a->ld_logits >= a->C1 && a->T > 0 && a->C1 > 1 && a->k > 0 && a->k <= 1024, "ud3d_postprocess_scene: bad sizes");
  UD3D_CHECK_ARG((long long)a->T * (a->C1 - 1) >= a->k, "ud3d_postprocess_scene: k=%d out of range for %lld scores", a->k,
                 (long long)a->T * (a->C1 - 1));
  if (ws_bytes < ud3d_postprocess_workspace_bytes(a)) {
    set_error("ud3d_postprocess_scene: workspace too small");
    return UD3D_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* p = (char*)ws;
  float* scores_all = (float*)p;
  p += align_up((size_t)a->T * (a->C1 - 1) * 4, 256);
  int32_t* query = (int32_t*)p;
  p += align_up((size_t)a->k * 4, 256);
  void* nms_ws = p;
  size_t nms_bytes = align_up(ud3d_nms_workspace_bytes(a->k), 256);
  p += nms_bytes;
  const int C = a->C1 - 1;
  softmax_scores_kernel<<<cdiv(a->T, 128), 128, 0, st>>>(a->logits, a->ld_logits, a->T, a->C1, scores_all);
  UD3D_LAUNCH_CHECK();
  topk_select_kernel<<<1, 1024, 0, st>>>(scores_all, a->T * C, C, a->k, a->scores, a->labels, query);
  UD3D_LAUNCH_CHECK();
  gather_rows_kernel<<<cdiv(a->k * a->box_dim, 256), 256, 0, st>>>(a->boxes, a->box_dim, query, a->k, a->cand);
  UD3D_LAUNCH_CHECK();
  int rc = ud3d_nms_multiclass(a->cand, a->box_dim, a->scores, a->labels, a->k, a->nms_mode, a->iou_thr, a->score_thr, a->keep,
                               a->n_keep, nms_ws, nms_bytes, stream);
  if (rc) return rc;
  if (a->use_trim) {
    size_t tb = ud3d_trim_workspace_bytes(a->n_sp, a->n_pts, a->k);
    rc = ud3d_trim_boxes(a->points, a->ld_pts, a->sp, a->n_pts, a->n_sp, a->cand, a->box_dim, a->keep, a->k, a->n_keep, a->low_thr,
                         a->up_thr, a->trimmed, p, tb, stream);
    if (rc) return rc;
  }
  return UD3D_OK;
}

}  // extern "C"
