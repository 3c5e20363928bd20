// Training-side kernels of the backbone (SURVEY.md section 8a R6 second half, 8f rank 2):
//   * batch statistics of train-mode (Sync)BatchNorm over all active voxels of the batch (spconv_unet.py:119-124,
//     unidet3d.py:104-107): deterministic two-stage per-channel sum / sum of squares in fp64; the cross-rank reduction
//     of SyncBatchNorm is one all-reduce of the [2C] sums between the two calls (host side, torch.distributed / NCCL);
//   * the fold into the (scale, shift) form every conv consumes, with the running-statistics update;
//   * weight gradient of a sparse convolution: dW[co, k, ci] = sum_o dY[o, co] * X[table[k][o], ci].
#include "common.cuh"

#include <stdlib.h>

namespace ud3d {

constexpr int kBnRowsPerBlock = 512;

// block b: rows [b * 512, ...).  C >= 256: thread t takes channels t, t + 256, ...  C < 256: the 256 threads form
// RG = 256 / C row groups (thread t: channel t % C, rows r0 + t / C, + RG, ...), so that narrow maps (32 channels at the
// finest level, the largest by rows) still use every thread and every row is read as full 128-byte lines; the groups'
// sums are combined through shared memory in group order (fixed order: deterministic).
__device__ __forceinline__ void bn_block_combine(double s, double q, int C, int RG, int c, int rg, int block, double* __restrict__ part) {
  __shared__ double sh[2][256];
  if (RG > 1) {
    sh[0][threadIdx.x] = s;
    sh[1][threadIdx.x] = q;
    __syncthreads();
    if (rg == 0 && c < C) {
      for (int g = 1; g < RG; ++g) {
        s += sh[0][g * C + c];
        q += sh[1][g * C + c];
      }
    }
  }
  if (rg == 0 && c < C) {
    part[((size_t)block * 2 + 0) * C + c] = s;
    part[((size_t)block * 2 + 1) * C + c] = q;
  }
}
__global__ void __launch_bounds__(256) bn_partial_sums_kernel(const float* __restrict__ x, int ld, int n, int C,
                                                              double* __restrict__ part) {
  const int r0 = blockIdx.x * kBnRowsPerBlock;
  const int r1 = min(n, r0 + kBnRowsPerBlock);
  if (C >= 256) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      double s = 0.0, q = 0.0;
      for (int r = r0; r < r1; ++r) {
        const double v = (double)x[(size_t)r * ld + c];
        s += v;
        q += v * v;
      }
      part[((size_t)blockIdx.x * 2 + 0) * C + c] = s;
      part[((size_t)blockIdx.x * 2 + 1) * C + c] = q;
    }
    return;
  }
  const int RG = 256 / C;
  const int rg = threadIdx.x / C, c = threadIdx.x - rg * C;
  double s = 0.0, q = 0.0;
  if (rg < RG) {
#pragma unroll 4
    for (int r = r0 + rg; r < r1; r += RG) {
      const double v = (double)x[(size_t)r * ld + c];
      s += v;
      q += v * v;
    }
  }
  bn_block_combine(s, q, C, RG, c, rg, blockIdx.x, part);
}
// fixed-order sum of the partials: sums[0][c] = sum x, sums[1][c] = sum x^2
__global__ void bn_final_sums_kernel(const double* __restrict__ part, int nblocks, int C, double* __restrict__ sums) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * C) return;
  const int which = i / C, c = i - which * C;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += part[((size_t)b * 2 + which) * C + c];
  sums[i] = s;
}
// torch.nn.(Sync)BatchNorm in training mode: biased variance normalises, unbiased variance updates running_var
__global__ void bn_train_fold_kernel(const double* __restrict__ sums, double count_host, const double* __restrict__ count_dev, int C, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float eps, float momentum, float* running_mean,
                                     float* running_var, float* __restrict__ scale, float* __restrict__ shift,
                                     float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double count = count_dev ? *count_dev : count_host;
  const double mean = sums[c] / count;
  double var = sums[C + c] / count - mean * mean;
  var = var > 0.0 ? var : 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale[c] = g * invstd;
  shift[c] = b - (float)mean * g * invstd;
  if (save_mean) save_mean[c] = (float)mean;
  if (save_invstd) save_invstd[c] = invstd;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
  if (running_var) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// ---------------------------------------------------------------- backward of the fused BatchNorm(train) + ReLU
// Forward (folded into the consumer conv's operand load): a = relu(x * scale + shift), scale = gamma * invstd,
// shift = beta - mean * scale.  Given dA (the conv's input gradient):
//   g      = dA * [x * scale + shift > 0]                       (relu == 0: g = dA)
//   dbeta  = sum_r g,   dgamma = sum_r g * xhat,   xhat = (x - mean) * invstd
//   dx     = scale * (g - dbeta / count - xhat * dgamma / count)       (batch statistics depend on every row)
// Two passes like the forward statistics: deterministic fp64 partial sums (all-reduced across ranks for SyncBatchNorm,
// the "bwd all_reduce of [2C]" of SURVEY.md 8e), then the element-wise pass.
__global__ void __launch_bounds__(256) bn_bwd_partial_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ da,
                                                             int ld_da, int n, int C, const float* __restrict__ scale,
                                                             const float* __restrict__ shift, const float* __restrict__ mean,
                                                             const float* __restrict__ invstd, int relu, double* __restrict__ part) {
  const int r0 = blockIdx.x * kBnRowsPerBlock;
  const int r1 = min(n, r0 + kBnRowsPerBlock);
  if (C >= 256) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float sc = scale[c], sh = shift[c], mu = mean[c], is = invstd[c];
      double s = 0.0, q = 0.0;
      for (int r = r0; r < r1; ++r) {
        const float xv = x[(size_t)r * ld_x + c];
        float g = da[(size_t)r * ld_da + c];
        if (relu && !(fmaf(xv, sc, sh) > 0.f)) g = 0.f;
        s += (double)g;
        q += (double)g * (double)((xv - mu) * is);
      }
      part[((size_t)blockIdx.x * 2 + 0) * C + c] = s;
      part[((size_t)blockIdx.x * 2 + 1) * C + c] = q;
    }
    return;
  }
  const int RG = 256 / C;                         // row groups, see bn_partial_sums_kernel
  const int rg = threadIdx.x / C, c = threadIdx.x - rg * C;
  double s = 0.0, q = 0.0;
  if (rg < RG) {
    const float sc = scale[c], sh = shift[c], mu = mean[c], is = invstd[c];
#pragma unroll 4
    for (int r = r0 + rg; r < r1; r += RG) {
      const float xv = x[(size_t)r * ld_x + c];
      float g = da[(size_t)r * ld_da + c];
      if (relu && !(fmaf(xv, sc, sh) > 0.f)) g = 0.f;
      s += (double)g;
      q += (double)g * (double)((xv - mu) * is);
    }
  }
  bn_block_combine(s, q, C, RG, c, rg, blockIdx.x, part);
}
__global__ void bn_bwd_apply_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ da, int ld_da, int n, int C,
                                    const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                                    const float* __restrict__ invstd, int relu, const double* __restrict__ sums, double count_host,
                                    const double* __restrict__ count_dev,
                                    float* __restrict__ dx, int ld_dx, int accumulate) {
  const double count = count_dev ? fmax(*count_dev, 1.0) : count_host;
  const long long total = (long long)n * C;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(t / C), c = (int)(t - (long long)r * C);
    const float xv = x[(size_t)r * ld_x + c];
    float g = da[(size_t)r * ld_da + c];
    if (relu && !(fmaf(xv, scale[c], shift[c]) > 0.f)) g = 0.f;
    const float xhat = (xv - mean[c]) * invstd[c];
    const float v = scale[c] * (g - (float)(sums[c] / count) - xhat * (float)(sums[C + c] / count));
    float* o = dx + (size_t)r * ld_dx + c;
    *o = accumulate ? *o + v : v;
  }
}
// a = relu?(x * scale + shift) as a plain fp32 map (the X operand of ud3d_conv_wgrad)
__global__ void bn_relu_apply_kernel(const float* __restrict__ x, int ld_x, int n, int C, const float* __restrict__ scale,
                                     const float* __restrict__ shift, int relu, float* __restrict__ out, int ld_out) {
  const long long total = (long long)n * C;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(t / C), c = (int)(t - (long long)r * C);
    float v = fmaf(x[(size_t)r * ld_x + c], scale[c], shift[c]);
    if (relu) v = fmaxf(v, 0.f);
    out[(size_t)r * ld_out + c] = v;
  }
}
// Backward of the superpoint mean-pool (unidet3d.py:130): d_vox[inverse[p], :] += d_pooled[seg[p], :] / count[seg[p]].
// One warp per point row segment; 64-bit fixed-point atomics (2^-32 resolution of a gradient) keep the scatter
// deterministic; a second pass converts.
__global__ void pool_bwd_scatter_kernel(const float* __restrict__ d_pooled, int C, const int32_t* __restrict__ gather,
                                        const int64_t* __restrict__ seg, const int32_t* __restrict__ cnt, int n, int n_seg,
                                        unsigned long long* __restrict__ acc) {
  const long long total = (long long)n * C;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(t / C), c = (int)(t - (long long)p * C);
    const long long s = seg[p];
    if (s < 0 || s >= n_seg) continue;
    const int v = gather ? gather[p] : p;
    if (v < 0) continue;
    const int k = cnt[s] > 1 ? cnt[s] : 1;
    const double g = (double)d_pooled[(size_t)s * C + c] / (double)k;
    atomicAdd(acc + (size_t)v * C + c, (unsigned long long)__double2ll_rn(g * 4294967296.0));
  }
}
__global__ void seg_count_kernel(const int64_t* __restrict__ seg, int n, int n_seg, int32_t* __restrict__ cnt) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const long long s = seg[p];
    if (s >= 0 && s < n_seg) atomicAdd(cnt + s, 1);
  }
}
__global__ void fix32_to_float_kernel(const unsigned long long* __restrict__ acc, long long total, float* __restrict__ out) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x)
    out[t] = (float)((double)(long long)acc[t] * (1.0 / 4294967296.0));
}

// ---------------------------------------------------------------- encoder-side backward pieces
// LayerNorm over the last dimension (encoder.py:38-39,77-78,189), y = (x - mu) * rstd * gamma + beta:
//   dx = rstd * (gy - mean_c(gy) - xhat * mean_c(gy * xhat)),  gy = dy * gamma;   dgamma = sum_r dy * xhat,  dbeta = sum_r dy.
// One warp per row (statistics recomputed from x); the per-channel sums are written as per-block partials and reduced by
// bn_final_sums_kernel (fixed order: deterministic).
constexpr int kLnRowsPerBlock = 64;      // 8 warps x 8 rows
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                            const float* __restrict__ gamma, int rows, int C, float eps,
                                                            float* __restrict__ dx, double* __restrict__ part) {
  extern __shared__ float s_acc[];                 // [8 warps][2][C]: every warp sums its own rows (no atomics: deterministic)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = threadIdx.x; c < 16 * C; c += blockDim.x) s_acc[c] = 0.f;
  __syncthreads();
  float* mine = s_acc + (size_t)warp * 2 * C;
  const int r0 = blockIdx.x * kLnRowsPerBlock;
  for (int rr = warp; rr < kLnRowsPerBlock; rr += 8) {
    const int r = r0 + rr;
    if (r >= rows) break;
    const float* xr = x + (size_t)r * C;
    const float* gr = dy + (size_t)r * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xr[c];
    const float mu = warp_sum(s) / (float)C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = xr[c] - mu; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    float a = 0.f, b = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float gy = gr[c] * gamma[c], xh = (xr[c] - mu) * rstd;
      a += gy;
      b = fmaf(gy, xh, b);
    }
    a = warp_sum(a) / (float)C;
    b = warp_sum(b) / (float)C;
    for (int c = lane; c < C; c += 32) {
      const float xh = (xr[c] - mu) * rstd;
      dx[(size_t)r * C + c] = rstd * (gr[c] * gamma[c] - a - xh * b);
      mine[c] += gr[c] * xh;
      mine[C + c] += gr[c];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += (double)s_acc[(size_t)w * 2 * C + c];
    part[(size_t)blockIdx.x * 2 * C + c] = t;
  }
}
// dX = dY * act'(pre):  act 1 = relu (pre > 0), 2 = gelu(erf): 0.5 (1 + erf(z / sqrt2)) + z exp(-z^2 / 2) / sqrt(2 pi)
__global__ void act_bwd_kernel(const float* __restrict__ pre, const float* __restrict__ dy, long long total, int act,
                               float* __restrict__ dx) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const float z = pre[t];
    float d;
    if (act == 1) d = z > 0.f ? 1.f : 0.f;
    else d = 0.5f * (1.f + erff(z * 0.70710678118654752440f)) + z * expf(-0.5f * z * z) * 0.39894228040143267794f;
    dx[t] = dy[t] * d;
  }
}

__global__ void act_fwd_kernel(const float* __restrict__ pre, long long total, int act, float* __restrict__ out) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const float z = pre[t];
    out[t] = act == 1 ? fmaxf(z, 0.f) : 0.5f * z * (1.f + erff(z * 0.70710678118654752440f));
  }
}

// ---------------------------------------------------------------- sparse-conv weight gradient
// dW[co][k][ci] = sum_o dY[o][co] * X[table[k][o]][ci]   (X = the conv's input AFTER its BatchNorm+ReLU).
// One CTA = (kernel offset k, 32 x 32 tile of (co, ci), slice of the output rows); rows are staged through shared memory
// 64 at a time; partial tiles of the row slices are summed in a fixed order by a second kernel (deterministic).
constexpr int kWgRows = 64;
__global__ void __launch_bounds__(256) conv_wgrad_partial_kernel(const float* __restrict__ x, int ld_x, int c_in,
                                                                 const float* __restrict__ dy, int ld_dy, int c_out,
                                                                 const int32_t* __restrict__ table, int n_out, int K,
                                                                 int rows_per_slice, float* __restrict__ part) {
  __shared__ float sx[kWgRows][33];
  __shared__ float sy[kWgRows][33];
  const int k = blockIdx.y;
  const int tiles_ci = (c_in + 31) / 32;
  const int tco = blockIdx.x / tiles_ci, tci = blockIdx.x - tco * tiles_ci;
  const int slice = blockIdx.z;
  const int r_begin = slice * rows_per_slice, r_end = min(n_out, r_begin + rows_per_slice);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // thread -> (ci = tx, co = ty, ty + 8, ty + 16, ty + 24)
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r0 = r_begin; r0 < r_end; r0 += kWgRows) {
    for (int i = threadIdx.x; i < kWgRows * 32; i += 256) {
      const int rr = i >> 5, cc = i & 31;
      const int o = r0 + rr;
      float vx = 0.f, vy = 0.f;
      if (o < r_end) {
        const int src = table ? __ldg(table + (size_t)k * n_out + o) : o;
        if (src >= 0) {
          if (tci * 32 + cc < c_in) vx = x[(size_t)src * ld_x + tci * 32 + cc];
          if (tco * 32 + cc < c_out) vy = dy[(size_t)o * ld_dy + tco * 32 + cc];
        }
      }
      sx[rr][cc] = vx;
      sy[rr][cc] = vy;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < kWgRows; ++rr) {
      const float xv = sx[rr][tx];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(sy[rr][ty + 8 * j], xv, acc[j]);
    }
    __syncthreads();
  }
  // part[slice][co][k][ci]
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int co = tco * 32 + ty + 8 * j, ci = tci * 32 + tx;
    if (co < c_out && ci < c_in) part[(((size_t)slice * c_out + co) * K + k) * c_in + ci] = acc[j];
  }
}
// Register-tiled version (the product path): a CTA owns a TM x TN tile of (co, ci) for one kernel offset and row slice;
// every thread accumulates a 4 x 4 block of it, reading its 4 dY values and 4 X values of a staged row as two 16-byte
// shared-memory loads (2 LDS.128 per 16 FMAs instead of 5 LDS.32 per 4 FMAs above: the FMA pipe is the limit, not the
// shared-memory port).  Tiles smaller than 64 x 64 leave threads over: the 256 threads then form G = 2 or 4 groups that
// take alternate rows of the 32 staged ones and combine their blocks through shared memory in a fixed order at the end.
constexpr int kWg2Rows = 32;
template <int TM, int TN>
__global__ void __launch_bounds__(256) conv_wgrad_tiled_kernel(const float* __restrict__ x, int ld_x, int c_in,
                                                               const float* __restrict__ dy, int ld_dy, int c_out,
                                                               const int32_t* __restrict__ table, int n_out, int K,
                                                               int rows_per_slice, int vec_x, int vec_y, float* __restrict__ part) {
  constexpr int TPG = (TM / 4) * (TN / 4);       // threads per group
  constexpr int G = 256 / TPG;                   // row-interleaved groups
  static_assert(TPG * G == 256 && (TM % 4) == 0 && (TN % 4) == 0, "tile shape");
  constexpr int kRedFloats = (G > 1) ? (G - 1) * TM * TN : 1;
  constexpr int kStageFloats = kWg2Rows * (TM + TN);
  __shared__ __align__(16) float smem[kStageFloats > kRedFloats ? kStageFloats : kRedFloats];
  float4(*sx)[TN / 4] = reinterpret_cast<float4(*)[TN / 4]>(smem);
  float4(*sy)[TM / 4] = reinterpret_cast<float4(*)[TM / 4]>(smem + kWg2Rows * TN);
  const int k = blockIdx.y;
  const int tiles_ci = (c_in + TN - 1) / TN;
  const int tco = blockIdx.x / tiles_ci, tci = blockIdx.x - tco * tiles_ci;
  const int slice = blockIdx.z;
  const int r_begin = slice * rows_per_slice, r_end = min(n_out, r_begin + rows_per_slice);
  const int grp = threadIdx.x / TPG, tig = threadIdx.x - grp * TPG;
  const int tx = tig % (TN / 4), ty = tig / (TN / 4);        // ci block tx, co block ty
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const int ci0 = tci * TN, co0 = tco * TM;
  for (int r0 = r_begin; r0 < r_end; r0 += kWg2Rows) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < kWg2Rows * (TN / 4); idx += 256) {
      const int rr = idx / (TN / 4), c4 = idx - rr * (TN / 4);
      const int o = r0 + rr;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (o < r_end) {
        const int src = table ? __ldg(table + (size_t)k * n_out + o) : o;
        const int c = ci0 + c4 * 4;
        if (src >= 0 && c < c_in) {
          const float* px = x + (size_t)src * ld_x + c;
          if (vec_x && c + 3 < c_in) {
            v = __ldg((const float4*)px);
          } else {
            v.x = px[0];
            if (c + 1 < c_in) v.y = px[1];
            if (c + 2 < c_in) v.z = px[2];
            if (c + 3 < c_in) v.w = px[3];
          }
        }
      }
      sx[rr][c4] = v;
    }
    for (int idx = threadIdx.x; idx < kWg2Rows * (TM / 4); idx += 256) {
      const int rr = idx / (TM / 4), c4 = idx - rr * (TM / 4);
      const int o = r0 + rr;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int c = co0 + c4 * 4;
      if (o < r_end && c < c_out) {
        const float* py = dy + (size_t)o * ld_dy + c;
        if (vec_y && c + 3 < c_out) {
          v = __ldg((const float4*)py);
        } else {
          v.x = py[0];
          if (c + 1 < c_out) v.y = py[1];
          if (c + 2 < c_out) v.z = py[2];
          if (c + 3 < c_out) v.w = py[3];
        }
      }
      sy[rr][c4] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int rr = grp; rr < kWg2Rows; rr += G) {
      const float4 xv = sx[rr][tx];
      const float4 yv = sy[rr][ty];
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ys[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(ys[a], xs[b], acc[a][b]);
    }
  }
  if (G > 1) {       // groups 1 .. G-1 park their blocks in shared memory, group 0 adds them in order
    __syncthreads();
    if (grp > 0) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) smem[((size_t)(grp - 1) * 16 + a * 4 + b) * TPG + tig] = acc[a][b];
    }
    __syncthreads();
    if (grp == 0) {
      for (int g2 = 1; g2 < G; ++g2)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] += smem[((size_t)(g2 - 1) * 16 + a * 4 + b) * TPG + tig];
    }
  }
  if (grp == 0) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int co = co0 + ty * 4 + a;
      if (co >= c_out) continue;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int ci = ci0 + tx * 4 + b;
        if (ci < c_in) part[(((size_t)slice * c_out + co) * K + k) * c_in + ci] = acc[a][b];      // part[slice][co][k][ci]
      }
    }
  }
}

template <int TM, int TN>
static void launch_wgrad_tiled(const float* x, int ld_x, int c_in, const float* dy, int ld_dy, int c_out, const int32_t* table, int n_out,
                               int K, int rps, int slices, float* part, cudaStream_t st) {
  const int vec_x = (ld_x % 4 == 0) && (((uintptr_t)x & 15) == 0);
  const int vec_y = (ld_dy % 4 == 0) && (((uintptr_t)dy & 15) == 0);
  dim3 grid(cdiv(c_out, TM) * cdiv(c_in, TN), K, slices);
  conv_wgrad_tiled_kernel<TM, TN><<<grid, 256, 0, st>>>(x, ld_x, c_in, dy, ld_dy, c_out, table, n_out, K, rps, vec_x, vec_y, part);
}

__global__ void conv_wgrad_reduce_kernel(const float* __restrict__ part, int n_slices, size_t n_w, float* __restrict__ dw,
                                         int accumulate) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_w) return;
  float s = 0.f;
  for (int sl = 0; sl < n_slices; ++sl) s += part[(size_t)sl * n_w + i];
  dw[i] = accumulate ? dw[i] + s : s;
}

static int wgrad_slices(int n_out) {
  int s = cdiv(n_out, 4096);
  if (s > 64) s = 64;
  return s < 1 ? 1 : s;
}

}  // namespace ud3d

using namespace ud3d;

extern "C" {

size_t ud3d_bn_batch_sums_workspace_bytes(int n, int C) {
  if (n < 0 || C <= 0) return 0;
  return (size_t)cdiv(n > 0 ? n : 1, kBnRowsPerBlock) * 2 * C * sizeof(double);
}

int ud3d_bn_batch_sums(const float* x, int ld, int n, int C, double* sums, void* ws, size_t ws_bytes, void* stream) {
  UD3D_CHECK_ARG(x && sums && ws && C > 0 && n >= 0 && ld >= C, "ud3d_bn_batch_sums: bad argument");
  if (ws_bytes < ud3d_bn_batch_sums_workspace_bytes(n, C)) {
    set_error("ud3d_bn_batch_sums: workspace too small");
    return UD3D_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = n > 0 ? cdiv(n, kBnRowsPerBlock) : 0;
  if (nb > 0) {
    bn_partial_sums_kernel<<<nb, 256, 0, st>>>(x, ld, n, C, (double*)ws);
    UD3D_LAUNCH_CHECK();
  }
  bn_final_sums_kernel<<<cdiv(2 * C, 128), 128, 0, st>>>((const double*)ws, nb, C, sums);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int ud3d_bn_train_fold(const double* sums, double count, int C, const float* gamma, const float* beta, float eps,
                       float momentum, float* running_mean, float* running_var, float* scale, float* shift,
                       float* save_mean, float* save_invstd, const double* count_dev, void* stream) {
  UD3D_CHECK_ARG(sums && scale && shift && C > 0 && (count_dev || count > 0.0), "ud3d_bn_train_fold: bad argument");
  bn_train_fold_kernel<<<cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(sums, count, count_dev, C, gamma, beta, eps, momentum, running_mean,
                                                                     running_var, scale, shift, save_mean, save_invstd);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int ud3d_bn_backward_sums(const float* x, int ld_x, const float* da, int ld_da, int n, int C, const float* scale, const float* shift,
                          const float* mean, const float* invstd, int relu, double* sums, void* ws, size_t ws_bytes, void* stream) {
  UD3D_CHECK_ARG(x && da && scale && shift && mean && invstd && sums && ws && C > 0 && n >= 0 && ld_x >= C && ld_da >= C,
                 "ud3d_bn_backward_sums: bad argument");
  if (ws_bytes < ud3d_bn_batch_sums_workspace_bytes(n, C)) {
    set_error("ud3d_bn_backward_sums: workspace too small");
    return UD3D_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int nblocks = cdiv(n > 0 ? n : 1, kBnRowsPerBlock);
  if (n == 0) {
    UD3D_CUDA(cudaMemsetAsync(sums, 0, (size_t)2 * C * sizeof(double), st));
    return UD3D_OK;
  }
  bn_bwd_partial_kernel<<<nblocks, 256, 0, st>>>(x, ld_x, da, ld_da, n, C, scale, shift, mean, invstd, relu, (double*)ws);
  UD3D_LAUNCH_CHECK();
  bn_final_sums_kernel<<<cdiv(2 * C, 128), 128, 0, st>>>((const double*)ws, nblocks, C, sums);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int ud3d_bn_backward_apply(const float* x, int ld_x, const float* da, int ld_da, int n, int C, const float* scale, const float* shift,
                           const float* mean, const float* invstd, int relu, const double* sums, double count, float* dx, int ld_dx,
                           int accumulate, const double* count_dev, void* stream) {
  UD3D_CHECK_ARG(x && da && scale && shift && mean && invstd && sums && dx && C > 0 && n >= 0 && ld_x >= C && ld_da >= C && ld_dx >= C &&
                     (count_dev || count > 0.0),
                 "ud3d_bn_backward_apply: bad argument");
  if (n == 0) return UD3D_OK;
  long long total = (long long)n * C;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  bn_bwd_apply_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, ld_x, da, ld_da, n, C, scale, shift, mean, invstd, relu, sums, count, count_dev,
                                                                dx, ld_dx, accumulate);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int ud3d_bn_relu_apply(const float* x, int ld_x, int n, int C, const float* scale, const float* shift, int relu, float* out, int ld_out,
                       void* stream) {
  UD3D_CHECK_ARG(x && scale && shift && out && C > 0 && n >= 0 && ld_x >= C && ld_out >= C, "ud3d_bn_relu_apply: bad argument");
  if (n == 0) return UD3D_OK;
  long long total = (long long)n * C;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  bn_relu_apply_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, ld_x, n, C, scale, shift, relu, out, ld_out);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

size_t ud3d_layernorm_backward_workspace_bytes(int rows, int C) {
  return (size_t)cdiv(rows > 0 ? rows : 1, kLnRowsPerBlock) * 2 * (size_t)(C > 0 ? C : 1) * sizeof(double);
}

int ud3d_layernorm_backward(const float* x, const float* dy, const float* gamma, int rows, int C, float eps, float* dx, double* dgamma_dbeta,
                            void* ws, size_t ws_bytes, void* stream) {
  UD3D_CHECK_ARG(x && dy && gamma && dx && dgamma_dbeta && ws && rows >= 0 && C > 0 && C <= 768, "ud3d_layernorm_backward: bad argument (C <= 768)");
  if (ws_bytes < ud3d_layernorm_backward_workspace_bytes(rows, C)) {
    set_error("ud3d_layernorm_backward: workspace too small");
    return UD3D_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) {
    UD3D_CUDA(cudaMemsetAsync(dgamma_dbeta, 0, (size_t)2 * C * sizeof(double), st));
    return UD3D_OK;
  }
  const int nblocks = cdiv(rows, kLnRowsPerBlock);
  layernorm_bwd_kernel<<<nblocks, 256, (size_t)16 * C * sizeof(float), st>>>(x, dy, gamma, rows, C, eps, dx, (double*)ws);
  UD3D_LAUNCH_CHECK();
  bn_final_sums_kernel<<<cdiv(2 * C, 128), 128, 0, st>>>((const double*)ws, nblocks, C, dgamma_dbeta);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int ud3d_activation_backward(const float* pre, const float* dy, long long total, int act, float* dx, void* stream) {
  UD3D_CHECK_ARG(pre && dy && dx && total >= 0 && (act == 1 || act == 2), "ud3d_activation_backward: bad argument (act 1 relu, 2 gelu)");
  if (total == 0) return UD3D_OK;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  act_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pre, dy, total, act, dx);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int ud3d_activation_forward(const float* pre, long long total, int act, float* out, void* stream) {
  UD3D_CHECK_ARG(pre && out && total >= 0 && (act == 1 || act == 2), "ud3d_activation_forward: bad argument (act 1 relu, 2 gelu)");
  if (total == 0) return UD3D_OK;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  act_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pre, total, act, out);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

size_t ud3d_segmented_mean_backward_workspace_bytes(int n_rows, int n_seg, int C) {
  return (size_t)(n_rows > 0 ? n_rows : 1) * (size_t)(C > 0 ? C : 1) * 8 + (size_t)(n_seg > 0 ? n_seg : 1) * 4;
}

int ud3d_segmented_mean_backward(const float* d_pooled, int C, const int32_t* gather, const int64_t* seg, int n, int n_seg, int n_rows,
                                 float* d_src, void* ws, size_t ws_bytes, void* stream) {
  UD3D_CHECK_ARG(d_pooled && seg && d_src && ws && C > 0 && n >= 0 && n_seg >= 0 && n_rows >= 0, "ud3d_segmented_mean_backward: bad argument");
  UD3D_CHECK_ARG(((uintptr_t)ws & 7) == 0, "ud3d_segmented_mean_backward: workspace must be 8-byte aligned");
  if (ws_bytes < ud3d_segmented_mean_backward_workspace_bytes(n_rows, n_seg, C)) {
    set_error("ud3d_segmented_mean_backward: workspace too small");
    return UD3D_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* acc = (unsigned long long*)ws;
  int32_t* cnt = (int32_t*)(acc + (size_t)(n_rows > 0 ? n_rows : 1) * C);
  UD3D_CUDA(cudaMemsetAsync(ws, 0, ud3d_segmented_mean_backward_workspace_bytes(n_rows, n_seg, C), st));
  if (n > 0 && n_seg > 0) {
    int b1 = cdiv(n, 256);
    if (b1 > 148 * 16) b1 = 148 * 16;
    seg_count_kernel<<<b1, 256, 0, st>>>(seg, n, n_seg, cnt);
    UD3D_LAUNCH_CHECK();
    long long total = (long long)n * C;
    int b2 = (int)((total + 255) / 256);
    if (b2 > 148 * 32) b2 = 148 * 32;
    pool_bwd_scatter_kernel<<<b2, 256, 0, st>>>(d_pooled, C, gather, seg, cnt, n, n_seg, acc);
    UD3D_LAUNCH_CHECK();
  }
  if (n_rows > 0) {
    long long total = (long long)n_rows * C;
    int b3 = (int)((total + 255) / 256);
    if (b3 > 148 * 16) b3 = 148 * 16;
    fix32_to_float_kernel<<<b3, 256, 0, st>>>(acc, total, d_src);
    UD3D_LAUNCH_CHECK();
  }
  return UD3D_OK;
}

size_t ud3d_conv_wgrad_workspace_bytes(int n_out, int K, int c_in, int c_out) {
  if (n_out < 0 || K <= 0 || c_in <= 0 || c_out <= 0) return 0;
  return (size_t)wgrad_slices(n_out) * K * c_in * c_out * sizeof(float);
}

int ud3d_conv_wgrad(const float* x, int ld_x, int c_in, const float* dy, int ld_dy, int c_out, const int32_t* table,
                    int n_out, int K, float* dw, int accumulate, void* ws, size_t ws_bytes, void* stream) {
  UD3D_CHECK_ARG(x && dy && dw && ws && c_in > 0 && c_out > 0 && K > 0 && n_out >= 0, "ud3d_conv_wgrad: bad argument");
  UD3D_CHECK_ARG(table || K == 1, "ud3d_conv_wgrad: identity gather requires K == 1");
  UD3D_CHECK_ARG(ld_x >= c_in && ld_dy >= c_out, "ud3d_conv_wgrad: leading dimension smaller than channel count");
  if (ws_bytes < ud3d_conv_wgrad_workspace_bytes(n_out, K, c_in, c_out)) {
    set_error("ud3d_conv_wgrad: workspace too small");
    return UD3D_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int slices = wgrad_slices(n_out);
  const int rps = cdiv(n_out > 0 ? n_out : 1, slices);
  static const bool use_v1 = [] { const char* e = getenv("UD3D_WGRAD"); return e && e[0] == 'v' && e[1] == '1'; }();   // measurement switch
  if (use_v1) {
    dim3 grid(cdiv(c_out, 32) * cdiv(c_in, 32), K, slices);
    conv_wgrad_partial_kernel<<<grid, 256, 0, st>>>(x, ld_x, c_in, dy, ld_dy, c_out, table, n_out, K, rps, (float*)ws);
  } else if (c_out > 32 && c_in > 32) {
    launch_wgrad_tiled<64, 64>(x, ld_x, c_in, dy, ld_dy, c_out, table, n_out, K, rps, slices, (float*)ws, st);
  } else if (c_out > 32) {
    launch_wgrad_tiled<64, 32>(x, ld_x, c_in, dy, ld_dy, c_out, table, n_out, K, rps, slices, (float*)ws, st);
  } else if (c_in > 32) {
    launch_wgrad_tiled<32, 64>(x, ld_x, c_in, dy, ld_dy, c_out, table, n_out, K, rps, slices, (float*)ws, st);
  } else {
    launch_wgrad_tiled<32, 32>(x, ld_x, c_in, dy, ld_dy, c_out, table, n_out, K, rps, slices, (float*)ws, st);
  }
  UD3D_LAUNCH_CHECK();
  const size_t n_w = (size_t)c_out * K * c_in;
  conv_wgrad_reduce_kernel<<<cdiv((long long)n_w, 256), 256, 0, st>>>((const float*)ws, slices, n_w, dw, accumulate);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

}  // extern "C"
