// Training-side kernels of the backbone (SURVEY.md section 8a R6 second half, 8f rank 2):
//   * batch statistics of train-mode (Sync)BatchNorm over all active voxels of the batch (spconv_unet.py:119-124,
//     unidet3d.py:104-107): deterministic two-stage per-channel sum / sum of squares in fp64; the cross-rank reduction
//     of SyncBatchNorm is one all-reduce of the [2C] sums between the two calls (host side, torch.distributed / NCCL);
//   * the fold into the (scale, shift) form every conv consumes, with the running-statistics update;
//   * weight gradient of a sparse convolution: dW[co, k, ci] = sum_o dY[o, co] * X[table[k][o], ci].
#include "common.cuh"

namespace ud3d {

constexpr int kBnRowsPerBlock = 512;

// block b: rows [b * 512, ...); thread t: channels t, t + 256, ... (coalesced along a row)
__global__ void __launch_bounds__(256) bn_partial_sums_kernel(const float* __restrict__ x, int ld, int n, int C,
                                                              double* __restrict__ part) {
  const int r0 = blockIdx.x * kBnRowsPerBlock;
  const int r1 = min(n, r0 + kBnRowsPerBlock);
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
__global__ void bn_train_fold_kernel(const double* __restrict__ sums, double count, int C, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float eps, float momentum, float* running_mean,
                                     float* running_var, float* __restrict__ scale, float* __restrict__ shift,
                                     float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
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
                       float* save_mean, float* save_invstd, void* stream) {
  UD3D_CHECK_ARG(sums && scale && shift && C > 0 && count > 0.0, "ud3d_bn_train_fold: bad argument");
  bn_train_fold_kernel<<<cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(sums, count, C, gamma, beta, eps, momentum, running_mean,
                                                                     running_var, scale, shift, save_mean, save_invstd);
  UD3D_LAUNCH_CHECK();
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
  dim3 grid(cdiv(c_out, 32) * cdiv(c_in, 32), K, slices);
  conv_wgrad_partial_kernel<<<grid, 256, 0, st>>>(x, ld_x, c_in, dy, ld_dy, c_out, table, n_out, K, rps, (float*)ws);
  UD3D_LAUNCH_CHECK();
  const size_t n_w = (size_t)c_out * K * c_in;
  conv_wgrad_reduce_kernel<<<cdiv((long long)n_w, 256), 256, 0, st>>>((const float*)ws, slices, n_w, dw, accumulate);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

}  // extern "C"
