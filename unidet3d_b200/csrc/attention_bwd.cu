// Backward of the varlen multi-head self-attention core (nn.MultiheadAttention inside SelfAttentionLayer, reference
// unidet3d/encoder.py:24-41; head_dim 32, no mask, dropout 0), fp32 on the CUDA cores -- correctness first: it completes
// the encoder's backward chain; the forward is the tcgen05 kernel of attention_tc.cu.
//   S = Q K^T / sqrt(32),  P = softmax_row(S),  O = P V
//   dV = P^T dO,   dP = dO V^T,   dS = P o (dP - D),  D_i = sum_c dO_ic O_ic,   dQ = dS K / sqrt(32),   dK = dS^T Q / sqrt(32)
// Three passes, one WARP per (token, head), lanes over the other side's tokens (each lane keeps its own 32-vector
// partial sums, combined at the end by shuffles in a fixed order: deterministic, no atomics, nothing T x T is stored):
//   1. stats : lse_i = log sum_j exp(S_ij), D_i
//   2. dQ    : lanes over keys
//   3. dK, dV: lanes over queries
#include "common.cuh"

namespace ud3d {

constexpr int kHd = 32;

__device__ __forceinline__ void load_row32(const float* __restrict__ p, float (&r)[kHd]) {
#pragma unroll
  for (int j = 0; j < kHd / 4; ++j) {
    const float4 v = __ldg((const float4*)p + j);
    r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
  }
}
__device__ __forceinline__ float dot32(const float (&a)[kHd], const float (&b)[kHd]) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kHd; ++c) s = fmaf(a[c], b[c], s);
  return s;
}
// scene [t0, t1) of token t
__device__ __forceinline__ void scene_of(const int32_t* __restrict__ cu, int B, int t, int& t0, int& t1) {
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (cu[mid] <= t) lo = mid; else hi = mid;
  }
  t0 = cu[lo];
  t1 = cu[lo + 1];
}
// sum over lanes of a 32-vector held per lane: result[c] in lane c (fixed butterfly order)
__device__ __forceinline__ float reduce_vec_to_lane(float (&v)[kHd], int lane) {
#pragma unroll
  for (int c = 0; c < kHd; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
  }
  float out = 0.f;
#pragma unroll
  for (int c = 0; c < kHd; ++c) out = lane == c ? v[c] : out;
  return out;
}

__global__ void __launch_bounds__(128) attn_bwd_stats_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ cu, int B,
                                                             int total_T, int H, const float* __restrict__ out,
                                                             const float* __restrict__ d_out, float* __restrict__ lse,
                                                             float* __restrict__ dsum) {
  const int lane = threadIdx.x & 31;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= (long long)total_T * H) return;
  const int t = (int)(w / H), h = (int)(w - (long long)t * H);
  const int d = H * kHd;
  int t0, t1;
  scene_of(cu, B, t, t0, t1);
  float q[kHd];
  load_row32(qkv + (size_t)t * 3 * d + h * kHd, q);
  const float scale = 0.17677669529663688110f;      // 1 / sqrt(32)
  float m = -INFINITY, l = 0.f;
  for (int j = t0 + lane; j < t1; j += 32) {
    float k[kHd];
    load_row32(qkv + (size_t)j * 3 * d + d + h * kHd, k);
    const float s = dot32(q, k) * scale;
    const float mn = fmaxf(m, s);
    l = l * expf(m - mn) + expf(s - mn);
    m = mn;
  }
  // combine (m, l) over the lanes
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), l2 = __shfl_xor_sync(0xffffffffu, l, o);
    const float mn = fmaxf(m, m2);
    l = (m == -INFINITY ? 0.f : l * expf(m - mn)) + (m2 == -INFINITY ? 0.f : l2 * expf(m2 - mn));
    m = mn;
  }
  const float dd = warp_sum(out[(size_t)t * d + h * kHd + lane] * d_out[(size_t)t * d + h * kHd + lane]);
  if (lane == 0) {
    lse[(size_t)t * H + h] = m + logf(l);
    dsum[(size_t)t * H + h] = dd;
  }
}

__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ cu, int B,
                                                          int total_T, int H, const float* __restrict__ d_out,
                                                          const float* __restrict__ lse, const float* __restrict__ dsum,
                                                          float* __restrict__ dqkv) {
  const int lane = threadIdx.x & 31;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= (long long)total_T * H) return;
  const int t = (int)(w / H), h = (int)(w - (long long)t * H);
  const int d = H * kHd;
  int t0, t1;
  scene_of(cu, B, t, t0, t1);
  float q[kHd], go[kHd], acc[kHd];
  load_row32(qkv + (size_t)t * 3 * d + h * kHd, q);
  load_row32(d_out + (size_t)t * d + h * kHd, go);
#pragma unroll
  for (int c = 0; c < kHd; ++c) acc[c] = 0.f;
  const float scale = 0.17677669529663688110f;
  const float L = lse[(size_t)t * H + h], D = dsum[(size_t)t * H + h];
  for (int j = t0 + lane; j < t1; j += 32) {
    float k[kHd], v[kHd];
    load_row32(qkv + (size_t)j * 3 * d + d + h * kHd, k);
    load_row32(qkv + (size_t)j * 3 * d + 2 * d + h * kHd, v);
    const float p = expf(dot32(q, k) * scale - L);
    const float ds = p * (dot32(go, v) - D) * scale;
#pragma unroll
    for (int c = 0; c < kHd; ++c) acc[c] = fmaf(ds, k[c], acc[c]);
  }
  const float r = reduce_vec_to_lane(acc, lane);
  dqkv[(size_t)t * 3 * d + h * kHd + lane] = r;
}

__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ cu, int B,
                                                           int total_T, int H, const float* __restrict__ d_out,
                                                           const float* __restrict__ lse, const float* __restrict__ dsum,
                                                           float* __restrict__ dqkv) {
  const int lane = threadIdx.x & 31;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= (long long)total_T * H) return;
  const int t = (int)(w / H), h = (int)(w - (long long)t * H);      // t = key / value token
  const int d = H * kHd;
  int t0, t1;
  scene_of(cu, B, t, t0, t1);
  float k[kHd], v[kHd], dk[kHd], dv[kHd];
  load_row32(qkv + (size_t)t * 3 * d + d + h * kHd, k);
  load_row32(qkv + (size_t)t * 3 * d + 2 * d + h * kHd, v);
#pragma unroll
  for (int c = 0; c < kHd; ++c) { dk[c] = 0.f; dv[c] = 0.f; }
  const float scale = 0.17677669529663688110f;
  for (int i = t0 + lane; i < t1; i += 32) {
    float q[kHd], go[kHd];
    load_row32(qkv + (size_t)i * 3 * d + h * kHd, q);
    load_row32(d_out + (size_t)i * d + h * kHd, go);
    const float p = expf(dot32(q, k) * scale - lse[(size_t)i * H + h]);
    const float ds = p * (dot32(go, v) - dsum[(size_t)i * H + h]) * scale;
#pragma unroll
    for (int c = 0; c < kHd; ++c) {
      dv[c] = fmaf(p, go[c], dv[c]);
      dk[c] = fmaf(ds, q[c], dk[c]);
    }
  }
  const float rk = reduce_vec_to_lane(dk, lane);
  const float rv = reduce_vec_to_lane(dv, lane);
  dqkv[(size_t)t * 3 * d + d + h * kHd + lane] = rk;
  dqkv[(size_t)t * 3 * d + 2 * d + h * kHd + lane] = rv;
}

// ---------------------------------------------------------------- register-resident variant (the product path)
// Same math, different work split: one THREAD per query (dQ pass) / per key (dK, dV pass) keeps its own 32-float rows and
// accumulators in registers; the other side's rows are staged through shared memory 64 at a time and read by all threads
// of the CTA at the same address (broadcast, conflict-free).  Per (query, key) pair: 96 / 128 FMAs against 16 shared
// loads of 16 bytes, no cross-thread reduction, no atomics -- the CUDA-core FMA pipe is the limit instead of the L1/L2
// row traffic of the warp-per-token kernels above (every lane of those fetches a 128-byte row per pair).
// Grid (ceil(total_T / 128), H, B): CTAs beyond the end of their scene exit at once.
constexpr int kStage = 64;

__device__ __forceinline__ void stage_rows(float4 (*dst)[kHd / 4], const float* __restrict__ base, size_t ld, int row0, int n_valid) {
  for (int idx = threadIdx.x; idx < kStage * (kHd / 4); idx += blockDim.x) {
    const int r = idx >> 3, c4 = idx & 7;
    dst[r][c4] = r < n_valid ? __ldg((const float4*)(base + (size_t)(row0 + r) * ld) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ float dot32s(const float (&a)[kHd], const float4* __restrict__ b) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;      // four independent chains (the FMA latency is 4 cycles)
#pragma unroll
  for (int j = 0; j < kHd / 4; ++j) {
    const float4 v = b[j];
    s0 = fmaf(a[4 * j], v.x, s0); s1 = fmaf(a[4 * j + 1], v.y, s1); s2 = fmaf(a[4 * j + 2], v.z, s2); s3 = fmaf(a[4 * j + 3], v.w, s3);
  }
  return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ void axpy32s(float (&acc)[kHd], float a, const float4* __restrict__ b) {
#pragma unroll
  for (int j = 0; j < kHd / 4; ++j) {
    const float4 v = b[j];
    acc[4 * j] = fmaf(a, v.x, acc[4 * j]); acc[4 * j + 1] = fmaf(a, v.y, acc[4 * j + 1]);
    acc[4 * j + 2] = fmaf(a, v.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(a, v.w, acc[4 * j + 3]);
  }
}
__device__ __forceinline__ void store_row32(float* __restrict__ p, const float (&r)[kHd]) {
#pragma unroll
  for (int j = 0; j < kHd / 4; ++j) ((float4*)p)[j] = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
}

// thread = query: softmax statistics (lse, D) + dQ
__global__ void __launch_bounds__(128) attn_bwd_dq_reg_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ cu, int H,
                                                              const float* __restrict__ out, const float* __restrict__ d_out,
                                                              float* __restrict__ lse, float* __restrict__ dsum,
                                                              float* __restrict__ dqkv) {
  __shared__ float4 sK[kStage][kHd / 4];
  __shared__ float4 sV[kStage][kHd / 4];
  const int h = blockIdx.y, t0 = cu[blockIdx.z], t1 = cu[blockIdx.z + 1];
  const int i0 = t0 + blockIdx.x * 128;
  if (i0 >= t1) return;
  const int i = i0 + threadIdx.x;
  const bool active = i < t1;
  const int d = H * kHd;
  const size_t ld = (size_t)3 * d;
  float q[kHd], go[kHd], acc[kHd];
  float D = 0.f;
#pragma unroll
  for (int c = 0; c < kHd; ++c) { q[c] = 0.f; go[c] = 0.f; acc[c] = 0.f; }
  if (active) {
    load_row32(qkv + (size_t)i * ld + h * kHd, q);
    load_row32(d_out + (size_t)i * d + h * kHd, go);
    float o[kHd];
    load_row32(out + (size_t)i * d + h * kHd, o);
    D = dot32(go, o);
  }
  const float scale = 0.17677669529663688110f;      // 1 / sqrt(32)
  const float* kbase = qkv + d + h * kHd;
  const float* vbase = qkv + 2 * d + h * kHd;
  // pass 1: running (max, sum of exponentials) over the keys of the scene
  float m = -INFINITY, l = 0.f;
  for (int j0 = t0; j0 < t1; j0 += kStage) {
    const int n = min(kStage, t1 - j0);
    __syncthreads();
    stage_rows(sK, kbase, ld, j0, n);
    __syncthreads();
    for (int jj = 0; jj < n; ++jj) {
      const float s = dot32s(q, sK[jj]) * scale;
      if (s > m) {
        l = l * expf(m - s) + 1.f;
        m = s;
      } else {
        l += expf(s - m);
      }
    }
  }
  const float L = m + logf(l);
  // pass 2: dQ_i = scale * sum_j P_ij (dP_ij - D_i) K_j
  for (int j0 = t0; j0 < t1; j0 += kStage) {
    const int n = min(kStage, t1 - j0);
    __syncthreads();
    stage_rows(sK, kbase, ld, j0, n);
    stage_rows(sV, vbase, ld, j0, n);
    __syncthreads();
    for (int jj = 0; jj < n; ++jj) {
      const float p = expf(dot32s(q, sK[jj]) * scale - L);
      const float ds = p * (dot32s(go, sV[jj]) - D) * scale;
      axpy32s(acc, ds, sK[jj]);
    }
  }
  if (active) {
    lse[(size_t)i * H + h] = L;
    dsum[(size_t)i * H + h] = D;
    store_row32(dqkv + (size_t)i * ld + h * kHd, acc);
  }
}

// thread = key: dK_j = scale * sum_i dS_ij Q_i,  dV_j = sum_i P_ij dO_i
__global__ void __launch_bounds__(128) attn_bwd_dkv_reg_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ cu, int H,
                                                               const float* __restrict__ d_out, const float* __restrict__ lse,
                                                               const float* __restrict__ dsum, float* __restrict__ dqkv) {
  __shared__ float4 sQ[kStage][kHd / 4];
  __shared__ float4 sG[kStage][kHd / 4];
  __shared__ float sL[kStage], sD[kStage];
  const int h = blockIdx.y, t0 = cu[blockIdx.z], t1 = cu[blockIdx.z + 1];
  const int j0 = t0 + blockIdx.x * 128;
  if (j0 >= t1) return;
  const int j = j0 + threadIdx.x;
  const bool active = j < t1;
  const int d = H * kHd;
  const size_t ld = (size_t)3 * d;
  float k[kHd], v[kHd], dk[kHd], dv[kHd];
#pragma unroll
  for (int c = 0; c < kHd; ++c) { k[c] = 0.f; v[c] = 0.f; dk[c] = 0.f; dv[c] = 0.f; }
  if (active) {
    load_row32(qkv + (size_t)j * ld + d + h * kHd, k);
    load_row32(qkv + (size_t)j * ld + 2 * d + h * kHd, v);
  }
  const float scale = 0.17677669529663688110f;
  for (int i0 = t0; i0 < t1; i0 += kStage) {
    const int n = min(kStage, t1 - i0);
    __syncthreads();
    stage_rows(sQ, qkv + h * kHd, ld, i0, n);
    stage_rows(sG, d_out + h * kHd, (size_t)d, i0, n);
    if (threadIdx.x < kStage) {
      const bool ok = (int)threadIdx.x < n;
      sL[threadIdx.x] = ok ? lse[(size_t)(i0 + threadIdx.x) * H + h] : 0.f;
      sD[threadIdx.x] = ok ? dsum[(size_t)(i0 + threadIdx.x) * H + h] : 0.f;
    }
    __syncthreads();
    for (int ii = 0; ii < n; ++ii) {
      const float p = expf(dot32s(k, sQ[ii]) * scale - sL[ii]);
      const float ds = p * (dot32s(v, sG[ii]) - sD[ii]) * scale;
      axpy32s(dv, p, sG[ii]);
      axpy32s(dk, ds, sQ[ii]);
    }
  }
  if (active) {
    store_row32(dqkv + (size_t)j * ld + d + h * kHd, dk);
    store_row32(dqkv + (size_t)j * ld + 2 * d + h * kHd, dv);
  }
}

}  // namespace ud3d

using namespace ud3d;

extern "C" {

size_t ud3d_attention_bwd_workspace_bytes(int total_T, int num_heads) {
  return (size_t)(total_T > 0 ? total_T : 1) * (size_t)(num_heads > 0 ? num_heads : 1) * 2 * sizeof(float);
}

int ud3d_attention_bwd(const float* qkv, const int32_t* cu_seqlens, int B, int total_T, int num_heads, const float* out,
                       const float* d_out, float* dqkv, void* ws, size_t ws_bytes, void* stream) {
  UD3D_CHECK_ARG(qkv && cu_seqlens && out && d_out && dqkv && ws, "ud3d_attention_bwd: NULL argument");
  UD3D_CHECK_ARG(B > 0 && total_T >= 0 && num_heads > 0, "ud3d_attention_bwd: bad sizes");
  UD3D_CHECK_ARG((((uintptr_t)qkv | (uintptr_t)d_out) & 15) == 0, "ud3d_attention_bwd: qkv / d_out must be 16-byte aligned");
  if (ws_bytes < ud3d_attention_bwd_workspace_bytes(total_T, num_heads)) {
    set_error("ud3d_attention_bwd: workspace too small");
    return UD3D_EWORKSPACE;
  }
  if (total_T == 0) return UD3D_OK;
  cudaStream_t st = (cudaStream_t)stream;
  float* lse = (float*)ws;
  float* dsum = lse + (size_t)total_T * num_heads;
  const long long warps = (long long)total_T * num_heads;
  const int blocks = (int)((warps + 3) / 4);
  attn_bwd_stats_kernel<<<blocks, 128, 0, st>>>(qkv, cu_seqlens, B, total_T, num_heads, out, d_out, lse, dsum);
  UD3D_LAUNCH_CHECK();
  attn_bwd_dq_kernel<<<blocks, 128, 0, st>>>(qkv, cu_seqlens, B, total_T, num_heads, d_out, lse, dsum, dqkv);
  UD3D_LAUNCH_CHECK();
  attn_bwd_dkv_kernel<<<blocks, 128, 0, st>>>(qkv, cu_seqlens, B, total_T, num_heads, d_out, lse, dsum, dqkv);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int ud3d_attention_bwd_reg(const float* qkv, const int32_t* cu_seqlens, int B, int total_T, int num_heads, const float* out,
                           const float* d_out, float* dqkv, void* ws, size_t ws_bytes, void* stream) {
  UD3D_CHECK_ARG(qkv && cu_seqlens && out && d_out && dqkv && ws, "ud3d_attention_bwd_reg: NULL argument");
  UD3D_CHECK_ARG(B > 0 && B <= 65535 && total_T >= 0 && num_heads > 0 && num_heads <= 65535, "ud3d_attention_bwd_reg: bad sizes");
  UD3D_CHECK_ARG((((uintptr_t)qkv | (uintptr_t)d_out | (uintptr_t)out | (uintptr_t)dqkv) & 15) == 0,
                 "ud3d_attention_bwd_reg: qkv / out / d_out / dqkv must be 16-byte aligned");
  if (ws_bytes < ud3d_attention_bwd_workspace_bytes(total_T, num_heads)) {
    set_error("ud3d_attention_bwd_reg: workspace too small");
    return UD3D_EWORKSPACE;
  }
  if (total_T == 0) return UD3D_OK;
  cudaStream_t st = (cudaStream_t)stream;
  float* lse = (float*)ws;
  float* dsum = lse + (size_t)total_T * num_heads;
  dim3 grid(cdiv(total_T, 128), num_heads, B);
  attn_bwd_dq_reg_kernel<<<grid, 128, 0, st>>>(qkv, cu_seqlens, num_heads, out, d_out, lse, dsum, dqkv);
  UD3D_LAUNCH_CHECK();
  attn_bwd_dkv_reg_kernel<<<grid, 128, 0, st>>>(qkv, cu_seqlens, num_heads, d_out, lse, dsum, dqkv);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

}  // extern "C"
