// Superpoint pooling and the non-GEMM pieces of the superpoint transformer encoder:
// segmented mean (warp-shuffle run-length reduction + 64-bit fixed-point red.add: deterministic), LayerNorm(+residual),
// varlen multi-head self-attention (flash-style, bf16 hi/lo split on tensor cores), box decode.
#include "common.cuh"

namespace ud3d {

// ---------------------------------------------------------------- segmented mean
// One warp owns a run of consecutive points; lane = channel (C <= 32).  Consecutive points that
// share a segment id are accumulated in registers (shuffle-broadcast ids), one red.global.add per
// run and channel.  The [n_pts, C] gather of the reference (x.features[inverse_mapping]) is never
// materialised; the output BatchNorm+ReLU is applied on the fly.
// DETERMINISTIC: a run's fp32 sum (fixed point order inside a fixed 128-point window) is converted to 64-bit fixed
// point (2^-24 resolution, |sum| < 5e11) before the atomic: integer addition is associative, so the result does not
// depend on the order in which the atomics of different warps land.  Counts are integer atomics.
constexpr int kPoolPtsPerWarp = 128;
constexpr float kPoolFixScale = 16777216.f;              // 2^24
constexpr double kPoolFixInv = 1.0 / 16777216.0;

__global__ void __launch_bounds__(256) segmented_sum_kernel(const float* __restrict__ src, int ld, int C,
                                                            const int32_t* __restrict__ gather,
                                                            const int64_t* __restrict__ seg, int n, int n_seg,
                                                            const float* __restrict__ scale,
                                                            const float* __restrict__ shift, int relu,
                                                            unsigned long long* acc_fix, int* cnt) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long p0 = (long long)gw * kPoolPtsPerWarp;
  if (p0 >= n) return;
  const int pend = (int)min((long long)n, p0 + kPoolPtsPerWarp);
  const float sc = (scale && lane < C) ? scale[lane] : 1.f;
  const float sh = (shift && lane < C) ? shift[lane] : 0.f;
  long long cur = -1;
  float acc = 0.f;
  int run = 0;
  for (int base = (int)p0; base < pend; base += 32) {
    int p = base + lane;
    long long my_id = p < pend ? seg[p] : -1;
    int my_g = p < pend ? (gather ? gather[p] : p) : -1;
    int cntp = min(32, pend - base);
    // (issuing the 32 row loads of a chunk back to back before the run-length walk was measured slower -- 98 vs 86 us:
    // 80 registers; with unsorted points the kernel is bound by the L2 atomics, one per point and channel)
    for (int j = 0; j < cntp; ++j) {
      long long id = __shfl_sync(0xffffffffu, my_id, j);
      int g = __shfl_sync(0xffffffffu, my_g, j);
      if (id != cur) {
        if (cur >= 0 && cur < n_seg) {
          if (lane < C) atomicAdd(acc_fix + (size_t)cur * C + lane, (unsigned long long)__float2ll_rn(acc * kPoolFixScale));
          if (lane == 0) atomicAdd(cnt + cur, run);
        }
        cur = id;
        acc = 0.f;
        run = 0;
      }
      float v = 0.f;
      if (lane < C && g >= 0) {
        v = __ldg(src + (size_t)g * ld + lane);
        v = fmaf(v, sc, sh);
        if (relu) v = fmaxf(v, 0.f);
      }
      acc += v;
      run += 1;
    }
  }
  if (cur >= 0 && cur < n_seg) {
    if (lane < C) atomicAdd(acc_fix + (size_t)cur * C + lane, (unsigned long long)__float2ll_rn(acc * kPoolFixScale));
    if (lane == 0) atomicAdd(cnt + cur, run);
  }
}

// C <= 4 (superpoint centres: xyz of a [n, 6] point matrix): one THREAD per point.  The lane-per-channel kernel above
// keeps 3 of 32 lanes busy and walks the points of a warp one after the other (98 us for 800k points); here every lane
// has its own point and issues its C fixed-point atomics directly (same sums: integer addition is order-independent).
__global__ void __launch_bounds__(256) segmented_sum_small_kernel(const float* __restrict__ src, int ld, int C,
                                                                  const int32_t* __restrict__ gather,
                                                                  const int64_t* __restrict__ seg, int n, int n_seg,
                                                                  const float* __restrict__ scale,
                                                                  const float* __restrict__ shift, int relu,
                                                                  unsigned long long* acc_fix, int* cnt) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const long long id = seg[p];
    if (id < 0 || id >= n_seg) continue;
    const int g = gather ? gather[p] : p;
    atomicAdd(cnt + id, 1);
    if (g < 0) continue;
    for (int c = 0; c < C; ++c) {
      float v = __ldg(src + (size_t)g * ld + c);
      if (scale) v = fmaf(v, scale[c], shift[c]);
      if (relu) v = fmaxf(v, 0.f);
      atomicAdd(acc_fix + (size_t)id * C + c, (unsigned long long)__float2ll_rn(v * kPoolFixScale));
    }
  }
}

__global__ void segmented_norm_kernel(float* out, const unsigned long long* __restrict__ acc_fix, const int* __restrict__ cnt,
                                      int n_seg, int C) {
  long long total = (long long)n_seg * C;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c = cnt[t / C];
    out[t] = (float)((double)(long long)acc_fix[t] * kPoolFixInv / (double)(c > 1 ? c : 1));
  }
}

// ---------------------------------------------------------------- LayerNorm (+ residual), warp per row
template <int PER_LANE>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float* __restrict__ out, float* __restrict__ out_split, int rows,
                                                        int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * C;
  const float* rr = res ? res + (size_t)row * C : nullptr;
  float v[PER_LANE];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < PER_LANE / 4; ++j) {
    int c = j * 128 + lane * 4;
    float4 t = *(const float4*)(xr + c);
    if (rr) {
      float4 u = *(const float4*)(rr + c);
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    v[j * 4 + 0] = t.x; v[j * 4 + 1] = t.y; v[j * 4 + 2] = t.z; v[j * 4 + 3] = t.w;
    s += t.x + t.y + t.z + t.w;
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < PER_LANE; ++j) {
    float d = v[j] - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int j = 0; j < PER_LANE / 4; ++j) {
    int c = j * 128 + lane * 4;
    float4 g = *(const float4*)(gamma + c), b = *(const float4*)(beta + c);
    float4 o;
    o.x = (v[j * 4 + 0] - mean) * rstd * g.x + b.x;
    o.y = (v[j * 4 + 1] - mean) * rstd * g.y + b.y;
    o.z = (v[j * 4 + 2] - mean) * rstd * g.z + b.z;
    o.w = (v[j * 4 + 3] - mean) * rstd * g.w + b.w;
    if (out) *(float4*)(out + (size_t)row * C + c) = o;
    if (out_split) {
      // operand form: chunk = c / 32, 4 channels at (c % 32): hi 8 bytes, lo 8 bytes (+64)
      uint32_t h0, l0, h1, l1;
      split_bf16x2(o.x, o.y, h0, l0);
      split_bf16x2(o.z, o.w, h1, l1);
      uint8_t* d = (uint8_t*)(out_split + (size_t)row * C) + (c >> 5) * 128 + (c & 31) * 2;
      *(uint2*)d = make_uint2(h0, h1);
      *(uint2*)(d + 64) = make_uint2(l0, l1);
    }
  }
}

// generic fallback: C % 32 == 0, scalar strided
__global__ void __launch_bounds__(256) layernorm_generic_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, float* __restrict__ out,
                                                                int rows, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * C;
  const float* rr = res ? res + (size_t)row * C : nullptr;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c] + (rr ? rr[c] : 0.f);
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) {
    float d = xr[c] + (rr ? rr[c] : 0.f) - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  for (int c = lane; c < C; c += 32) {
    float d = xr[c] + (rr ? rr[c] : 0.f) - mean;
    out[(size_t)row * C + c] = d * rstd * gamma[c] + beta[c];
  }
}

// ---------------------------------------------------------------- attention (head_dim 32, varlen)
// CTA = 64 queries of one (scene, head); 4 warps x 16 query rows.  K/V tiles of 64 keys are staged in
// shared memory as bf16 hi/lo (V transposed); S = QK^T and O += PV run on mma.sync m16n8k16 with the
// three-term hi/lo split (fp32-grade accuracy), online softmax in registers (exp2, log2e folded in Q).
constexpr int kAttQ = 64, kAttKV = 64, kHeadDim = 32;
constexpr int kKPad = kHeadDim + 8;   // bf16 elements per K row (pad -> conflict-free fragment loads)
constexpr int kVPad = kAttKV + 8;

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <bool SPLIT_OUT>
__global__ void __launch_bounds__(128) attention_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ cu,
                                                        int num_heads, float* __restrict__ out) {
  const int b = blockIdx.z, h = blockIdx.y;
  const int t0 = cu[b];
  const int T = cu[b + 1] - t0;
  const int q0 = blockIdx.x * kAttQ;
  if (q0 >= T) return;
  const int d_model = num_heads * kHeadDim;
  const int ld = 3 * d_model;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;

  __shared__ __align__(16) __nv_bfloat16 sKh[kAttKV][kKPad], sKl[kAttKV][kKPad];
  __shared__ __align__(16) __nv_bfloat16 sVh[kHeadDim][kVPad], sVl[kHeadDim][kVPad];

  // ---- Q fragments (scaled by log2(e)/sqrt(32)), hi/lo
  const float qscale = 1.44269504088896340736f * 0.17677669529663688110f;
  uint32_t qh[2][4], ql[2][4];
  {
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
      for (int part = 0; part < 4; ++part) {
        int row = (part & 1) ? r1 : r0;
        int col = ks * 16 + 2 * t + ((part >> 1) ? 8 : 0);
        float2 v = make_float2(0.f, 0.f);
        if (row < T) v = *(const float2*)(qkv + (size_t)(t0 + row) * ld + h * kHeadDim + col);
        split_bf16x2(v.x * qscale, v.y * qscale, qh[ks][part], ql[ks][part]);
      }
    }
  }
  float o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  for (int kv0 = 0; kv0 < T; kv0 += kAttKV) {
    __syncthreads();   // previous tile fully consumed
    {
      // stage K (row-major) and V (transposed): thread -> (key = tid/2, 16 dims)
      const int key = tid >> 1, half = tid & 1;
      const int kr = kv0 + key;
      float kf[16], vf[16];
      if (kr < T) {
        const float* kp = qkv + (size_t)(t0 + kr) * ld + d_model + h * kHeadDim + half * 16;
        const float* vp = kp + d_model;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 a = __ldg((const float4*)kp + j), c = __ldg((const float4*)vp + j);
          kf[4 * j] = a.x; kf[4 * j + 1] = a.y; kf[4 * j + 2] = a.z; kf[4 * j + 3] = a.w;
          vf[4 * j] = c.x; vf[4 * j + 1] = c.y; vf[4 * j + 2] = c.z; vf[4 * j + 3] = c.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) kf[j] = vf[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t hi, lo;
        split_bf16x2(kf[2 * j], kf[2 * j + 1], hi, lo);
        *(uint32_t*)&sKh[key][half * 16 + 2 * j] = hi;
        *(uint32_t*)&sKl[key][half * 16 + 2 * j] = lo;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        __nv_bfloat16 hi = __float2bfloat16_rn(vf[j]);
        __nv_bfloat16 lo = __float2bfloat16_rn(vf[j] - __bfloat162float(hi));
        sVh[half * 16 + j][key] = hi;
        sVl[half * 16 + j][key] = lo;
      }
    }
    __syncthreads();

    // ---- S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int j = 0; j < 4; ++j) s[nt][j] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int key = nt * 8 + g;
        uint32_t bh0 = *(const uint32_t*)&sKh[key][ks * 16 + 2 * t], bh1 = *(const uint32_t*)&sKh[key][ks * 16 + 2 * t + 8];
        uint32_t bl0 = *(const uint32_t*)&sKl[key][ks * 16 + 2 * t], bl1 = *(const uint32_t*)&sKl[key][ks * 16 + 2 * t + 8];
        mma_bf16_16816(s[nt], qh[ks], bh0, bh1);
        mma_bf16_16816(s[nt], ql[ks], bh0, bh1);
        mma_bf16_16816(s[nt], qh[ks], bl0, bl1);
      }
    }
    // ---- mask keys >= T, online softmax
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int kc = kv0 + nt * 8 + 2 * t;
      if (kc >= T) s[nt][0] = s[nt][2] = -INFINITY;
      if (kc + 1 >= T) s[nt][1] = s[nt][3] = -INFINITY;
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
    float alpha[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      float m_new = fmaxf(m_run[r], mx[r]);
      alpha[r] = exp2f(m_run[r] - m_new);
      m_run[r] = m_new;
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - m_run[0]);
      s[nt][1] = exp2f(s[nt][1] - m_run[0]);
      s[nt][2] = exp2f(s[nt][2] - m_run[1]);
      s[nt][3] = exp2f(s[nt][3] - m_run[1]);
      rs[0] += s[nt][0] + s[nt][1];
      rs[1] += s[nt][2] + s[nt][3];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
      l_run[r] = l_run[r] * alpha[r] + rs[r];
    }
#pragma unroll
    for (int dn = 0; dn < 4; ++dn) {
      o[dn][0] *= alpha[0]; o[dn][1] *= alpha[0];
      o[dn][2] *= alpha[1]; o[dn][3] *= alpha[1];
    }
    // ---- O += P V   (P from the S accumulators, hi/lo split)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t ph[4], pl[4];
      split_bf16x2(s[2 * j][0], s[2 * j][1], ph[0], pl[0]);
      split_bf16x2(s[2 * j][2], s[2 * j][3], ph[1], pl[1]);
      split_bf16x2(s[2 * j + 1][0], s[2 * j + 1][1], ph[2], pl[2]);
      split_bf16x2(s[2 * j + 1][2], s[2 * j + 1][3], ph[3], pl[3]);
#pragma unroll
      for (int dn = 0; dn < 4; ++dn) {
        const int dim = dn * 8 + g;
        uint32_t vh0 = *(const uint32_t*)&sVh[dim][j * 16 + 2 * t], vh1 = *(const uint32_t*)&sVh[dim][j * 16 + 2 * t + 8];
        uint32_t vl0 = *(const uint32_t*)&sVl[dim][j * 16 + 2 * t], vl1 = *(const uint32_t*)&sVl[dim][j * 16 + 2 * t + 8];
        mma_bf16_16816(o[dn], ph, vh0, vh1);
        mma_bf16_16816(o[dn], pl, vh0, vh1);
        mma_bf16_16816(o[dn], ph, vl0, vl1);
      }
    }
  }
  // ---- finalize
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
#pragma unroll
  for (int dn = 0; dn < 4; ++dn) {
    const int col = h * kHeadDim + dn * 8 + 2 * t;
    if constexpr (SPLIT_OUT) {
      // operand form: head h == 32-channel chunk h; hi pair at (dn*8+2t)*2 bytes, lo pair +64
      uint32_t hi, lo;
      if (r0 < T) {
        split_bf16x2(o[dn][0] * inv0, o[dn][1] * inv0, hi, lo);
        uint8_t* d = (uint8_t*)(out + (size_t)(t0 + r0) * d_model) + h * 128 + (dn * 8 + 2 * t) * 2;
        *(uint32_t*)d = hi;
        *(uint32_t*)(d + 64) = lo;
      }
      if (r1 < T) {
        split_bf16x2(o[dn][2] * inv1, o[dn][3] * inv1, hi, lo);
        uint8_t* d = (uint8_t*)(out + (size_t)(t0 + r1) * d_model) + h * 128 + (dn * 8 + 2 * t) * 2;
        *(uint32_t*)d = hi;
        *(uint32_t*)(d + 64) = lo;
      }
    } else {
      if (r0 < T) *(float2*)(out + (size_t)(t0 + r0) * d_model + col) = make_float2(o[dn][0] * inv0, o[dn][1] * inv0);
      if (r1 < T) *(float2*)(out + (size_t)(t0 + r1) * d_model + col) = make_float2(o[dn][2] * inv1, o[dn][3] * inv1);
    }
  }
}

// ---------------------------------------------------------------- attention on operand-form q|k|v
// Input: the QKV projection already in operand form (per token, per head: 64 B bf16 hi | 64 B bf16 lo for each of
// q, k, v; written once by the QKV GEMM epilogue).  CTA = 128 queries of one (scene, head), 8 warps x 16 rows;
// K/V tiles of 64 keys stream through a 2-stage cp.async ring as raw 128-byte rows (XOR-swizzled 16-byte chunks:
// conflict-free fragment loads, V fragments via ldmatrix.trans) -- no per-tile conversion, loads overlap the MMAs.
// 2^x on the MUFU unit, flush-to-zero, no range fix-up code around it (relative error 2^-22)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int kAtt2Q = 128, kAtt2KV = 64;

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t smem_addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_addr));
}

__global__ void __launch_bounds__(256) attention_opform_kernel(const uint8_t* __restrict__ qkv, const int32_t* __restrict__ cu,
                                                               int num_heads, uint8_t* __restrict__ out) {
  const int b = blockIdx.z, h = blockIdx.y;
  const int t0 = cu[b];
  const int T = cu[b + 1] - t0;
  const int q0 = blockIdx.x * kAtt2Q;
  if (q0 >= T) return;
  const int d_model = num_heads * kHeadDim;
  const size_t ldb = (size_t)3 * d_model * 4;          // bytes per token row of qkv
  const size_t ldo = (size_t)d_model * 4;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;

  __shared__ __align__(128) uint8_t sK[2][kAtt2KV * 128];
  __shared__ __align__(128) uint8_t sV[2][kAtt2KV * 128];

  auto load_tile = [&](int stage, int kv0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int id = tid + 256 * i;
      const int which = id >> 9, rem = id & 511;
      const int key = rem >> 3, j = rem & 7;
      const int kr = kv0 + key;
      const uint8_t* src = qkv + (size_t)(t0 + (kr < T ? kr : 0)) * ldb + (size_t)((which ? 2 : 1) * num_heads + h) * 128 + j * 16;
      uint8_t* dst = (which ? sV[stage] : sK[stage]) + key * 128 + ((j ^ (key & 7)) << 4);
      cp_async_16_zfill(smem_u32(dst), src, kr < T ? 16u : 0u);
    }
    cp_async_commit();
  };

  // ---- Q fragments straight from the operand form
  uint32_t qh[2][4], ql[2][4];
  {
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
      for (int part = 0; part < 4; ++part) {
        const int row = (part & 1) ? r1 : r0;
        const int col = ks * 16 + 2 * t + ((part >> 1) ? 8 : 0);
        uint32_t vh = 0, vl = 0;
        if (row < T) {
          const uint8_t* qp = qkv + (size_t)(t0 + row) * ldb + (size_t)h * 128 + col * 2;
          vh = *(const uint32_t*)qp;
          vl = *(const uint32_t*)(qp + 64);
        }
        qh[ks][part] = vh;
        ql[ks][part] = vl;
      }
    }
  }
  const float qscale = 1.44269504088896340736f * 0.17677669529663688110f;   // log2(e) / sqrt(32)
  float o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  const int n_tiles = (T + kAtt2KV - 1) / kAtt2KV;
  load_tile(0, 0);
  for (int it = 0; it < n_tiles; ++it) {
    const int stage = it & 1;
    const int kv0 = it * kAtt2KV;
    if (it + 1 < n_tiles) {
      load_tile(stage ^ 1, kv0 + kAtt2KV);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint8_t* Ks = sK[stage];
    const uint32_t Vs = smem_u32(sV[stage]);

    // ---- S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int j = 0; j < 4; ++j) s[nt][j] = 0.f;
      const uint8_t* krow = Ks + (nt * 8 + g) * 128;      // key & 7 == g
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint32_t bh0 = *(const uint32_t*)(krow + (((2 * ks) ^ g) << 4) + 4 * t);
        const uint32_t bh1 = *(const uint32_t*)(krow + (((2 * ks + 1) ^ g) << 4) + 4 * t);
        const uint32_t bl0 = *(const uint32_t*)(krow + (((4 + 2 * ks) ^ g) << 4) + 4 * t);
        const uint32_t bl1 = *(const uint32_t*)(krow + (((5 + 2 * ks) ^ g) << 4) + 4 * t);
        mma_bf16_16816(s[nt], qh[ks], bh0, bh1);
        mma_bf16_16816(s[nt], ql[ks], bh0, bh1);
        mma_bf16_16816(s[nt], qh[ks], bl0, bl1);
      }
    }
    // ---- mask keys >= T, online softmax (scores are scaled by log2(e)/sqrt(d) inside the exponent)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int kc = kv0 + nt * 8 + 2 * t;
      if (kc >= T) s[nt][0] = s[nt][2] = -INFINITY;
      if (kc + 1 >= T) s[nt][1] = s[nt][3] = -INFINITY;
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
    float alpha[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      mx[r] *= qscale;                      // the running maximum lives in the scaled (log2) domain; qscale > 0
      float m_new = fmaxf(m_run[r], mx[r]);
      alpha[r] = ex2_approx(m_run[r] - m_new);
      m_run[r] = m_new;
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = ex2_approx(fmaf(s[nt][0], qscale, -m_run[0]));     // scale folded into the exponent's FMA
      s[nt][1] = ex2_approx(fmaf(s[nt][1], qscale, -m_run[0]));
      s[nt][2] = ex2_approx(fmaf(s[nt][2], qscale, -m_run[1]));
      s[nt][3] = ex2_approx(fmaf(s[nt][3], qscale, -m_run[1]));
      rs[0] += s[nt][0] + s[nt][1];
      rs[1] += s[nt][2] + s[nt][3];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
      l_run[r] = l_run[r] * alpha[r] + rs[r];
    }
#pragma unroll
    for (int dn = 0; dn < 4; ++dn) {
      o[dn][0] *= alpha[0]; o[dn][1] *= alpha[0];
      o[dn][2] *= alpha[1]; o[dn][3] *= alpha[1];
    }
    // ---- O += P V : V fragments by ldmatrix.trans from the row-major (key, dim) tile
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t ph[4], pl[4];
      split_bf16x2(s[2 * j][0], s[2 * j][1], ph[0], pl[0]);
      split_bf16x2(s[2 * j][2], s[2 * j][3], ph[1], pl[1]);
      split_bf16x2(s[2 * j + 1][0], s[2 * j + 1][1], ph[2], pl[2]);
      split_bf16x2(s[2 * j + 1][2], s[2 * j + 1][3], ph[3], pl[3]);
      // lane -> (matrix = lane >> 3, row = lane & 7): matrices {keys 0-7, keys 8-15} x {dim tile dn, dn + 1}
      const int key = 16 * j + (lane & 7) + ((lane >> 3) & 1) * 8;
      const uint32_t vrow = Vs + key * 128;
      const int csel = lane >> 4;
#pragma unroll
      for (int dp = 0; dp < 2; ++dp) {
        uint32_t vh[4], vl[4];
        ldmatrix_x4_trans(vh, vrow + (((2 * dp + csel) ^ (key & 7)) << 4));
        ldmatrix_x4_trans(vl, vrow + (((4 + 2 * dp + csel) ^ (key & 7)) << 4));
        mma_bf16_16816(o[2 * dp], ph, vh[0], vh[1]);
        mma_bf16_16816(o[2 * dp], pl, vh[0], vh[1]);
        mma_bf16_16816(o[2 * dp], ph, vl[0], vl[1]);
        mma_bf16_16816(o[2 * dp + 1], ph, vh[2], vh[3]);
        mma_bf16_16816(o[2 * dp + 1], pl, vh[2], vh[3]);
        mma_bf16_16816(o[2 * dp + 1], ph, vl[2], vl[3]);
      }
    }
    __syncthreads();   // everyone is done with this stage before it is refilled
  }
  // ---- finalize: operand-form output (head h == chunk h)
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
#pragma unroll
  for (int dn = 0; dn < 4; ++dn) {
    uint32_t hi, lo;
    if (r0 < T) {
      split_bf16x2(o[dn][0] * inv0, o[dn][1] * inv0, hi, lo);
      uint8_t* d = out + (size_t)(t0 + r0) * ldo + h * 128 + (dn * 8 + 2 * t) * 2;
      *(uint32_t*)d = hi;
      *(uint32_t*)(d + 64) = lo;
    }
    if (r1 < T) {
      split_bf16x2(o[dn][2] * inv1, o[dn][3] * inv1, hi, lo);
      uint8_t* d = out + (size_t)(t0 + r1) * ldo + h * 128 + (dn * 8 + 2 * t) * 2;
      *(uint32_t*)d = hi;
      *(uint32_t*)(d + 64) = lo;
    }
  }
}

// ---------------------------------------------------------------- box decode / column gather
__global__ void bbox_decode_kernel(const float* __restrict__ raw, int ld_raw, const float* __restrict__ centers, int T,
                                   int with_angle, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  const float* r = raw + (size_t)i * ld_raw;
  float e[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) e[j] = expf(r[j]);
  const float* c = centers + (size_t)i * 3;
  float xc = c[0] + (e[1] - e[0]) / 2.f;
  float yc = c[1] + (e[3] - e[2]) / 2.f;
  float zc = c[2] + (e[5] - e[4]) / 2.f;
  if (!with_angle) {
    float* o = out + (size_t)i * 6;
    o[0] = xc; o[1] = yc; o[2] = zc;
    o[3] = e[0] + e[1]; o[4] = e[2] + e[3]; o[5] = e[4] + e[5];
  } else {
    float scale = e[0] + e[1] + e[2] + e[3];
    float q = expf(sqrtf(r[6] * r[6] + r[7] * r[7]));
    float alpha = 0.5f * atan2f(r[6], r[7]);
    float* o = out + (size_t)i * 7;
    o[0] = xc; o[1] = yc; o[2] = zc;
    o[3] = scale / (1.f + q);
    o[4] = scale / (1.f + q) * q;
    o[5] = e[5] + e[4];
    o[6] = alpha;
  }
}

__global__ void gather_columns_kernel(const float* __restrict__ src, int ld_src, const int32_t* __restrict__ cols, int n_cols,
                                      int T, float* __restrict__ out) {
  long long total = (long long)T * n_cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int t = (int)(i / n_cols), j = (int)(i % n_cols);
    out[i] = src[(size_t)t * ld_src + cols[j]];
  }
}

}  // namespace ud3d

using namespace ud3d;

extern "C" {

size_t ud3d_segmented_mean_workspace_bytes(int n_seg, int C) {
  return (size_t)(n_seg > 0 ? n_seg : 0) * ((size_t)(C > 0 ? C : 0) * 8 + 4);
}

int ud3d_segmented_mean(const float* src, int ld_src, int C, const int32_t* gather, const int64_t* seg, int n, int n_seg,
                        const float* scale, const float* shift, int relu, float* out, void* ws, size_t ws_bytes,
                        void* stream) {
  UD3D_CHECK_ARG(src && seg && out && ws, "ud3d_segmented_mean: NULL argument");
  UD3D_CHECK_ARG(C > 0 && C <= 32 && ld_src >= C && n >= 0 && n_seg >= 0, "ud3d_segmented_mean: need 0 < C <= 32, ld >= C");
  UD3D_CHECK_ARG((scale == nullptr) == (shift == nullptr), "ud3d_segmented_mean: scale/shift must both be set");
  UD3D_CHECK_ARG(((uintptr_t)ws & 7) == 0, "ud3d_segmented_mean: workspace must be 8-byte aligned");
  if (ws_bytes < ud3d_segmented_mean_workspace_bytes(n_seg, C)) {
    set_error("ud3d_segmented_mean: workspace too small");
    return UD3D_EWORKSPACE;
  }
  if (n_seg == 0) return UD3D_OK;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* acc_fix = (unsigned long long*)ws;            // [n_seg, C] fixed-point sums
  int* cnt = (int*)(acc_fix + (size_t)n_seg * C);                   // [n_seg] point counts
  UD3D_CUDA(cudaMemsetAsync(ws, 0, ud3d_segmented_mean_workspace_bytes(n_seg, C), st));
  if (n > 0 && C <= 4) {
    int blocks = cdiv(n, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    segmented_sum_small_kernel<<<blocks, 256, 0, st>>>(src, ld_src, C, gather, seg, n, n_seg, scale, shift, relu, acc_fix, cnt);
    UD3D_LAUNCH_CHECK();
  } else if (n > 0) {
    int warps = cdiv(n, kPoolPtsPerWarp);
    segmented_sum_kernel<<<cdiv(warps, 8), 256, 0, st>>>(src, ld_src, C, gather, seg, n, n_seg, scale, shift, relu, acc_fix, cnt);
    UD3D_LAUNCH_CHECK();
  }
  long long total = (long long)n_seg * C;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  segmented_norm_kernel<<<blocks, 256, 0, st>>>(out, acc_fix, cnt, n_seg, C);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

static int layernorm_impl(const float* x, const float* residual, const float* gamma, const float* beta, float* out,
                          float* out_split, int rows, int C, float eps, void* stream, const char* who) {
  UD3D_CHECK_ARG(x && gamma && beta && (out || out_split), "%s: NULL argument", who);
  UD3D_CHECK_ARG(C > 0 && C % 32 == 0 && C <= 1024 && rows >= 0, "%s: need C %% 32 == 0, C <= 1024", who);
  if (rows == 0) return UD3D_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = cdiv(rows, 8);
  bool vec = (C % 128 == 0) &&
             (((uintptr_t)x | (uintptr_t)out | (uintptr_t)out_split | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)residual) & 15) == 0;
  if (vec && C == 256)
    layernorm_kernel<8><<<blocks, 256, 0, st>>>(x, residual, gamma, beta, out, out_split, rows, C, eps);
  else if (vec && C == 128)
    layernorm_kernel<4><<<blocks, 256, 0, st>>>(x, residual, gamma, beta, out, out_split, rows, C, eps);
  else {
    UD3D_CHECK_ARG(!out_split, "%s: operand-form output needs C in {128, 256} and 16-byte aligned pointers", who);
    layernorm_generic_kernel<<<blocks, 256, 0, st>>>(x, residual, gamma, beta, out, rows, C, eps);
  }
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int ud3d_layernorm(const float* x, const float* residual, const float* gamma, const float* beta, float* out, int rows,
                   int C, float eps, void* stream) {
  UD3D_CHECK_ARG(out, "ud3d_layernorm: NULL out");
  return layernorm_impl(x, residual, gamma, beta, out, nullptr, rows, C, eps, stream, "ud3d_layernorm");
}

int ud3d_layernorm_split(const float* x, const float* residual, const float* gamma, const float* beta, float* out,
                         float* out_split, int rows, int C, float eps, void* stream) {
  return layernorm_impl(x, residual, gamma, beta, out, out_split, rows, C, eps, stream, "ud3d_layernorm_split");
}

static int attention_impl(const float* qkv, const int32_t* cu_seqlens, int B, int max_T, int num_heads, float* out,
                          bool split, void* stream) {
  UD3D_CHECK_ARG(qkv && cu_seqlens && out, "ud3d_attention_fwd: NULL argument");
  UD3D_CHECK_ARG(B > 0 && num_heads > 0 && max_T >= 0, "ud3d_attention_fwd: bad sizes");
  UD3D_CHECK_ARG((((uintptr_t)qkv | (uintptr_t)out) & 15) == 0, "ud3d_attention_fwd: pointers must be 16-byte aligned");
  if (max_T == 0) return UD3D_OK;
  dim3 grid(cdiv(max_T, kAttQ), num_heads, B);
  if (split)
    attention_kernel<true><<<grid, 128, 0, (cudaStream_t)stream>>>(qkv, cu_seqlens, num_heads, out);
  else
    attention_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(qkv, cu_seqlens, num_heads, out);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int ud3d_attention_fwd(const float* qkv, const int32_t* cu_seqlens, int B, int max_T, int num_heads, float* out,
                       void* stream) {
  return attention_impl(qkv, cu_seqlens, B, max_T, num_heads, out, false, stream);
}

int ud3d_attention_fwd_split(const float* qkv, const int32_t* cu_seqlens, int B, int max_T, int num_heads,
                             float* out_split, void* stream) {
  return attention_impl(qkv, cu_seqlens, B, max_T, num_heads, out_split, true, stream);
}

int ud3d_attention_fwd_opform(const float* qkv_split, const int32_t* cu_seqlens, int B, int max_T, int num_heads,
                              float* out_split, void* stream) {
  UD3D_CHECK_ARG(qkv_split && cu_seqlens && out_split, "ud3d_attention_fwd_opform: NULL argument");
  UD3D_CHECK_ARG(B > 0 && num_heads > 0 && max_T >= 0, "ud3d_attention_fwd_opform: bad sizes");
  UD3D_CHECK_ARG((((uintptr_t)qkv_split | (uintptr_t)out_split) & 15) == 0, "ud3d_attention_fwd_opform: pointers must be 16-byte aligned");
  if (max_T == 0) return UD3D_OK;
  dim3 grid(cdiv(max_T, kAtt2Q), num_heads, B);
  attention_opform_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)qkv_split, cu_seqlens, num_heads, (uint8_t*)out_split);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int ud3d_bbox_decode(const float* raw, int ld_raw, const float* centers, int T, int with_angle, float* out, void* stream) {
  UD3D_CHECK_ARG(raw && centers && out && ld_raw >= 8 && T >= 0, "ud3d_bbox_decode: bad argument");
  if (T == 0) return UD3D_OK;
  bbox_decode_kernel<<<cdiv(T, 128), 128, 0, (cudaStream_t)stream>>>(raw, ld_raw, centers, T, with_angle, out);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int ud3d_gather_columns(const float* src, int ld_src, const int32_t* cols, int n_cols, int T, float* out, void* stream) {
  UD3D_CHECK_ARG(src && cols && out && n_cols > 0 && T >= 0, "ud3d_gather_columns: bad argument");
  if (T == 0) return UD3D_OK;
  long long total = (long long)T * n_cols;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  gather_columns_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, ld_src, cols, n_cols, T, out);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

}  // extern "C"
