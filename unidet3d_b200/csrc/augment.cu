// GPU side of the two training-pipeline transforms that sit directly in front of the hot path:
//   ElasticTransfrom  (reference unidet3d/transforms_3d.py:12-83): blur of the noise grids + trilinear displacement
//   PointSample_      (reference unidet3d/transforms_3d.py:233-295): re-indexing of instance / superpoint ids after sampling
// The random draws (noise grids, sample indices) stay with the caller -- with numpy's generator they reproduce the
// reference bit for bit; the arithmetic below follows scipy's (float32 grids, double accumulation, double coordinates).
#include "common.cuh"

namespace ud3d {

// one pass of scipy.ndimage.convolve(n, ones(3)/3 along `axis`, mode='constant', cval=0) over the 3 noise volumes:
// double accumulation w*in[i-1] + w*in[i] + w*in[i+1] (w = float32(1/3)), rounded to float32 per pass like ndimage
__global__ void elastic_blur_pass_kernel(const float* __restrict__ in, float* __restrict__ out, int X, int Y, int Z, int axis) {
  const long long vol = (long long)X * Y * Z;
  const double w = (double)(1.0f / 3.0f);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < 3 * vol; t += (long long)gridDim.x * blockDim.x) {
    const long long r = t % vol;
    const int z = (int)(r % Z), y = (int)((r / Z) % Y), x = (int)(r / ((long long)Z * Y));
    int pos, len;
    long long stride;
    if (axis == 0) { pos = x; len = X; stride = (long long)Y * Z; }
    else if (axis == 1) { pos = y; len = Y; stride = Z; }
    else { pos = z; len = Z; stride = 1; }
    const double a = pos > 0 ? (double)in[t - stride] : 0.0;
    const double b = (double)in[t];
    const double c = pos + 1 < len ? (double)in[t + stride] : 0.0;
    out[t] = (float)(w * a + w * b + w * c);
  }
}

// coords (voxel units) of the reference: float32 division points[:, :3] / voxel_size, then promoted to double
__global__ void voxel_units_kernel(const float* __restrict__ pts, int ld, int n, float voxel_size, double* __restrict__ out) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < 3LL * n; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / 3), k = (int)(t % 3);
    out[t] = (double)__fdiv_rn(pts[(size_t)i * ld + k], voxel_size);
  }
}

// x + RegularGridInterpolator(linspace(-(b-1) gran, (b-1) gran, b), noise_k, linear, fill 0)(x) * mag   per point
__global__ void elastic_apply_kernel(const double* __restrict__ x, int n, const float* __restrict__ noise, int X, int Y, int Z,
                                     double gran, double mag, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int dims[3] = {X, Y, Z};
  const double p[3] = {x[3 * (size_t)i], x[3 * (size_t)i + 1], x[3 * (size_t)i + 2]};
  int idx[3];
  double d[3];
  bool inside = true;
  const double step = 2.0 * gran;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int b = dims[k];
    const double g0 = -(double)(b - 1) * gran;
    const double g1 = (double)(b - 1) * gran;
    inside = inside && p[k] >= g0 && p[k] <= g1;
    int j = (int)floor((p[k] - g0) / step);
    j = j < 0 ? 0 : (j > b - 2 ? b - 2 : j);
    // the interval scipy picks: the largest j with grid[j] <= x (searchsorted), clipped to [0, b - 2]
    while (j > 0 && g0 + step * j > p[k]) --j;
    while (j < b - 2 && g0 + step * (j + 1) <= p[k]) ++j;
    idx[k] = j;
    d[k] = (p[k] - (g0 + step * j)) / ((g0 + step * (j + 1)) - (g0 + step * j));
  }
  const long long vol = (long long)X * Y * Z;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double v = 0.0;
    if (inside) {
      const float* nz = noise + c * vol;
#pragma unroll
      for (int corner = 0; corner < 8; ++corner) {
        const int bx = (corner >> 2) & 1, by = (corner >> 1) & 1, bz = corner & 1;
        double wgt = 1.0;                                   // same multiplication order as scipy's product over dims
        wgt = wgt * (bx ? d[0] : 1.0 - d[0]);
        wgt = wgt * (by ? d[1] : 1.0 - d[1]);
        wgt = wgt * (bz ? d[2] : 1.0 - d[2]);
        v += wgt * (double)nz[((long long)(idx[0] + bx) * Y + (idx[1] + by)) * Z + (idx[2] + bz)];
      }
    }
    out[3 * (size_t)i + c] = p[c] + v * mag;
  }
}

// ---- voxel coordinates of the elastic coordinates: floor(el - el.min(0)) per scene, in DOUBLE like the reference
// (unidet3d.py:162-166 subtracts the per-scene minimum of the float64 `elastic_coords`; MinkowskiEngine floors)
__device__ __forceinline__ unsigned long long dbl_key(double v) {          // order-preserving double -> uint64
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_dbl(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}
__global__ void elastic_min_kernel(const double* __restrict__ el, const int32_t* __restrict__ offs, unsigned long long* __restrict__ mn) {
  const int b = blockIdx.y;
  const int beg = offs[b], end = offs[b + 1];
  unsigned long long m[3] = {~0ull, ~0ull, ~0ull};
  for (int p = beg + blockIdx.x * blockDim.x + threadIdx.x; p < end; p += gridDim.x * blockDim.x) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const unsigned long long k = dbl_key(el[3 * (size_t)p + a]);
      m[a] = k < m[a] ? k : m[a];
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    unsigned long long v = m[a];
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
      v = t < v ? t : v;
    }
    if ((threadIdx.x & 31) == 0 && v != ~0ull) atomicMin(mn + 3 * b + a, v);
  }
}
__global__ void elastic_coords_kernel(const double* __restrict__ el, const int32_t* __restrict__ offs,
                                      const unsigned long long* __restrict__ mn, int32_t* __restrict__ coords, int32_t* max_coord) {
  const int b = blockIdx.y;
  const int beg = offs[b], end = offs[b + 1];
  const double m0 = key_dbl(mn[3 * b]), m1 = key_dbl(mn[3 * b + 1]), m2 = key_dbl(mn[3 * b + 2]);
  int mx[3] = {0, 0, 0};
  for (int p = beg + blockIdx.x * blockDim.x + threadIdx.x; p < end; p += gridDim.x * blockDim.x) {
    const int cx = (int)floor(__dsub_rn(el[3 * (size_t)p], m0));
    const int cy = (int)floor(__dsub_rn(el[3 * (size_t)p + 1], m1));
    const int cz = (int)floor(__dsub_rn(el[3 * (size_t)p + 2], m2));
    reinterpret_cast<int4*>(coords)[p] = make_int4(b, cx, cy, cz);
    mx[0] = max(mx[0], cx); mx[1] = max(mx[1], cy); mx[2] = max(mx[2], cz);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    int v = mx[a];
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(&max_coord[a], v);
  }
}

// ---- dense re-indexing of ids (np.unique(..., return_inverse=True)[1]; negative ids pass through as -1)
__global__ void ids_mark_kernel(const int64_t* __restrict__ ids, int n, int64_t hi, uint32_t* __restrict__ bits, int* __restrict__ bad) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int64_t v = ids[i];
    if (v < 0) continue;
    if (v > hi) { *bad = 1; continue; }
    atomicOr(bits + (v >> 5), 1u << (v & 31));
  }
}
// one CTA: exclusive popcount prefix over the bitmap words; total -> n_unique
__global__ void __launch_bounds__(1024) ids_scan_kernel(const uint32_t* __restrict__ bits, int n_words, uint32_t* __restrict__ prefix,
                                                        int32_t* __restrict__ n_unique) {
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_words; base += 1024) {
    const int w = base + threadIdx.x;
    const uint32_t c = w < n_words ? __popc(bits[w]) : 0u;
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t t = warp_tot[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, t, o);
        if (threadIdx.x >= o) t += u;
      }
      warp_tot[threadIdx.x] = t;
    }
    __syncthreads();
    const uint32_t before = carry + (threadIdx.x >= 32 ? warp_tot[(threadIdx.x >> 5) - 1] : 0u) + inc - c;
    if (w < n_words) prefix[w] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + c;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_unique = (int32_t)carry;
}
__global__ void ids_rank_kernel(const int64_t* __restrict__ ids, int n, int64_t hi, const uint32_t* __restrict__ bits,
                                const uint32_t* __restrict__ prefix, int64_t* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int64_t v = ids[i];
    if (v < 0 || v > hi) { out[i] = -1; continue; }
    const uint32_t word = bits[v >> 5];
    out[i] = (int64_t)(prefix[v >> 5] + __popc(word & ((1u << (v & 31)) - 1u)));
  }
}

}  // namespace ud3d

using namespace ud3d;

extern "C" {

size_t ud3d_elastic_workspace_bytes(const int32_t dims_host[3]) {
  if (!dims_host) return 0;
  return (size_t)3 * dims_host[0] * dims_host[1] * dims_host[2] * 4;
}

int ud3d_elastic_blur(float* noise, const int32_t dims_host[3], void* ws, size_t ws_bytes, void* stream) {
  UD3D_CHECK_ARG(noise && dims_host && ws, "ud3d_elastic_blur: NULL argument");
  const int X = dims_host[0], Y = dims_host[1], Z = dims_host[2];
  UD3D_CHECK_ARG(X > 0 && Y > 0 && Z > 0 && (long long)X * Y * Z * 3 < (1LL << 31), "ud3d_elastic_blur: bad noise grid size");
  if (ws_bytes < ud3d_elastic_workspace_bytes(dims_host)) {
    set_error("ud3d_elastic_blur: workspace too small");
    return UD3D_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = 3LL * X * Y * Z;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  float* a = noise;
  float* b = (float*)ws;
  for (int pass = 0; pass < 6; ++pass) {       // blur0, blur1, blur2, blur0, blur1, blur2 (transforms_3d.py:70-74)
    elastic_blur_pass_kernel<<<blocks, 256, 0, st>>>(a, b, X, Y, Z, pass % 3);
    UD3D_LAUNCH_CHECK();
    float* t = a; a = b; b = t;
  }
  return UD3D_OK;                               // six passes: the result is back in `noise`
}

int ud3d_points_to_voxel_units(const float* points, int ld, int n, float voxel_size, double* out, void* stream) {
  UD3D_CHECK_ARG(points && out && ld >= 3 && n >= 0 && voxel_size > 0.f, "ud3d_points_to_voxel_units: bad argument");
  if (n == 0) return UD3D_OK;
  int blocks = cdiv(3 * n, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  voxel_units_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(points, ld, n, voxel_size, out);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int ud3d_elastic_apply(const double* x, int n, const float* noise, const int32_t dims_host[3], double gran, double mag, double* out,
                       void* stream) {
  UD3D_CHECK_ARG(x && noise && dims_host && out && n >= 0, "ud3d_elastic_apply: NULL argument");
  UD3D_CHECK_ARG(dims_host[0] >= 2 && dims_host[1] >= 2 && dims_host[2] >= 2 && gran > 0.0, "ud3d_elastic_apply: bad noise grid");
  if (n == 0) return UD3D_OK;
  elastic_apply_kernel<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(x, n, noise, dims_host[0], dims_host[1], dims_host[2], gran, mag, out);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

size_t ud3d_elastic_voxel_coords_workspace_bytes(int B) { return (size_t)(B > 0 ? B : 1) * 3 * 8; }

int ud3d_elastic_voxel_coords(const double* elastic, int n, const int32_t* scene_offsets, int B, int32_t* coords, int32_t* max_coord,
                              void* ws, size_t ws_bytes, void* stream) {
  UD3D_CHECK_ARG(elastic && scene_offsets && coords && max_coord && ws, "ud3d_elastic_voxel_coords: NULL argument");
  UD3D_CHECK_ARG(B > 0 && n >= 0 && ((uintptr_t)ws & 7) == 0 && ((uintptr_t)coords & 15) == 0, "ud3d_elastic_voxel_coords: bad B / n / alignment");
  if (ws_bytes < ud3d_elastic_voxel_coords_workspace_bytes(B)) {
    set_error("ud3d_elastic_voxel_coords: workspace too small");
    return UD3D_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  UD3D_CUDA(cudaMemsetAsync(ws, 0xff, ud3d_elastic_voxel_coords_workspace_bytes(B), st));
  UD3D_CUDA(cudaMemsetAsync(max_coord, 0, 12, st));
  if (n == 0) return UD3D_OK;
  int bx = cdiv(cdiv(n, B), 256);
  bx = bx < 1 ? 1 : (bx > 64 ? 64 : bx);
  dim3 grid(bx, B);
  elastic_min_kernel<<<grid, 256, 0, st>>>(elastic, scene_offsets, (unsigned long long*)ws);
  UD3D_LAUNCH_CHECK();
  elastic_coords_kernel<<<grid, 256, 0, st>>>(elastic, scene_offsets, (const unsigned long long*)ws, coords, max_coord);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

size_t ud3d_compact_ids_workspace_bytes(int64_t max_id) {
  const size_t words = (size_t)((max_id < 0 ? 0 : max_id) / 32 + 1);
  return words * 8 + 16;
}

int ud3d_compact_ids(const int64_t* ids, int n, int64_t max_id, int64_t* out, int32_t* n_unique, void* ws, size_t ws_bytes, void* stream) {
  UD3D_CHECK_ARG(ids && out && n_unique && ws && n >= 0, "ud3d_compact_ids: NULL argument");
  UD3D_CHECK_ARG(max_id >= 0 && max_id < (1LL << 31), "ud3d_compact_ids: max_id out of range");
  UD3D_CHECK_ARG(((uintptr_t)ws & 3) == 0, "ud3d_compact_ids: workspace must be 4-byte aligned");
  if (ws_bytes < ud3d_compact_ids_workspace_bytes(max_id)) {
    set_error("ud3d_compact_ids: workspace too small");
    return UD3D_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int n_words = (int)(max_id / 32 + 1);
  uint32_t* bits = (uint32_t*)ws;
  uint32_t* prefix = bits + n_words;
  int* bad = (int*)(prefix + n_words);
  UD3D_CUDA(cudaMemsetAsync(ws, 0, ud3d_compact_ids_workspace_bytes(max_id), st));
  if (n > 0) {
    int blocks = cdiv(n, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    ids_mark_kernel<<<blocks, 256, 0, st>>>(ids, n, max_id, bits, bad);
    UD3D_LAUNCH_CHECK();
  }
  ids_scan_kernel<<<1, 1024, 0, st>>>(bits, n_words, prefix, n_unique);
  UD3D_LAUNCH_CHECK();
  if (n > 0) {
    int blocks = cdiv(n, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    ids_rank_kernel<<<blocks, 256, 0, st>>>(ids, n, max_id, bits, prefix, out);
    UD3D_LAUNCH_CHECK();
  }
  return UD3D_OK;
}

}  // extern "C"
