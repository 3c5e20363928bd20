// Voxelisation, occupancy grid ("perfect spatial hash") and rulebook kernels.
// Integer paths are bit-exact against oracle/voxelize.py and oracle/rulebook.py.
// All kernels here are HBM/L2-latency bound integer work: coalesced, grid-stride, no tensor cores.
#include "common.cuh"

#include <float.h>
#include <atomic>
#include <map>
#include <mutex>

namespace ud3d {

// ---------------------------------------------------------------- error / counters (library-wide)
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};          // process-wide: batches may be issued from several host threads
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---------------------------------------------------------------- per-device context
struct DeviceCtx {
  int device = -1;
  int sms = 0;
  std::mutex mu;
  struct Cfg { size_t smem = 0; int carve = -2; };
  std::map<const void*, Cfg> configured;      // kernel -> what its function attributes were last set to on this device
};
static std::mutex g_ctx_mu;
static DeviceCtx* g_ctx[256] = {nullptr};

DeviceCtx* device_ctx(int* device_out) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 256) {
    set_error("ud3d: cannot query the current CUDA device");
    return nullptr;
  }
  if (device_out) *device_out = dev;
  std::lock_guard<std::mutex> lock(g_ctx_mu);
  if (!g_ctx[dev]) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
      set_error("ud3d: cannot query the SM count of device %d", dev);
      return nullptr;
    }
    DeviceCtx* c = new DeviceCtx();
    c->device = dev;
    c->sms = sms;
    g_ctx[dev] = c;
  }
  return g_ctx[dev];
}
int ctx_sm_count(const DeviceCtx* c) { return c ? c->sms : 0; }
void ctx_lock(DeviceCtx* c) { c->mu.lock(); }
void ctx_unlock(DeviceCtx* c) { c->mu.unlock(); }
bool ctx_needs_config(DeviceCtx* c, const void* key, size_t smem, int carve) {      // caller holds the lock
  DeviceCtx::Cfg& f = c->configured[key];
  const bool need = smem > f.smem || carve != f.carve;
  if (smem > f.smem) f.smem = smem;
  f.carve = carve;
  return need;
}

// ---------------------------------------------------------------- point statistics + coords
// stats layout per scene: [min x,y,z (as ordered float), sum x,y,z (double)] in workspace
struct SceneAcc {
  float mn[3];
  int pad;
  double sum[3];
  double pad2;
};

__global__ void init_scene_acc(SceneAcc* acc, int B, int32_t* max_coord) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    for (int a = 0; a < 3; ++a) {
      acc[i].mn[a] = FLT_MAX;
      acc[i].sum[a] = 0.0;
    }
  }
  if (i < 3) max_coord[i] = 0;
}

__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
  // valid for any mix of signs when the cell starts at +FLT_MAX
  if (v >= 0.f)
    atomicMin((int*)addr, __float_as_int(v));
  else
    atomicMax((unsigned int*)addr, __float_as_uint(v));
}

// grid: (chunks, B); block 256
__global__ void scene_stats_kernel(const float* __restrict__ pts, const int32_t* __restrict__ offs, SceneAcc* acc) {
  int b = blockIdx.y;
  int beg = offs[b], end = offs[b + 1];
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
  double sm[3] = {0, 0, 0};
  for (int p = beg + blockIdx.x * blockDim.x + threadIdx.x; p < end; p += gridDim.x * blockDim.x) {
    const float* q = pts + (size_t)p * 6;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float v = q[a];
      mn[a] = fminf(mn[a], v);
      sm[a] += (double)v;
    }
  }
  __shared__ float s_mn[8][3];
  __shared__ double s_sm[8][3];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      sm[a] += __shfl_xor_sync(0xffffffffu, sm[a], o);
    }
    if (lane == 0) {
      s_mn[w][a] = mn[a];
      s_sm[w][a] = sm[a];
    }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    int a = threadIdx.x;
    float m = FLT_MAX;
    double s = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
      m = fminf(m, s_mn[i][a]);
      s += s_sm[i][a];
    }
    if (end > beg) {
      atomic_min_float(&acc[b].mn[a], m);
      atomicAdd(&acc[b].sum[a], s);
    }
  }
}

__global__ void point_coords_kernel(const float* __restrict__ pts, const int32_t* __restrict__ offs,
                                    const SceneAcc* __restrict__ acc, float voxel_size, int32_t* __restrict__ coords,
                                    float* __restrict__ feats, float* __restrict__ stats, int32_t* max_coord) {
  int b = blockIdx.y;
  int beg = offs[b], end = offs[b + 1];
  int n = end - beg;
  float mn[3], mean[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    mn[a] = acc[b].mn[a];
    mean[a] = n > 0 ? (float)(acc[b].sum[a] / (double)n) : 0.f;
  }
  if (blockIdx.x == 0 && threadIdx.x < 6) stats[b * 6 + threadIdx.x] = threadIdx.x < 3 ? mn[threadIdx.x] : mean[threadIdx.x - 3];
  int mx[3] = {0, 0, 0};
  for (int p = beg + blockIdx.x * blockDim.x + threadIdx.x; p < end; p += gridDim.x * blockDim.x) {
    const float* q = pts + (size_t)p * 6;
    float x = q[0], y = q[1], z = q[2];
    int cx = (int)floorf(__fdiv_rn(__fsub_rn(x, mn[0]), voxel_size));
    int cy = (int)floorf(__fdiv_rn(__fsub_rn(y, mn[1]), voxel_size));
    int cz = (int)floorf(__fdiv_rn(__fsub_rn(z, mn[2]), voxel_size));
    reinterpret_cast<int4*>(coords)[p] = make_int4(b, cx, cy, cz);
    float* f = feats + (size_t)p * 6;
    f[0] = q[3]; f[1] = q[4]; f[2] = q[5];
    f[3] = __fsub_rn(x, mean[0]); f[4] = __fsub_rn(y, mean[1]); f[5] = __fsub_rn(z, mean[2]);
    mx[0] = max(mx[0], cx); mx[1] = max(mx[1], cy); mx[2] = max(mx[2], cz);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    int v = mx[a];
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(&max_coord[a], v);
  }
}

// ---------------------------------------------------------------- grid build
__global__ void grid_set_bits_kernel(const int4* __restrict__ coords, int n, GridDims g, uint32_t* words) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int4 c = coords[i];
    if (c.x < 0) continue;
    if ((unsigned)c.x >= (unsigned)g.B || (unsigned)c.y >= (unsigned)g.X || (unsigned)c.z >= (unsigned)g.Y ||
        (unsigned)c.w >= (unsigned)g.Z)
      continue;  // outside the declared box: ignored (caller sized the box from max_coord)
    uint32_t* w = words + grid_word_index(g, c.x, c.y, c.z, c.w);
    const uint32_t bit = 1u << (c.w & 31);
    // many inputs map to the same cell (points -> voxels -> coarse ancestors): test before the atomic
    if (!(*(volatile uint32_t*)w & bit)) atomicOr(w, bit);
  }
}

// coarse occupancy bitmap of a k=2,s=2 down-sampling straight from the finer bitmap: cell (b,x,y,z) -> (b,x/2,y/2,z/2);
// coordinates outside the coarse box are dropped (spconv's odd-extent rule: out_shape = (in_shape-2)/2+1, the coarse
// box is min(out_shape, (extent+1)/2)).  One thread per fine word: the 32 z-cells of the word fold into 16 coarse
// z-cells = one half of a coarse word -> at most one atomicOr per non-empty fine word.
__global__ void grid_downsample_kernel(const uint32_t* __restrict__ fine, GridDims gf, uint32_t* coarse, GridDims gc) {
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < gf.nwords; w += (long long)gridDim.x * blockDim.x) {
    uint32_t bits = fine[w];
    if (!bits) continue;
    long long t = w;
    const int zw = (int)(t % gf.Zw); t /= gf.Zw;
    const int y = (int)(t % gf.Y); t /= gf.Y;
    const int x = (int)(t % gf.X);
    const int b = (int)(t / gf.X);
    const int cx = x >> 1, cy = y >> 1;
    if (cx >= gc.X || cy >= gc.Y) continue;
    uint32_t m = (bits | (bits >> 1)) & 0x55555555u;       // coarse cell i <- fine cells 2i, 2i+1 (at bit 2i)
    m = (m | (m >> 1)) & 0x33333333u;
    m = (m | (m >> 2)) & 0x0f0f0f0fu;
    m = (m | (m >> 4)) & 0x00ff00ffu;
    m = (m | (m >> 8)) & 0x0000ffffu;                       // 16 coarse cells, z = zw * 16 + i
    const int cz0 = zw * 16;
    if (cz0 >= gc.Z) continue;
    if (cz0 + 16 > gc.Z) m &= (1u << (gc.Z - cz0)) - 1u;    // coarse z >= extent: dropped
    if (!m) continue;
    uint32_t* dst = coarse + grid_word_index(gc, b, cx, cy, cz0);
    const uint32_t val = m << (cz0 & 31);
    if ((*(volatile uint32_t*)dst & val) != val) atomicOr(dst, val);
  }
}

// exclusive scan of popc(words): phase 1 block sums
__global__ void __launch_bounds__(256) grid_scan_phase1(const uint32_t* __restrict__ words, long long nwords, uint32_t* bsum) {
  long long base = (long long)blockIdx.x * kScanBlockWords;
  uint32_t s = 0;
  for (int j = threadIdx.x; j < kScanBlockWords; j += 256) {
    long long w = base + j;
    if (w < nwords) s += __popc(words[w]);
  }
  __shared__ uint32_t sh[8];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int i = 0; i < 8; ++i) t += sh[i];
    bsum[blockIdx.x] = t;
  }
}
// phase 2: single CTA exclusive scan of block sums (in place), total -> bsum[nblocks] and *n_unique
__global__ void __launch_bounds__(1024) grid_scan_phase2(uint32_t* bsum, int nblocks, int32_t* n_unique) {
  __shared__ uint32_t sh[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    int i = base + threadIdx.x;
    uint32_t v = i < nblocks ? bsum[i] : 0;
    uint32_t x = v;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) sh[w] = x;
    __syncthreads();
    if (w == 0) {
      uint32_t t = sh[lane];
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      sh[lane] = t;
    }
    __syncthreads();
    uint32_t incl = x + (w > 0 ? sh[w - 1] : 0) + carry;
    if (i < nblocks) bsum[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    bsum[nblocks] = carry;
    if (n_unique) *n_unique = (int32_t)carry;
  }
}
// phase 3: per-word exclusive prefix
__global__ void __launch_bounds__(256) grid_scan_phase3(const uint32_t* __restrict__ words, long long nwords,
                                                        const uint32_t* __restrict__ bsum, uint32_t* __restrict__ prefix) {
  // each thread owns 16 consecutive words
  long long base = (long long)blockIdx.x * kScanBlockWords + threadIdx.x * 16;
  uint32_t cnt[16];
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    long long w = base + j;
    cnt[j] = w < nwords ? __popc(words[w]) : 0;
    s += cnt[j];
  }
  __shared__ uint32_t sh[8];
  int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  uint32_t x = s;
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) sh[wp] = x;
  __syncthreads();
  uint32_t off = bsum[blockIdx.x];
  for (int i = 0; i < wp; ++i) off += sh[i];
  uint32_t run = off + x - s;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    long long w = base + j;
    if (w < nwords) prefix[w] = run;
    run += cnt[j];
  }
}

__global__ void grid_rank_kernel(const int4* __restrict__ coords, int n, GridDims g, const uint32_t* __restrict__ words,
                                 const uint32_t* __restrict__ prefix, int32_t* __restrict__ rank) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int4 c = coords[i];
    rank[i] = c.x < 0 ? -1 : grid_rank(g, words, prefix, c.x, c.y, c.z, c.w);
  }
}

__global__ void grid_coords_kernel(GridDims g, const uint32_t* __restrict__ words, const uint32_t* __restrict__ prefix,
                                   int4* __restrict__ out, int n_unique) {
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < g.nwords; w += (long long)gridDim.x * blockDim.x) {
    uint32_t bits = words[w];
    if (!bits) continue;
    int zw = (int)(w % g.Zw);
    long long col = w / g.Zw;
    int y = (int)(col % g.Y);
    long long t = col / g.Y;
    int x = (int)(t % g.X);
    int b = (int)(t / g.X);
    uint32_t r = prefix[w];
    while (bits) {
      int bit = __ffs(bits) - 1;
      bits &= bits - 1;
      if ((int)r < n_unique) out[r] = make_int4(b, x, y, zw * 32 + bit);
      ++r;
    }
  }
}

// ---------------------------------------------------------------- voxel feature mean
__global__ void voxel_accum_kernel(const float* __restrict__ feats, const int32_t* __restrict__ rank, int n, int C,
                                   float* out, float* cnt) {
  long long total = (long long)n * C;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int p = (int)(t / C), c = (int)(t % C);
    int r = rank[p];
    if (r < 0) continue;
    atomicAdd(out + (size_t)r * C + c, feats[t]);
    if (c == 0) atomicAdd(cnt + r, 1.0f);
  }
}
__global__ void voxel_norm_kernel(float* out, const float* __restrict__ cnt, int n_vox, int C) {
  long long total = (long long)n_vox * C;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    float c = cnt[t / C];
    out[t] = __fdiv_rn(out[t], fmaxf(c, 1.0f));
  }
}

// ---------------------------------------------------------------- rulebooks
__global__ void row_of_rank_kernel(const int4* __restrict__ coords, int n, GridDims g, const uint32_t* __restrict__ words,
                                   const uint32_t* __restrict__ prefix, int32_t* __restrict__ row_of_rank) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int4 c = coords[i];
    int r = grid_rank(g, words, prefix, c.x, c.y, c.z, c.w);
    if (r >= 0) row_of_rank[r] = i;
  }
}

// one thread per voxel, 27 lookups; table is offset-major so writes are coalesced
__global__ void __launch_bounds__(128) subm3_table_kernel(const int4* __restrict__ coords, int n, GridDims g,
                                                          const uint32_t* __restrict__ words,
                                                          const uint32_t* __restrict__ prefix,
                                                          const int32_t* __restrict__ row_of_rank,
                                                          int32_t* __restrict__ table, uint32_t* tile_mask) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t mask = 0;
  if (i < n) {
    int4 c = coords[i];
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      int dx = k / 9 - 1, dy = (k / 3) % 3 - 1, dz = k % 3 - 1;
      int r = grid_rank(g, words, prefix, c.x, c.y + dx, c.z + dy, c.w + dz);
      if (r >= 0 && row_of_rank) r = row_of_rank[r];
      table[(size_t)k * n + i] = r;
      if (r >= 0) mask |= 1u << k;
    }
  }
  if (tile_mask) {
    mask = __reduce_or_sync(0xffffffffu, mask);
    // blockDim == UD3D_TILE_M == 128: one tile per CTA
    if ((threadIdx.x & 31) == 0 && mask) atomicOr(tile_mask + blockIdx.x, mask);
  }
}

__global__ void down2_parents_kernel(const int4* __restrict__ coords, int n, int ox, int oy, int oz, int4* __restrict__ parents) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int4 c = coords[i];
    int px = c.y >> 1, py = c.z >> 1, pz = c.w >> 1;
    bool keep = c.x >= 0 && px < ox && py < oy && pz < oz;
    parents[i] = make_int4(keep ? c.x : -1, px, py, pz);
  }
}

// ancestor of every input voxel after `levels` successive k2/s2 down-samplings, with the per-level drop rule
// (coordinate/2 >= out_shape of that level -> the voxel and all its descendants contribute nothing deeper).
// Lets the grids (and counts) of ALL coarse levels be built from the finest coordinates in one go, without
// knowing the intermediate voxel counts on the host.
__global__ void down_ancestors_kernel(const int4* __restrict__ coords, int n, int sx, int sy, int sz, int levels,
                                      int4* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int4 c = coords[i];
    int x = c.y, y = c.z, z = c.w, ax = sx, ay = sy, az = sz;
    bool keep = c.x >= 0;
    for (int l = 0; l < levels && keep; ++l) {
      const int ox = (ax - 2) / 2 + 1, oy = (ay - 2) / 2 + 1, oz = (az - 2) / 2 + 1;
      x >>= 1; y >>= 1; z >>= 1;
      keep = x < ox && y < oy && z < oz;
      ax = ox; ay = oy; az = oz;
    }
    out[i] = make_int4(keep ? c.x : -1, x, y, z);
  }
}

__global__ void down2_fill_kernel(const int4* __restrict__ coords, const int4* __restrict__ parents, int n_fine, int n_coarse,
                                  GridDims g, const uint32_t* __restrict__ words, const uint32_t* __restrict__ prefix,
                                  int32_t* __restrict__ child, int32_t* __restrict__ up, uint32_t* child_mask,
                                  uint32_t* up_mask) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_fine; i += gridDim.x * blockDim.x) {
    int4 p = parents[i];
    if (p.x < 0) continue;
    int4 c = coords[i];
    int r = grid_rank(g, words, prefix, p.x, p.y, p.z, p.w);
    if (r < 0 || r >= n_coarse) continue;
    int s = (c.y & 1) * 4 + (c.z & 1) * 2 + (c.w & 1);
    child[(size_t)s * n_coarse + r] = i;
    up[(size_t)s * n_fine + i] = r;
    if (child_mask) atomicOr(child_mask + r / UD3D_TILE_M, 1u << s);
    if (up_mask) atomicOr(up_mask + i / UD3D_TILE_M, 1u << s);
  }
}


// ---------------------------------------------------------------- SubM3 tile order
// The gather-GEMM skips a kernel offset for a whole 128-row tile only when NO row of the tile has that neighbour.  In
// canonical (b,x,y,z) order nearly every tile sees all 27 offsets although a voxel has ~11-14 neighbours.  Rows are
// therefore regrouped by their neighbourhood pattern: a 16-bit key of the rarest offsets (8 corners, then the 8 edges
// with dx = 0 or dy = 0), counting-sorted (histogram -> scan -> scatter).  The order inside a bucket is whatever the
// atomics produce: it only changes which rows share a tile, never a result (every output row accumulates its own
// offsets in ascending k).  Measured on ScanNet-like scenes: 26.2 -> 19.6 active offsets per tile at 2 cm, 25.4 -> 16.5
// at 4 cm.
constexpr int kOrderKeyBits = 16;
__device__ __constant__ int8_t kOrderKeyOffsets[kOrderKeyBits] = {0, 2, 6, 8, 18, 20, 24, 26, 9, 11, 15, 17, 3, 5, 21, 23};

__device__ __forceinline__ uint32_t order_key(const int32_t* __restrict__ table, int n, int i) {
  uint32_t key = 0;
#pragma unroll
  for (int b = 0; b < kOrderKeyBits; ++b) key = (key << 1) | (table[(size_t)kOrderKeyOffsets[b] * n + i] >= 0 ? 1u : 0u);
  return key;
}

// warp-aggregated atomicAdd on counters[key]: returns this lane's slot
__device__ __forceinline__ uint32_t bucket_take(uint32_t* counters, uint32_t key, bool active) {
  const uint32_t act = __ballot_sync(0xffffffffu, active);
  uint32_t slot = 0;
  if (active) {
    const uint32_t peers = __match_any_sync(act, key);
    const int leader = __ffs(peers) - 1;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(counters + key, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    slot = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
  }
  return slot;
}

__global__ void __launch_bounds__(256) order_hist_kernel(const int32_t* __restrict__ table, int n, uint16_t* __restrict__ keys,
                                                         uint32_t* __restrict__ hist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < n;
  uint32_t key = 0;
  if (active) {
    key = order_key(table, n, i);
    keys[i] = (uint16_t)key;
  }
  bucket_take(hist, key, active);
}

// exclusive scan of the 65536 bucket counts (one CTA, 1024 threads x 64 consecutive buckets)
__global__ void __launch_bounds__(1024) order_scan_kernel(uint32_t* __restrict__ hist) {
  __shared__ uint32_t warp_sums[32];
  constexpr int kPer = (1 << kOrderKeyBits) / 1024;
  const int t = threadIdx.x;
  uint32_t v[kPer];
  uint32_t sum = 0;
  const uint4* src = (const uint4*)(hist + (size_t)t * kPer);
#pragma unroll
  for (int j = 0; j < kPer / 4; ++j) {
    uint4 q = src[j];
    v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
    sum += q.x + q.y + q.z + q.w;
  }
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
    if ((t & 31) >= o) incl += u;
  }
  if ((t & 31) == 31) warp_sums[t >> 5] = incl;
  __syncthreads();
  if (t < 32) {
    uint32_t w = warp_sums[t], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t u = __shfl_up_sync(0xffffffffu, wi, o);
      if (t >= o) wi += u;
    }
    warp_sums[t] = wi - w;
  }
  __syncthreads();
  uint32_t run = warp_sums[t >> 5] + incl - sum;
  uint4* dst = (uint4*)(hist + (size_t)t * kPer);
#pragma unroll
  for (int j = 0; j < kPer / 4; ++j) {
    uint4 q;
    q.x = run; run += v[4 * j];
    q.y = run; run += v[4 * j + 1];
    q.z = run; run += v[4 * j + 2];
    q.w = run; run += v[4 * j + 3];
    dst[j] = q;
  }
}

__global__ void __launch_bounds__(256) order_scatter_kernel(const uint16_t* __restrict__ keys, int n, uint32_t* __restrict__ cursor,
                                                            int32_t* __restrict__ perm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < n;
  const uint32_t key = active ? keys[i] : 0u;
  const uint32_t slot = bucket_take(cursor, key, active);
  if (active) perm[slot] = i;
}

// table_p[k][i] = table[k][perm[i]] and the tile mask of the regrouped rows (one CTA per 128-row tile)
__global__ void __launch_bounds__(UD3D_TILE_M) order_permute_kernel(const int32_t* __restrict__ table, const int32_t* __restrict__ perm,
                                                                    int n, int32_t* __restrict__ table_p, uint32_t* __restrict__ tile_mask_p) {
  __shared__ uint32_t s_mask[UD3D_TILE_M / 32];
  const int i = blockIdx.x * UD3D_TILE_M + threadIdx.x;
  uint32_t mask = 0;
  if (i < n) {
    const int r = perm[i];
    int32_t v[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) v[k] = __ldg(table + (size_t)k * n + r);
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      table_p[(size_t)k * n + i] = v[k];
      if (v[k] >= 0) mask |= 1u << k;
    }
  }
  mask = __reduce_or_sync(0xffffffffu, mask);
  if ((threadIdx.x & 31) == 0) s_mask[threadIdx.x >> 5] = mask;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t m = 0;
#pragma unroll
    for (int w = 0; w < UD3D_TILE_M / 32; ++w) m |= s_mask[w];
    tile_mask_p[blockIdx.x] = m;
  }
}

static int grid_blocks(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > 148 * 16) b = 148 * 16;
  return (int)b;
}

}  // namespace ud3d

using namespace ud3d;

extern "C" {

int ud3d_version(void) { return 100; }
const ud3d_ctx* ud3d_ctx_current(void) { return (const ud3d_ctx*)device_ctx(); }
int ud3d_ctx_device(const ud3d_ctx* ctx) { return ctx ? ((const DeviceCtx*)ctx)->device : -1; }
int ud3d_ctx_sm_count(const ud3d_ctx* ctx) { return ctx_sm_count((const DeviceCtx*)ctx); }
const char* ud3d_last_error(void) { return g_err; }
int64_t ud3d_launch_count(int reset) {
  return reset ? g_launches.exchange(0, std::memory_order_relaxed) : g_launches.load(std::memory_order_relaxed);
}

size_t ud3d_point_coords_workspace_bytes(int B) { return align_up(sizeof(SceneAcc) * (size_t)(B > 0 ? B : 1), 256); }

int ud3d_point_coords(const float* points, int n, const int32_t* scene_offsets, int B, float voxel_size,
                      int32_t* coords, float* feats, float* stats, int32_t* max_coord, void* ws, size_t ws_bytes,
                      void* stream) {
  UD3D_CHECK_ARG(points && scene_offsets && coords && feats && stats && max_coord && ws, "ud3d_point_coords: NULL argument");
  UD3D_CHECK_ARG(B > 0 && n >= 0 && voxel_size > 0.f, "ud3d_point_coords: bad B/n/voxel_size");
  if (ws_bytes < ud3d_point_coords_workspace_bytes(B)) {
    set_error("ud3d_point_coords: workspace too small");
    return UD3D_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  SceneAcc* acc = (SceneAcc*)ws;
  init_scene_acc<<<cdiv(B > 3 ? B : 3, 128), 128, 0, st>>>(acc, B, max_coord);
  UD3D_LAUNCH_CHECK();
  if (n == 0) return UD3D_OK;
  int per_scene_blocks = grid_blocks((n + B - 1) / B, 256);
  if (per_scene_blocks > 296) per_scene_blocks = 296;
  dim3 grid(per_scene_blocks, B);
  scene_stats_kernel<<<grid, 256, 0, st>>>(points, scene_offsets, acc);
  UD3D_LAUNCH_CHECK();
  point_coords_kernel<<<grid, 256, 0, st>>>(points, scene_offsets, acc, voxel_size, coords, feats, stats, max_coord);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

size_t ud3d_grid_workspace_bytes(const int32_t dims_host[4]) {
  if (!dims_host) return 0;
  GridDims g = make_grid_dims(dims_host);
  return grid_ws_bytes(g);
}

static int check_dims(const int32_t d[4], const char* who) {
  UD3D_CHECK_ARG(d && d[0] > 0 && d[1] > 0 && d[2] > 0 && d[3] > 0, "%s: bad grid dims", who);
  GridDims g = make_grid_dims(d);
  UD3D_CHECK_ARG(g.nwords < (1ll << 31), "%s: grid of %lld words is too large", who, g.nwords);
  return UD3D_OK;
}

int ud3d_grid_build(const int32_t* coords, int n, const int32_t dims_host[4], void* ws, size_t ws_bytes,
                    int32_t* n_unique, void* stream) {
  int rc = check_dims(dims_host, "ud3d_grid_build");
  if (rc) return rc;
  UD3D_CHECK_ARG(ws && (coords || n == 0), "ud3d_grid_build: NULL argument");
  UD3D_CHECK_ARG(((uintptr_t)coords & 15) == 0, "ud3d_grid_build: coords must be 16-byte aligned");
  GridDims g = make_grid_dims(dims_host);
  if (ws_bytes < grid_ws_bytes(g)) {
    set_error("ud3d_grid_build: workspace too small (%zu < %zu)", ws_bytes, grid_ws_bytes(g));
    return UD3D_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  GridView v = grid_view(g, ws);
  UD3D_CUDA(cudaMemsetAsync(v.words, 0, (size_t)g.nwords * 4, st));
  if (n > 0) {
    grid_set_bits_kernel<<<grid_blocks(n, 256), 256, 0, st>>>((const int4*)coords, n, g, v.words);
    UD3D_LAUNCH_CHECK();
  }
  grid_scan_phase1<<<v.nblocks, 256, 0, st>>>(v.words, g.nwords, v.bsum);
  UD3D_LAUNCH_CHECK();
  grid_scan_phase2<<<1, 1024, 0, st>>>(v.bsum, v.nblocks, n_unique);
  UD3D_LAUNCH_CHECK();
  grid_scan_phase3<<<v.nblocks, 256, 0, st>>>(v.words, g.nwords, v.bsum, v.prefix);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int ud3d_grid_build_coarser(const int32_t fine_dims_host[4], const void* fine_ws, const int32_t dims_host[4], void* ws,
                            size_t ws_bytes, int32_t* n_unique, void* stream) {
  int rc = check_dims(dims_host, "ud3d_grid_build_coarser");
  if (rc) return rc;
  rc = check_dims(fine_dims_host, "ud3d_grid_build_coarser");
  if (rc) return rc;
  UD3D_CHECK_ARG(ws && fine_ws, "ud3d_grid_build_coarser: NULL argument");
  UD3D_CHECK_ARG(dims_host[0] == fine_dims_host[0], "ud3d_grid_build_coarser: batch sizes differ");
  GridDims gf = make_grid_dims(fine_dims_host), g = make_grid_dims(dims_host);
  if (ws_bytes < grid_ws_bytes(g)) {
    set_error("ud3d_grid_build_coarser: workspace too small (%zu < %zu)", ws_bytes, grid_ws_bytes(g));
    return UD3D_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  GridView vf = grid_view(gf, fine_ws), v = grid_view(g, ws);
  UD3D_CUDA(cudaMemsetAsync(v.words, 0, (size_t)g.nwords * 4, st));
  grid_downsample_kernel<<<grid_blocks(gf.nwords, 256), 256, 0, st>>>(vf.words, gf, v.words, g);
  UD3D_LAUNCH_CHECK();
  grid_scan_phase1<<<v.nblocks, 256, 0, st>>>(v.words, g.nwords, v.bsum);
  UD3D_LAUNCH_CHECK();
  grid_scan_phase2<<<1, 1024, 0, st>>>(v.bsum, v.nblocks, n_unique);
  UD3D_LAUNCH_CHECK();
  grid_scan_phase3<<<v.nblocks, 256, 0, st>>>(v.words, g.nwords, v.bsum, v.prefix);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int ud3d_grid_rank(const int32_t* coords, int n, const int32_t dims_host[4], const void* ws, int32_t* rank_out,
                   void* stream) {
  int rc = check_dims(dims_host, "ud3d_grid_rank");
  if (rc) return rc;
  UD3D_CHECK_ARG(ws && (n == 0 || (coords && rank_out)), "ud3d_grid_rank: NULL argument");
  if (n == 0) return UD3D_OK;
  GridDims g = make_grid_dims(dims_host);
  GridView v = grid_view(g, ws);
  grid_rank_kernel<<<grid_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>((const int4*)coords, n, g, v.words, v.prefix, rank_out);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int ud3d_grid_coords(const int32_t dims_host[4], const void* ws, int32_t* coords_out, int n_unique, void* stream) {
  int rc = check_dims(dims_host, "ud3d_grid_coords");
  if (rc) return rc;
  UD3D_CHECK_ARG(ws && (n_unique == 0 || coords_out), "ud3d_grid_coords: NULL argument");
  if (n_unique == 0) return UD3D_OK;
  GridDims g = make_grid_dims(dims_host);
  GridView v = grid_view(g, ws);
  grid_coords_kernel<<<grid_blocks(g.nwords, 256), 256, 0, (cudaStream_t)stream>>>(g, v.words, v.prefix, (int4*)coords_out, n_unique);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int ud3d_voxel_mean(const float* feats_pts, const int32_t* rank, int n, int C, int n_vox, float* out, void* ws,
                    size_t ws_bytes, void* stream) {
  UD3D_CHECK_ARG(feats_pts && rank && out && ws && C > 0 && n >= 0 && n_vox >= 0, "ud3d_voxel_mean: bad argument");
  if (ws_bytes < (size_t)n_vox * 4) {
    set_error("ud3d_voxel_mean: workspace too small");
    return UD3D_EWORKSPACE;
  }
  if (n_vox == 0) return UD3D_OK;
  cudaStream_t st = (cudaStream_t)stream;
  UD3D_CUDA(cudaMemsetAsync(out, 0, (size_t)n_vox * C * 4, st));
  UD3D_CUDA(cudaMemsetAsync(ws, 0, (size_t)n_vox * 4, st));
  if (n > 0) {
    voxel_accum_kernel<<<grid_blocks((long long)n * C, 256), 256, 0, st>>>(feats_pts, rank, n, C, out, (float*)ws);
    UD3D_LAUNCH_CHECK();
  }
  voxel_norm_kernel<<<grid_blocks((long long)n_vox * C, 256), 256, 0, st>>>(out, (const float*)ws, n_vox, C);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int ud3d_rulebook_subm3(const int32_t* coords, int n, const int32_t dims_host[4], const void* ws, int32_t* row_of_rank,
                        int32_t* table, uint32_t* tile_mask, void* stream) {
  int rc = check_dims(dims_host, "ud3d_rulebook_subm3");
  if (rc) return rc;
  UD3D_CHECK_ARG(ws && (n == 0 || (coords && table)), "ud3d_rulebook_subm3: NULL argument");
  if (n == 0) return UD3D_OK;
  cudaStream_t st = (cudaStream_t)stream;
  GridDims g = make_grid_dims(dims_host);
  GridView v = grid_view(g, ws);
  if (row_of_rank) {
    row_of_rank_kernel<<<grid_blocks(n, 256), 256, 0, st>>>((const int4*)coords, n, g, v.words, v.prefix, row_of_rank);
    UD3D_LAUNCH_CHECK();
  }
  int tiles = cdiv(n, UD3D_TILE_M);
  if (tile_mask) UD3D_CUDA(cudaMemsetAsync(tile_mask, 0, (size_t)tiles * 4, st));
  subm3_table_kernel<<<tiles, UD3D_TILE_M, 0, st>>>((const int4*)coords, n, g, v.words, v.prefix, row_of_rank, table, tile_mask);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

size_t ud3d_subm3_tile_order_workspace_bytes(int n) {
  return align_up((size_t)(1 << kOrderKeyBits) * 4, 256) + align_up((size_t)(n > 0 ? n : 1) * 2, 256);
}

int ud3d_subm3_tile_order(const int32_t* table, int n, int32_t* perm, int32_t* table_p, uint32_t* tile_mask_p, void* ws,
                          size_t ws_bytes, void* stream) {
  UD3D_CHECK_ARG(n >= 0 && (n == 0 || (table && perm && table_p && tile_mask_p && ws)), "ud3d_subm3_tile_order: NULL argument");
  UD3D_CHECK_ARG(ws_bytes >= ud3d_subm3_tile_order_workspace_bytes(n), "ud3d_subm3_tile_order: workspace too small");
  UD3D_CHECK_ARG(((uintptr_t)ws & 15) == 0, "ud3d_subm3_tile_order: workspace must be 16-byte aligned");
  if (n == 0) return UD3D_OK;
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* hist = (uint32_t*)ws;
  uint16_t* keys = (uint16_t*)((uint8_t*)ws + align_up((size_t)(1 << kOrderKeyBits) * 4, 256));
  UD3D_CUDA(cudaMemsetAsync(hist, 0, (size_t)(1 << kOrderKeyBits) * 4, st));
  const int blocks = cdiv(n, 256);
  order_hist_kernel<<<blocks, 256, 0, st>>>(table, n, keys, hist);
  UD3D_LAUNCH_CHECK();
  order_scan_kernel<<<1, 1024, 0, st>>>(hist);
  UD3D_LAUNCH_CHECK();
  order_scatter_kernel<<<blocks, 256, 0, st>>>(keys, n, hist, perm);
  UD3D_LAUNCH_CHECK();
  order_permute_kernel<<<cdiv(n, UD3D_TILE_M), UD3D_TILE_M, 0, st>>>(table, perm, n, table_p, tile_mask_p);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int ud3d_down2_parents(const int32_t* coords, int n, const int32_t in_shape_host[3], int32_t* parents, void* stream) {
  UD3D_CHECK_ARG(in_shape_host && (n == 0 || (coords && parents)), "ud3d_down2_parents: NULL argument");
  if (n == 0) return UD3D_OK;
  int ox = (in_shape_host[0] - 2) / 2 + 1, oy = (in_shape_host[1] - 2) / 2 + 1, oz = (in_shape_host[2] - 2) / 2 + 1;
  down2_parents_kernel<<<grid_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>((const int4*)coords, n, ox, oy, oz, (int4*)parents);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int ud3d_down_ancestors(const int32_t* coords, int n, const int32_t in_shape_host[3], int levels, int32_t* ancestors,
                        void* stream) {
  UD3D_CHECK_ARG(in_shape_host && levels >= 1 && (n == 0 || (coords && ancestors)), "ud3d_down_ancestors: bad argument");
  if (n == 0) return UD3D_OK;
  down_ancestors_kernel<<<grid_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>((const int4*)coords, n, in_shape_host[0],
                                                                              in_shape_host[1], in_shape_host[2], levels,
                                                                              (int4*)ancestors);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int ud3d_rulebook_down2(const int32_t* coords, const int32_t* parents, int n_fine, int n_coarse,
                        const int32_t coarse_dims_host[4], const void* coarse_ws, int32_t* child, int32_t* up,
                        uint32_t* child_mask, uint32_t* up_mask, void* stream) {
  int rc = check_dims(coarse_dims_host, "ud3d_rulebook_down2");
  if (rc) return rc;
  UD3D_CHECK_ARG(coarse_ws && coords && parents && child && up, "ud3d_rulebook_down2: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  GridDims g = make_grid_dims(coarse_dims_host);
  GridView v = grid_view(g, coarse_ws);
  UD3D_CUDA(cudaMemsetAsync(child, 0xFF, (size_t)8 * n_coarse * 4, st));
  UD3D_CUDA(cudaMemsetAsync(up, 0xFF, (size_t)8 * n_fine * 4, st));
  if (child_mask) UD3D_CUDA(cudaMemsetAsync(child_mask, 0, (size_t)cdiv(n_coarse, UD3D_TILE_M) * 4, st));
  if (up_mask) UD3D_CUDA(cudaMemsetAsync(up_mask, 0, (size_t)cdiv(n_fine, UD3D_TILE_M) * 4, st));
  if (n_fine > 0 && n_coarse > 0) {
    down2_fill_kernel<<<grid_blocks(n_fine, 256), 256, 0, st>>>((const int4*)coords, (const int4*)parents, n_fine, n_coarse, g,
                                                                v.words, v.prefix, child, up, child_mask, up_mask);
    UD3D_LAUNCH_CHECK();
  }
  return UD3D_OK;
}

}  // extern "C"
