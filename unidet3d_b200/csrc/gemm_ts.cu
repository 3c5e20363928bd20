// Gather-GEMM, "TS" form: the gathered A operand never touches shared memory.
//
//   out[o,:] = act( sum_k in[table[k][o],:] @ W_k + bias ) + residual[o,:]          (in = operand-form feature map)
//
// Why: with the A tile in shared memory (gemm.cu) a K-step of 128 rows x 32 channels costs the shared-memory port
// 16 KB of LDGSTS writes plus 24 KB of tensor-core reads (the hi/lo split reads A three times) plus the weight reads --
// the level-1 convolutions (N = 32) ran at ~400 cycles per step per SM, i.e. at the shared-memory bandwidth, not at the
// gather latency.  Here
//   * 8 producer warps gather rows with 32-byte loads (LDG.256: 4 lanes per 128-byte row-chunk, 8 full lines per
//     instruction, L1-allocating so that rows shared by several kernel offsets of a tile are L2-fetched once) into
//     registers, software-pipelined kDepth steps deep, and move them with tcgen05.st (16 lanes x 256 bit) straight into
//     a TMEM A stage; the operand form is laid out so that one 32-byte load is exactly one thread's share of that store;
//   * tcgen05.mma reads A from TMEM (no shared-memory traffic) and the pre-packed weight tile from shared memory
//     (cp.async.bulk, one per step); 6 MMAs per step: hi.hi, lo.hi, hi.lo;
//   * the CTA is persistent (one per SM): a table warp prefetches the rulebook slice of the next tile, the accumulator is
//     double-buffered in TMEM, and 4 dedicated epilogue warps (residual rows prefetched before the accumulator is ready)
//     drain tile i while the producers / MMA warp are already on tile i+1.
// Roles (15 warps): 0-3 epilogue (TMEM lane quadrant = warp), 4-11 A producers (quadrant (w-4)&3, 16-lane half (w-4)>>2),
// 12 rulebook slices, 13 weight tiles, 14 MMA issue.  All hand-offs are mbarriers; no block barrier after the prologue.
#include "gemm_common.cuh"

namespace ud3d {

template <int N_TILE> struct TsCfg {
  // ring depth: TMEM A stages (32 columns each, next to the two N_TILE-column accumulators) == weight stages in smem
  static constexpr int kStages = N_TILE <= 64 ? 8 : N_TILE <= 96 ? 6 : N_TILE <= 128 ? 5 : 4;
};
constexpr int kTsThreads = 32 * 15;
constexpr int kTsEpiWarps = 4;
constexpr int kTsProdWarps = 8;
constexpr int kTsWarpTbl = 12, kTsWarpB = 13, kTsWarpMma = 14;

// 16 TMEM lanes x 32 columns (one A stage of a half quadrant); register layout of .16x256b.x4: registers 4g, 4g+1 =
// columns 8g + 2q, 8g + 2q + 1 of lane base + t/4; registers 4g+2, 4g+3 = the same columns of lane base + 8 + t/4
__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint32_t (&a)[8], const uint32_t (&b)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(a[0]), "r"(a[1]), "r"(b[0]), "r"(b[1]), "r"(a[2]), "r"(a[3]), "r"(b[2]), "r"(b[3]), "r"(a[4]), "r"(a[5]), "r"(b[4]),
      "r"(b[5]), "r"(a[6]), "r"(a[7]), "r"(b[6]), "r"(b[7])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_bf16_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                   uint32_t accumulate) {
  if (elect_one_sync()) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// ---------------------------------------------------------------- weight packing for the TS kernel
// Same tile image as pack_weight_kernel (gemm.cu) -- [n_tile][k][chunk][N_TILE rows][8 x 16 B], 16-byte pieces
// XOR-swizzled with the row -- but the K order inside a 32-channel chunk follows the TMEM A operand: the row's 64
// bf16 positions are four K=16 slices, slice s in {hi(P0), hi(P1), lo(P0), lo(P1)}, and position 4q + j of a slice is
// channel 8q + j (P0) or 8q + 4 + j (P1)  (q = the producer lane's 32-byte quarter of the row-chunk).
__global__ void pack_weight_ts_kernel(const float* __restrict__ w, int K, int c_in, int c_out, int n_tile_sz, int n_tiles,
                                      int n_chunks, uint4* __restrict__ out) {
  long long total = (long long)n_tiles * K * n_chunks * n_tile_sz * 8;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int jj = (int)(t & 7);
    long long r = t >> 3;
    int n = (int)(r % n_tile_sz);
    r /= n_tile_sz;
    int c = (int)(r % n_chunks);
    r /= n_chunks;
    int k = (int)(r % K);
    int nt = (int)(r / K);
    int ng = nt * n_tile_sz + n;
    const int slice = jj >> 1;                  // 0,1 hi; 2,3 lo; odd = P1
    uint32_t v[4];
#pragma unroll
    for (int e2 = 0; e2 < 4; ++e2) {
      float x[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int e = 2 * e2 + hh;              // position 8 (jj & 1) + e of the slice
        const int q = 2 * (jj & 1) + (e >> 2), j = e & 3;
        const int ch = c * kChunk + 8 * q + j + ((slice & 1) ? 4 : 0);
        x[hh] = (ng < c_out && ch < c_in) ? w[((size_t)ng * K + k) * c_in + ch] : 0.f;
      }
      uint32_t hi, lo;
      split_bf16x2(x[0], x[1], hi, lo);
      v[e2] = (slice < 2) ? hi : lo;
    }
    size_t tile = ((size_t)nt * K + k) * n_chunks + c;
    out[tile * n_tile_sz * 8 + (size_t)n * 8 + (jj ^ (n & 7))] = make_uint4(v[0], v[1], v[2], v[3]);
  }
}

int launch_pack_weight_ts(const float* w, int K, int c_in, int c_out, int nts, void* packed, cudaStream_t st) {
  int n_tiles = cdiv(c_out, nts), n_chunks = cdiv(c_in, kChunk);
  long long total = (long long)n_tiles * K * n_chunks * nts * 8;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  pack_weight_ts_kernel<<<blocks, 256, 0, st>>>(w, K, c_in, c_out, nts, n_tiles, n_chunks, (uint4*)packed);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

// ---------------------------------------------------------------- the kernel
// A rulebook slot in shared memory (two slots, filled two tiles ahead by the table warp):
//   int32 tbl[K + 1][128]   the tile's slice of the gather table; row K is all -1 (the "null" offset)
//   int32 nsteps, pad[3]
//   uint16 steps[kTsMaxSteps]   (k << 8) | chunk of every K-step of the tile, k == K for a null step
// nsteps = max(active offsets x chunks, kTsMinSteps): a tile of at least kTsMinSteps steps keeps a producer's load cursor
// at most one tile ahead of its store cursor, which is what makes two slots enough (no deadlock).
constexpr int kTsMaxSteps = 256;
constexpr int kTsMinSteps = 4;
constexpr int kTsInFlight = 2;        // K-steps in flight per producer warp (32 registers each); the two warps of a TMEM lane
                                      // quadrant take alternate steps -> 4 steps (64 KB) in flight per SM

__device__ uint4 g_ts_zero_row[8];    // 128 bytes of zeros: the source row of a missing neighbour (always an L1 hit)

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((uint16_t)v) : "memory");
}
__device__ __forceinline__ void ldg256(const void* p, uint32_t (&v)[8]) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
}

// one K-step of the MMA warp at ring stage S (compile-time): wait, 6 MMAs, commit.  All tensor-memory / descriptor
// operands are compile-time offsets from two loop-invariant values, so the elected lane issues from uniform registers.
#define UD3D_TS_MMA_STEP(S)                                                                                   \
  if (t < nsteps) {                                                                                           \
    mbar_wait(&full[S], use & 1u);                                                                            \
    tc_fence_after_sync();                                                                                    \
    {                                                                                                         \
      const uint32_t at = tmem_base + COL_A + (uint32_t)((S) * 32);                                           \
      const uint64_t bd = bdesc0 + (uint64_t)((S) * (B_BYTES >> 4));                                          \
      umma_bf16_ts_elect(d_tmem, at + 0, bd + 0, IDESC, t > 0);                                               \
      umma_bf16_ts_elect(d_tmem, at + 8, bd + 2, IDESC, 1);                                                   \
      umma_bf16_ts_elect(d_tmem, at + 16, bd + 0, IDESC, 1);                                                  \
      umma_bf16_ts_elect(d_tmem, at + 24, bd + 2, IDESC, 1);                                                  \
      umma_bf16_ts_elect(d_tmem, at + 0, bd + 4, IDESC, 1);                                                   \
      umma_bf16_ts_elect(d_tmem, at + 8, bd + 6, IDESC, 1);                                                   \
      umma_commit_elect(&empty[S]);                                                                           \
    }                                                                                                         \
    ++t;                                                                                                      \
    if ((S) == STAGES - 1) ++use;                                                                             \
  }

template <int N_TILE>
__global__ void __launch_bounds__(kTsThreads, 1) gather_gemm_ts_kernel(const GemmParams p) {
  constexpr int STAGES = TsCfg<N_TILE>::kStages;
  constexpr int B_BYTES = N_TILE * 128;
  constexpr uint32_t IDESC = umma_idesc_bf16_m128(N_TILE);
  constexpr uint32_t COL_A = 2 * N_TILE;            // TMEM: [acc 0 | acc 1 | A stage 0 | A stage 1 | ...]
  static_assert(COL_A + STAGES * 32 <= 512, "TMEM budget");
  static_assert(STAGES <= 8, "the MMA warp's stage dispatch covers 8 stages");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const ud3d_gemm_args& a = p.a;
  const bool has_table = a.table != nullptr;
  const int tbl_bytes = has_table ? (a.K + 1) * kTileM * 4 : 0;
  const int slot_bytes = tbl_bytes + 16 + kTsMaxSteps * 2;
  uint8_t* sB = smem;                                        // [STAGES][N_TILE x 128 B]
  uint8_t* sEpi = sB + STAGES * B_BYTES;                     // [4 warps][4 KB] row-segment transposition
  uint8_t* sTbl = sEpi + kTsEpiWarps * 4096;                 // [2 slots]
  uint64_t* bars = (uint64_t*)(sTbl + 2 * slot_bytes);
  uint64_t* full = bars;                   // [STAGES] 4 producer warps + 1 weight copy (+ tx bytes)
  uint64_t* empty = full + 8;              // [STAGES] tcgen05.commit
  uint64_t* tbl_full = empty + 8;          // [2] table warp
  uint64_t* tbl_empty = tbl_full + 2;      // [2] 8 producer warps + weight warp + MMA warp
  uint64_t* acc_full = tbl_empty + 2;      // [2] tcgen05.commit
  uint64_t* acc_empty = acc_full + 2;      // [2] 4 epilogue warps
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_row_tiles = (a.n_out + kTileM - 1) / kTileM;
  const int n_ntiles = (a.c_out + N_TILE - 1) / N_TILE;
  const int n_items = n_row_tiles * n_ntiles;
  const uint32_t sTbl_u32 = smem_u32(sTbl);

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], kTsProdWarps / 2 + 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tbl_full[s], 1);
      mbar_init(&tbl_empty[s], kTsProdWarps + 2);
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], kTsEpiWarps);
    }
    fence_mbar_init();
  }
  if (has_table && tid < 2 * kTileM)      // the null offset: row K of both slots
    sts_u32(sTbl_u32 + (uint32_t)((tid >> 7) * slot_bytes + (a.K * kTileM + (tid & 127)) * 4), 0xffffffffu);
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kTsEpiWarps) {
    // ================================================================= epilogue: TMEM -> registers -> global
    const int quad = warp;
    const uint32_t stage = smem_u32(sEpi) + (uint32_t)warp * 4096u;
    int it = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const int rt = w / n_ntiles, nt = w - rt * n_ntiles;
      const int m0 = rt * kTileM, n0 = nt * N_TILE;
      const int row = quad * 32 + lane;
      const bool row_ok = m0 + row < a.n_out;
      const int grow = (row_ok && a.row_perm) ? __ldg(a.row_perm + m0 + row) : m0 + row;
      const int buf = it & 1;
      mbar_wait(&acc_full[buf], (uint32_t)(it >> 1) & 1u);
      tc_fence_after_sync();
#pragma unroll 1
      for (int c0 = 0; c0 < N_TILE; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * N_TILE + c0), r);
        tmem_ld_wait();
        if (c0 + 32 >= N_TILE) {
          // the accumulator has been read: hand it back to the MMA warp before the stores
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        epilogue_store_chunk(p, r, grow, n0 + c0, row_ok, stage, lane);
      }
    }
  } else if (warp < kTsEpiWarps + kTsProdWarps) {
    // ================================================================= A producers: global -> registers -> TMEM
    const int pw = warp - kTsEpiWarps;
    const int quad = pw & 3, par = pw >> 2;             // this warp fills the quadrant's 32 lanes for steps g == par (mod 2)
    const int rsub = lane >> 2, q = lane & 3;
    const int r0 = quad * 32 + rsub;                    // this thread's four tile rows: r0, r0 + 8, r0 + 16, r0 + 24
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + COL_A;
    const size_t row_bytes = (size_t)a.ld_in * 4u;
    const uint8_t* in_q = (const uint8_t*)a.in + q * 32;
    const uint8_t* zero_q = (const uint8_t*)g_ts_zero_row + q * 32;

    // two cursors over the CTA's flattened K-step sequence: `is` issues the loads kTsInFlight of this warp's steps
    // ahead of `cs`, which stores them to TMEM
    struct Cur {
      int it, w, t, nsteps, m0;
      uint32_t tbl, steps;         // shared-memory addresses of the tile's rulebook slice / step list
    };
    auto open_tile = [&](Cur& cu, bool wait) {
      if (cu.w >= n_items) return;
      if (wait) mbar_wait(&tbl_full[cu.it & 1], (uint32_t)(cu.it >> 1) & 1u);
      cu.tbl = sTbl_u32 + (uint32_t)((cu.it & 1) * slot_bytes);
      cu.steps = cu.tbl + (uint32_t)tbl_bytes + 16u;
      cu.nsteps = (int)lds_u32(cu.tbl + (uint32_t)tbl_bytes);
      cu.m0 = (cu.w / n_ntiles) * kTileM;
    };
    auto advance = [&](Cur& cu, bool is_issue) {
      cu.t += 2;
      if (cu.t >= cu.nsteps) {
        cu.t -= cu.nsteps;            // (< 2 <= kTsMinSteps <= the next tile's step count)
        if (!is_issue) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&tbl_empty[cu.it & 1]);       // this warp is done with the tile's slot
        }
        ++cu.it;
        cu.w += gridDim.x;
        open_tile(cu, is_issue);
      }
    };
    Cur is, cs;
    is.it = cs.it = 0;
    is.w = cs.w = blockIdx.x;
    is.t = cs.t = par;
    is.nsteps = cs.nsteps = 0; is.m0 = cs.m0 = 0; is.tbl = cs.tbl = is.steps = cs.steps = 0;
    open_tile(is, true);
    open_tile(cs, false);

    uint32_t buf[kTsInFlight][4][8];
    auto issue = [&](uint32_t (&v)[4][8]) {
      if (is.w >= n_items) return;
      const uint32_t st = lds_u16(is.steps + 2u * (uint32_t)is.t);
      const int k = (int)(st >> 8), c = (int)(st & 255u);
      int idx[4];
      if (has_table) {
        const uint32_t trow = is.tbl + (uint32_t)((k * kTileM + r0) * 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) idx[j] = (int)lds_u32(trow + 32u * j);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) idx[j] = (k == 0 && is.m0 + r0 + 8 * j < a.n_out) ? is.m0 + r0 + 8 * j : -1;
      }
      const uint8_t* src = in_q + c * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) ldg256(idx[j] >= 0 ? src + (size_t)idx[j] * row_bytes : zero_q, v[j]);
      advance(is, true);
    };
    int gs = par;              // ring stage / use count of this warp's next step (global step g == par mod 2)
    uint32_t guse = 0;
    if (gs >= STAGES) { gs -= STAGES; ++guse; }
    auto consume = [&](const uint32_t (&v)[4][8]) {
      if (cs.w >= n_items) return;
      if (guse) mbar_wait(&empty[gs], (guse & 1u) ^ 1u);
      tc_fence_after_sync();
      const uint32_t ta = t_lane + (uint32_t)(gs * 32);
      tmem_st_16x256b_x4(ta, v[0], v[1]);                        // lanes +0..15  (rows r0, r0 + 8)
      tmem_st_16x256b_x4(ta + (16u << 16), v[2], v[3]);          // lanes +16..31 (rows r0 + 16, r0 + 24)
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[gs]);
      gs += 2;
      if (gs >= STAGES) { gs -= STAGES; ++guse; }
      advance(cs, false);
    };
#pragma unroll
    for (int j = 0; j < kTsInFlight; ++j) issue(buf[j]);
    while (cs.w < n_items) {
#pragma unroll
      for (int j = 0; j < kTsInFlight; ++j) {
        consume(buf[j]);
        issue(buf[j]);
      }
    }
  } else if (warp == kTsWarpTbl) {
    // ================================================================= rulebook slices + step lists, two tiles ahead
    int it = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const int rt = w / n_ntiles;
      const int m0 = rt * kTileM;
      if (it >= 2) mbar_wait(&tbl_empty[it & 1], (uint32_t)((it >> 1) - 1) & 1u);
      const uint32_t slot = sTbl_u32 + (uint32_t)((it & 1) * slot_bytes);
      uint32_t mask = 1u;
      if (has_table) {
        const bool full_tile = m0 + kTileM <= a.n_out && (a.n_out & 3) == 0 && ((uintptr_t)a.table & 15) == 0;
        if (full_tile) {
          for (int i = lane; i < a.K * (kTileM / 4); i += 32) {
            const int k = i >> 5, r4 = (i & 31) * 4;
            cp_async_16(slot + (uint32_t)((k * kTileM + r4) * 4), a.table + (size_t)k * a.n_out + m0 + r4);
          }
          cp_async_commit();
          cp_async_wait<0>();
        } else {
          for (int i = lane; i < a.K * kTileM; i += 32) {
            const int k = i >> 7, r = i & 127;
            sts_u32(slot + (uint32_t)(i * 4), (m0 + r < a.n_out) ? (uint32_t)__ldg(a.table + (size_t)k * a.n_out + m0 + r) : 0xffffffffu);
          }
        }
        __syncwarp();
        if (a.tile_mask) {
          mask = __ldg(a.tile_mask + rt);
        } else {
          mask = 0u;
          for (int k = 0; k < a.K; ++k) {
            int4 v;
            asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(slot + (uint32_t)((k * kTileM + lane * 4) * 4)));
            const bool any = (v.x >= 0) | (v.y >= 0) | (v.z >= 0) | (v.w >= 0);
            if (__any_sync(0xffffffffu, any)) mask |= 1u << k;
          }
        }
      }
      // step list: (active offset, chunk) pairs, padded with null steps up to kTsMinSteps
      const int nact = __popc(mask);
      const int nreal = nact * p.n_chunks;
      const int nsteps = nreal < kTsMinSteps ? kTsMinSteps : nreal;
      const uint32_t steps = slot + (uint32_t)tbl_bytes + 16u;
      for (int i = lane; i < nsteps; i += 32) {
        uint32_t e = (uint32_t)(has_table ? a.K : 1) << 8;         // null step
        if (i < nreal) {
          const int ks = i / p.n_chunks, c = i - ks * p.n_chunks;
          // ks-th set bit of the mask
          uint32_t m = mask;
          for (int j = 0; j < ks; ++j) m &= m - 1u;
          e = ((uint32_t)(__ffs(m) - 1) << 8) | (uint32_t)c;
        }
        sts_u16(steps + 2u * (uint32_t)i, e);
      }
      if (lane == 0) sts_u32(slot + (uint32_t)tbl_bytes, (uint32_t)nsteps);
      __syncwarp();
      if (lane == 0) mbar_arrive(&tbl_full[it & 1]);
    }
  } else if (warp == kTsWarpB) {
    // ================================================================= weight tiles: one bulk copy per step
    int it = 0, s = 0;
    uint32_t use = 0;
    const int knull = has_table ? a.K : 1;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const int rt = w / n_ntiles, nt = w - rt * n_ntiles;
      mbar_wait(&tbl_full[it & 1], (uint32_t)(it >> 1) & 1u);
      const uint32_t slot = sTbl_u32 + (uint32_t)((it & 1) * slot_bytes);
      const int nsteps = (int)lds_u32(slot + (uint32_t)tbl_bytes);
      const uint32_t steps = slot + (uint32_t)tbl_bytes + 16u;
      const uint8_t* wp = (const uint8_t*)a.w_packed_ts + (size_t)nt * a.K * p.n_chunks * B_BYTES;
      for (int t = 0; t < nsteps; ++t) {
        const uint32_t st = lds_u16(steps + 2u * (uint32_t)t);
        int k = (int)(st >> 8);
        const int c = (int)(st & 255u);
        if (k == knull) k = 0;                 // a null step multiplies all-zero A rows: any weight tile will do
        if (use) mbar_wait(&empty[s], (use & 1u) ^ 1u);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full[s], B_BYTES);
          bulk_copy_g2s(sB + s * B_BYTES, wp + ((size_t)k * p.n_chunks + c) * B_BYTES, B_BYTES, &full[s]);
        }
        if (++s == STAGES) { s = 0; ++use; }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&tbl_empty[it & 1]);
    }
  } else {
    // ================================================================= MMA issue (whole warp converged, elect.sync)
    const uint64_t bdesc0 = umma_desc_sw128(smem_u32(sB));
    int it = 0;
    uint32_t gbase = 0;             // K-steps issued so far by this CTA: ring stage = gbase % STAGES, use = gbase / STAGES
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      mbar_wait(&tbl_full[it & 1], (uint32_t)(it >> 1) & 1u);
      const int nsteps = (int)lds_u32(sTbl_u32 + (uint32_t)((it & 1) * slot_bytes + tbl_bytes));
      __syncwarp();
      if (lane == 0) mbar_arrive(&tbl_empty[it & 1]);         // (only the step count is needed)
      const int buf = it & 1;
      if (it >= 2) mbar_wait(&acc_empty[buf], (uint32_t)((it >> 1) - 1) & 1u);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * N_TILE);
      int t = 0;
      uint32_t s = gbase % STAGES, use = gbase / STAGES;
      while (t < nsteps) {
        // enter the stage sequence at the ring's current stage, then run the stages in order: every tensor-memory
        // address / descriptor offset is a compile-time constant
        switch (s) {
          case 0: UD3D_TS_MMA_STEP(0)
          case 1: UD3D_TS_MMA_STEP(1)
          case 2: UD3D_TS_MMA_STEP(2)
          case 3: UD3D_TS_MMA_STEP(3)
          case 4: if (STAGES > 4) { UD3D_TS_MMA_STEP(4) }
          case 5: if (STAGES > 5) { UD3D_TS_MMA_STEP(5) }
          case 6: if (STAGES > 6) { UD3D_TS_MMA_STEP(6) }
          case 7: if (STAGES > 7) { UD3D_TS_MMA_STEP(7) }
          default: break;
        }
        s = 0;
      }
      gbase += (uint32_t)nsteps;
      umma_commit_elect(&acc_full[buf]);
    }
    __syncwarp();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int N_TILE>
static size_t ts_smem_bytes(int K, bool has_table) {
  const size_t slot = (has_table ? (size_t)(K + 1) * kTileM * 4 : 0) + 16 + kTsMaxSteps * 2;
  return 1024 + (size_t)TsCfg<N_TILE>::kStages * N_TILE * 128 + kTsEpiWarps * 4096 + 2 * slot + (16 + 8) * 8 + 16;
}

template <int N_TILE>
static int launch_ts(const GemmParams& p, int num_sms, cudaStream_t st) {
  const size_t smem = ts_smem_bytes<N_TILE>(p.a.K, p.a.table != nullptr);
  // per-device configuration (cudaFuncSetAttribute applies to the current device)
  static size_t configured[64] = {0};
  int dev = 0;
  UD3D_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { set_error("ud3d_gemm_fwd: device index %d out of range", dev); return UD3D_EINVAL; }
  if (smem > configured[dev]) {
    UD3D_CUDA(cudaFuncSetAttribute(gather_gemm_ts_kernel<N_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // leave the rest of the 256 KB to L1: gathered rows are re-used across the kernel offsets of a tile
    int pct = (int)((smem + 1024) * 100 / (228 * 1024)) + 1;
    if (pct > 100) pct = 100;
    UD3D_CUDA(cudaFuncSetAttribute(gather_gemm_ts_kernel<N_TILE>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    configured[dev] = smem;
  }
  const int n_items = cdiv(p.a.n_out, kTileM) * cdiv(p.a.c_out, N_TILE);
  const int grid = n_items < num_sms ? n_items : num_sms;
  gather_gemm_ts_kernel<N_TILE><<<grid, kTsThreads, smem, st>>>(p);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int launch_gemm_ts(const GemmParams& p, int nts, int num_sms, cudaStream_t st) {
  switch (nts) {
    case 32: return launch_ts<32>(p, num_sms, st);
    case 64: return launch_ts<64>(p, num_sms, st);
    case 96: return launch_ts<96>(p, num_sms, st);
    case 128: return launch_ts<128>(p, num_sms, st);
    case 160: return launch_ts<160>(p, num_sms, st);
    default: set_error("ud3d_gemm_fwd: no TS kernel for N_TILE %d", nts); return UD3D_EINVAL;
  }
}

}  // namespace ud3d
