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
// Warp roles (28 warps = 7 warpgroups): 0-3 epilogue; 4-23 A producers, five per TMEM lane quadrant (quadrant = warp & 3,
// turn = (warp - 4) >> 2: the warp fills K-steps g == turn (mod 5)); 24, 25 weight tiles; 26 MMA issue; 27 idle.
// Memory-level parallelism comes from the NUMBER of producer warps, each with a single K-step (32 rows x 32 B per lane)
// in flight: one warp cannot keep several independent load batches in flight without waiting for the newest (loads
// that share a hardware scoreboard complete "in order" as far as a dependent instruction is concerned).
constexpr int kTsThreads = 32 * 28;
constexpr int kTsEpiWarps = 4;
constexpr int kTsTurns = 5;           // producer warps per quadrant
constexpr int kTsProdWarps = 4 * kTsTurns;
constexpr int kTsWarpB = kTsEpiWarps + kTsProdWarps;    // 24, 25: weight tiles (alternate K-steps)
constexpr int kTsWarpMma = kTsWarpB + 2;
// registers per thread after the role split (setmaxnreg, per warpgroup): the launch allocates 65536 / 896 = 72
constexpr int kTsRegsEpi = 120, kTsRegsProd = 64, kTsRegsCtl = 64;   // 120: the pool is what the others release (6144 registers)
#ifndef UD3D_TS_NO_SETMAXNREG
#define UD3D_TS_SETMAXNREG(dir, n) asm volatile("setmaxnreg." dir ".sync.aligned.u32 %0;" ::"n"(n))
#else
#define UD3D_TS_SETMAXNREG(dir, n) do { } while (0)
#endif

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
// Every role derives the K-step sequence of a tile from ONE word, the tile's mask of active kernel offsets
// (ud3d_gemm_args.tile_mask; 1 for a dense GEMM): steps = (active offset in ascending order) x (32-channel chunk).
// Nothing about the rulebook is staged in shared memory: a producer thread reads the four row indices of its next
// step straight from the gather table (one 32-byte sector per 8 rows), two of its steps ahead of the row loads, which
// are two steps ahead of the TMEM store -- the dependent index -> row chain is software-pipelined in registers.
constexpr int kTsInFlight = 2;        // K-steps in flight per producer warp (32 registers each); the two warps of a TMEM lane
                                      // quadrant take alternate steps -> 4 steps (64 KB) in flight per SM

__device__ uint4 g_ts_zero_row[8];    // 128 bytes of zeros: the source row of a missing neighbour (always an L1 hit)

__device__ __forceinline__ void ldg256(const void* p, uint32_t (&v)[8]) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
}
// idx = ok ? *p : -1, as a predicated load (no select on the loaded value: the result is not needed before its use two
// iterations later, so the warp must not wait for it here)
__device__ __forceinline__ int ldg_s32_or_m1(const int32_t* p, bool ok) {
  int v;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %2, 0;\n\t"
      "mov.b32 %0, -1;\n\t"
      "@p ld.global.nc.s32 %0, [%1];\n\t}"
      : "=r"(v)
      : "l"(p), "r"((int)ok));
  return v;
}
__device__ __forceinline__ uint32_t ldg_u32_raw(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// The 6 MMAs of one K-step + the commit that frees the stage, issued by ONE elected lane from one asm block (one
// elect / predicate region instead of seven).  A slices (8 TMEM columns each): hi(P0), hi(P1), lo(P0), lo(P1); B slices
// (+32 B = +2 in the descriptor's address field): hi(P0), hi(P1), lo(P0), lo(P1); products hi.hi, lo.hi, hi.lo.
__device__ __forceinline__ void umma_ts_step(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate, uint32_t empty_bar) {
  if (elect_one_sync()) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b32 a1, a2, a3;\n\t.reg .b64 b2, b4, b6;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.b32 q, 0, 0;\n\t"
        "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\t"
        "add.u64 b2, %2, 2;\n\tadd.u64 b4, %2, 4;\n\tadd.u64 b6, %2, 6;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b2, %3, q;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], %2, %3, q;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], b2, %3, q;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], b4, %3, q;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b6, %3, q;\n\t"
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(empty_bar)
        : "memory");
  }
}

// debug trace (ud3d_debug_set_trace(buf >= 4096 int64, block)): clock64 of CTA `block`; per K-step g < 256: [8g + 0]
// MMA warp saw the stage full, +1 MMAs issued, +2 producer (quadrant 0) starts waiting for the stage, +3 has it, +4 stage
// published, +5 next loads issued, +6 weight copy issued; per tile it < 64: [2048 + 4 it + 0] epilogue sees the
// accumulator, +1 epilogue done, +2 MMA warp starts the tile
#define UD3D_TS_TR(idx)                                                              \
  do {                                                                               \
    if (traced && lane == 0 && (idx) < 4096) p.trace[(idx)] = clock64();             \
  } while (0)

// one K-step of the MMA warp at ring stage S (compile-time): wait, 6 MMAs, commit.  All tensor-memory / descriptor
// operands are compile-time offsets from loop-invariant values.
#define UD3D_TS_MMA_STEP(S)                                                                                   \
  if (t < nsteps) {                                                                                           \
    mbar_wait(&full[S], use & 1u);                                                                            \
    tc_fence_after_sync();                                                                                    \
    if (traced && gbase + t < 256) UD3D_TS_TR(8 * (gbase + t) + 0);                                           \
    if (!UD3D_DBG(p, 1)) {                                                                                       \
      umma_ts_step(d_tmem, tmem_base + COL_A + (uint32_t)((S) * 32), bdesc0 + (uint64_t)((S) * (B_BYTES >> 4)), IDESC, \
                   t > 0, empty_u32 + 8u * (S));                                                              \
    } else if (lane == 0) {                                                                                   \
      mbar_arrive(&empty[S]);                                                                                 \
    }                                                                                                         \
    if (traced && gbase + t < 256) UD3D_TS_TR(8 * (gbase + t) + 1);                                           \
    ++t;                                                                                                      \
    if ((S) == STAGES - 1) ++use;                                                                             \
  }

// A tile's mask of active kernel offsets as every role sees it: the tile_mask word, or 1 when it is 0 (a tile without
// any input still runs one all-zero K-step per chunk so that its accumulator is defined) or when there is no table.
template <int N_TILE>
__global__ void __launch_bounds__(kTsThreads, 1) gather_gemm_ts_kernel(const GemmParams p) {
  constexpr int STAGES = TsCfg<N_TILE>::kStages;
  constexpr int B_BYTES = N_TILE * 128;
  constexpr uint32_t IDESC = umma_idesc_bf16_m128(N_TILE);
  constexpr uint32_t COL_A = 2 * N_TILE;            // TMEM: [acc 0 | acc 1 | A stage 0 | A stage 1 | ...]
  static_assert(COL_A + STAGES * 32 <= 512, "TMEM budget");
  static_assert(STAGES <= 8, "the MMA warp's stage dispatch covers 8 stages");
  // weight tiles: one bulk copy (TMA engine) per step for the wide tiles -- its issue costs ~300 cycles, hidden behind >= 192
  // cycles of MMAs per step and two alternating warps -- and 16-byte cp.async for the 4 KB tiles of N = 32, where a step is
  // too short for that
  constexpr bool kBulkB = N_TILE >= 64;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const ud3d_gemm_args& a = p.a;
  const bool has_table = a.table != nullptr;
  uint8_t* sB = smem;                                        // [STAGES][N_TILE x 128 B]
  uint8_t* sEpi = sB + STAGES * B_BYTES;                     // [4 warps][4 KB] row-segment transposition
  uint8_t* sIdx = sEpi + kTsEpiWarps * 4096;                 // [producer warp][2][32 lanes x 16 B] row indices of the next step
  uint64_t* bars = (uint64_t*)(sIdx + kTsProdWarps * 1024);
  uint64_t* full = bars;                   // [STAGES] 4 producer warps (one per quadrant) + 32 lanes of a weight warp (cp.async arrivals)
  uint64_t* empty = full + 8;              // [STAGES] tcgen05.commit
  uint64_t* acc_full = empty + 8;          // [2] tcgen05.commit
  uint64_t* acc_empty = acc_full + 2;      // [2] 4 epilogue warps
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_row_tiles = (a.n_out + kTileM - 1) / kTileM;
  const int n_ntiles = (a.c_out + N_TILE - 1) / N_TILE;
  const int n_items = n_row_tiles * n_ntiles;
  const bool traced = UD3D_TRACE_BUF(p) && (int)blockIdx.x == p.trace_block;
  // trace_block == -3: globaltimer at start / end + SM id of every CTA: p.trace[4 * blockIdx.x + {0, 1, 2}]
  if (UD3D_TRACE_BUF(p) && p.trace_block == -3 && tid == 0) {
    unsigned long long gt;
    unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    p.trace[4 * blockIdx.x + 0] = (long long)gt;
    p.trace[4 * blockIdx.x + 2] = smid;
  }

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 4 + (kBulkB ? 1 : 32));
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], kTsEpiWarps);
    }
    fence_mbar_init();
  }
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
    UD3D_TS_SETMAXNREG("inc", kTsRegsEpi);
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
      if (warp == 0 && it < 64) UD3D_TS_TR(2048 + 4 * it + 0);
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
        epilogue_store_chunk(p, r, grow, n0 + c0, row_ok && !UD3D_DBG(p, 16), stage, lane);
      }
      if (warp == 0 && it < 64) UD3D_TS_TR(2048 + 4 * it + 1);
    }
  } else if (warp < kTsEpiWarps + kTsProdWarps) {
    // ================================================================= A producers: global -> registers -> TMEM
    UD3D_TS_SETMAXNREG("dec", kTsRegsProd);
    const int pw = warp - kTsEpiWarps;
    const int quad = pw & 3, turn = pw >> 2;            // this warp fills the quadrant's 32 lanes for steps g == turn (mod 5)
    const int rsub = lane >> 2, q = lane & 3;
    const int r0 = quad * 32 + rsub;                    // this thread's four tile rows: r0, r0 + 8, r0 + 16, r0 + 24
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + COL_A;
    const uint32_t row_bytes = (uint32_t)a.ld_in * 4u;
    const uint8_t* in_q = (const uint8_t*)a.in + q * 32;
    const uint8_t* zero_q = (const uint8_t*)g_ts_zero_row + q * 32;
    const uint32_t full_u32 = smem_u32(full);
    constexpr uint32_t kInvalid = 0xffffffffu;

    // ---- the step generator.  Cursor = (tile, remaining active offsets, chunk); gen() emits this warp's next step --
    //      meta = chunk byte offset (or kInvalid past the end) and the four row indices, loaded from the gather table
    //      (not waited for here) -- and moves kTsTurns steps ahead.
    int g_w = blockIdx.x;                 // work item
    uint32_t g_m = 0u;                    // active offsets of the tile not yet passed (lowest set bit = current offset)
    int g_c = 0, g_m0 = 0;
    uint32_t g_mnext = 0u;                // raw mask of the CTA's next tile (loaded one tile ahead)
    auto g_open = [&]() {                 // g_w is a valid item whose (raw) mask is in g_mnext
      g_m = (has_table && g_mnext) ? g_mnext : 1u;
      g_c = 0;
      g_m0 = (g_w / n_ntiles) * kTileM;
      const int wn = g_w + (int)gridDim.x;
      if (has_table && wn < n_items) g_mnext = ldg_u32_raw(a.tile_mask + wn / n_ntiles);    // used one tile later
    };
    auto g_step1 = [&]() {                // advance the cursor by one K-step
      if (++g_c == p.n_chunks) {
        g_c = 0;
        g_m &= g_m - 1u;
        if (g_m == 0u) {
          g_w += (int)gridDim.x;
          if (g_w < n_items) g_open();
        }
      }
    };
    // The four row indices of a step travel global -> shared memory by cp.async (4 bytes each, this lane's own 16-byte
    // slot) and are read back with one LDS.128 an iteration later: an asynchronous copy is tracked by the cp.async group
    // counter, not by a register scoreboard, so having it in flight does not stall the instructions that touch the
    // previously loaded indices or the row registers.  Rows beyond n_out (last tile) read a clamped, valid table entry:
    // whatever they gather lands in accumulator rows the epilogue never stores.
    const uint32_t idx_slot = smem_u32(sIdx) + (uint32_t)(pw * 1024 + lane * 16);
    auto gen = [&](uint32_t& meta, uint32_t slot) {
      if (g_w >= n_items) { meta = kInvalid; return; }
      meta = (uint32_t)g_c * 128u;
      if (has_table) {
        const int k = __ffs(g_m) - 1;
        const int32_t* trow = a.table + (size_t)k * a.n_out;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int r = g_m0 + r0 + 8 * j;
          r = r < a.n_out ? r : a.n_out - 1;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(idx_slot + slot * 512u + 4u * j), "l"(trow + r) : "memory");
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = g_m0 + r0 + 8 * j;
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(idx_slot + slot * 512u + 4u * j), "r"(r < a.n_out ? r : -1) : "memory");
        }
      }
      cp_async_commit();
#pragma unroll 1
      for (int j = 0; j < kTsTurns && g_w < n_items; ++j) g_step1();
    };
    if (g_w < n_items) {
      if (has_table) g_mnext = ldg_u32_raw(a.tile_mask + g_w / n_ntiles);
      g_open();
#pragma unroll 1
      for (int j = 0; j < turn && g_w < n_items; ++j) g_step1();
    }
    uint32_t g = (uint32_t)turn;      // global K-step index of this warp's next step
    uint32_t meta_n, it = 0;
    gen(meta_n, 0u);
    while (meta_n != kInvalid) {
      if (traced && quad == 0 && g < 256) UD3D_TS_TR(8 * g + 6);
      const uint32_t meta = meta_n;
      cp_async_wait<0>();
      int4 iv;
      asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(iv.x), "=r"(iv.y), "=r"(iv.z), "=r"(iv.w) : "r"(idx_slot + (it & 1u) * 512u));
      const int idx[4] = {iv.x, iv.y, iv.z, iv.w};
      uint32_t rows[4][8];
      const uint8_t* src = in_q + meta;
#pragma unroll
      for (int r = 0; r < 4; ++r)
        ldg256((idx[r] >= 0 && !UD3D_DBG(p, 32)) ? src + (size_t)((uint32_t)idx[r]) * row_bytes : zero_q, rows[r]);
      if (traced && quad == 0 && g < 256) UD3D_TS_TR(8 * g + 7);
      ++it;
      gen(meta_n, it & 1u);        // indices of this warp's next step: in flight while the rows arrive and are stored
      const uint32_t gs = g % STAGES, gph = ((g / STAGES) & 1u) ^ 1u;
      if (traced && quad == 0 && g < 256) UD3D_TS_TR(8 * g + 2);
      mbar_wait(&empty[gs], gph);
      tc_fence_after_sync();
      if (traced && quad == 0 && g < 256) UD3D_TS_TR(8 * g + 3);
      const uint32_t ta = t_lane + gs * 32u;
      tmem_st_16x256b_x4(ta, rows[0], rows[1]);                    // lanes +0..15  (rows r0, r0 + 8)
      tmem_st_16x256b_x4(ta + (16u << 16), rows[2], rows[3]);      // lanes +16..31 (rows r0 + 16, r0 + 24)
      if (traced && quad == 0 && g < 256) UD3D_TS_TR(8 * g + 5);
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_u32 + 8u * gs) : "memory");
      if (traced && quad == 0 && g < 256) UD3D_TS_TR(8 * g + 4);
      g += kTsTurns;
    }
  } else {
   // (one setmaxnreg for the whole control warpgroup, before its warps part ways)
   UD3D_TS_SETMAXNREG("dec", kTsRegsCtl);
   if (warp < kTsWarpMma) {
    // ================================================================= weight tiles: 16-byte async copies (LDGSTS), the two
    //                                                                   warps take alternate K-steps
    const int bw = warp - kTsWarpB;
    constexpr int kPer = N_TILE * 8 / 32;                // 16-byte pieces per lane and step
    uint32_t s = (uint32_t)bw, ph = 1u;
    if (s >= STAGES) { s -= STAGES; ph ^= 1u; }
    uint32_t g = 0;                                      // global step index
    int gtr = bw;
    const uint32_t sB_u32 = smem_u32(sB) + (uint32_t)lane * 16u;
    uint32_t mnext = (has_table && (int)blockIdx.x < n_items) ? ldg_u32_raw(a.tile_mask + blockIdx.x / n_ntiles) : 1u;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const int rt = w / n_ntiles, nt = w - rt * n_ntiles;
      uint32_t m = (has_table && mnext) ? mnext : 1u;
      if (has_table && w + (int)gridDim.x < n_items) mnext = ldg_u32_raw(a.tile_mask + (w + (int)gridDim.x) / n_ntiles);
      const uint8_t* wp = (const uint8_t*)a.w_packed_ts + (size_t)nt * a.K * p.n_chunks * B_BYTES + lane * 16;
      while (m) {
        const int k = __ffs(m) - 1;
        m &= m - 1u;
        for (int c = 0; c < p.n_chunks; ++c, ++g) {
          if ((g & 1u) != (uint32_t)bw) continue;
          mbar_wait(&empty[s], ph);
          const uint8_t* src = wp + ((size_t)k * p.n_chunks + c) * B_BYTES;
          if (kBulkB) {
            if (elect_one_sync()) {
              mbar_arrive_expect_tx(&full[s], B_BYTES);
              bulk_copy_g2s(sB + s * B_BYTES, src - lane * 16, B_BYTES, &full[s]);
            }
          } else {
            const uint32_t dst = sB_u32 + s * (uint32_t)B_BYTES;
#pragma unroll
            for (int i = 0; i < kPer; ++i) cp_async_16(dst + (uint32_t)i * 512u, src + i * 512);
            cp_async_mbar_arrive_noinc(&full[s]);
          }
          gtr += 2;
          s += 2u;
          if (s >= STAGES) { s -= STAGES; ph ^= 1u; }
        }
      }
    }
   } else if (warp == kTsWarpMma) {
    // ================================================================= MMA issue (whole warp converged, elect.sync)
    const uint64_t bdesc0 = umma_desc_sw128(smem_u32(sB));
    const uint32_t empty_u32 = smem_u32(empty);
    int it = 0;
    uint32_t gbase = 0;             // K-steps issued so far by this CTA: ring stage = gbase % STAGES, use = gbase / STAGES
    uint32_t mnext = (has_table && (int)blockIdx.x < n_items) ? ldg_u32_raw(a.tile_mask + blockIdx.x / n_ntiles) : 1u;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const int nsteps = __popc((has_table && mnext) ? mnext : 1u) * p.n_chunks;
      if (has_table && w + (int)gridDim.x < n_items) mnext = ldg_u32_raw(a.tile_mask + (w + (int)gridDim.x) / n_ntiles);
      const int buf = it & 1;
      if (it >= 2) mbar_wait(&acc_empty[buf], (uint32_t)((it >> 1) - 1) & 1u);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * N_TILE);
      if (it < 64) UD3D_TS_TR(2048 + 4 * it + 2);
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
      if (!UD3D_DBG(p, 1)) {
        umma_commit_elect(&acc_full[buf]);
      } else if (lane == 0) {
        mbar_arrive(&acc_full[buf]);
      }
    }
    __syncwarp();
   }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
  if (UD3D_TRACE_BUF(p) && p.trace_block == -3 && tid == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.trace[4 * blockIdx.x + 1] = (long long)gt;
  }
}

template <int N_TILE>
static size_t ts_smem_bytes() {
  return 1024 + (size_t)TsCfg<N_TILE>::kStages * N_TILE * 128 + kTsEpiWarps * 4096 + kTsProdWarps * 1024 + (16 + 4) * 8 + 16;
}

template <int N_TILE>
static int launch_ts(const GemmParams& p, int num_sms, cudaStream_t st) {
  const size_t smem = ts_smem_bytes<N_TILE>();
  // per-device configuration (cudaFuncSetAttribute applies to the current device)
  DeviceCtx* ctx = device_ctx();
  if (!ctx) return UD3D_ECUDA;
  {
    CtxGuard guard(ctx);
    if (ctx_needs_config(ctx, (const void*)gather_gemm_ts_kernel<N_TILE>, smem)) {
    UD3D_CUDA(cudaFuncSetAttribute(gather_gemm_ts_kernel<N_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // leave the rest of the 256 KB to L1: gathered rows are re-used across the kernel offsets of a tile
    int pct = (int)((smem + 1024) * 100 / (228 * 1024)) + 1;
    if (pct > 100) pct = 100;
    UD3D_CUDA(cudaFuncSetAttribute(gather_gemm_ts_kernel<N_TILE>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    }
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
