// Gather-GEMM: sparse convolution (SubM / strided / inverse) and dense linear layers on the
// 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM), sm_100a.
//
//   out[o,:] = act( sum_k pre(in[table[k][o],:]) @ W_k + bias ) + residual[o,:]
//
// Design (one CTA = 128 output rows x N_TILE output channels, 10 warps, warp-specialised):
//   * the K-loop runs over (kernel offset k, 32-channel chunk c).  Operand-form inputs (pre-activated bf16 hi|lo rows,
//     128 bytes per row-chunk) are gathered with 16-byte cp.async copies (zero-fill for missing neighbours) straight
//     into a 128B-swizzled K-major smem tile; the producer warps are bound to ring stages and never wait for their own
//     copies (cp.async.mbarrier.arrive.noinc on the stage's full barrier).  fp32 inputs take a register path (coalesced
//     16-byte loads, folded BatchNorm affine + ReLU, bf16 hi/lo split, st.shared);
//   * the matching weight tile (pre-packed on the host side of the ABI into the identical swizzled image, hi|lo per
//     output channel) is fetched by ONE cp.async.bulk (TMA engine) onto the same full barrier;
//   * the MMA warp issues 6 tcgen05.mma (M=128, N=N_TILE, K=16; elect.sync): hi*hi, lo*hi, hi*lo -> fp32 accumulation
//     in TMEM (~2e-5 relative error end to end, vs 1.5e-3 for single-pass TF32), then tcgen05.commit -> empty[stage];
//   * a 3-5 stage smem ring linked only by mbarriers -- no block barrier in the mainloop; the tile's rulebook slice is
//     staged in smem first; offsets with no active input in the tile are skipped (rulebook tile mask);
//   * launches that cannot fill 148 SMs are split along K across the CTAs of a (1,1,z) thread-block cluster, which
//     reduce their partial tiles through distributed shared memory (deterministic, no atomics, no second pass);
//   * epilogue: tcgen05.ld (32 lanes x 32 columns) -> bias / activation / residual -> fp32 and/or operand-form stores.
// HBM/L2 roofline: see DESIGN.md section 4 (algorithmic bytes per layer).
#include "gemm_common.cuh"
#include <stdlib.h>


namespace ud3d {

static long long* g_trace = nullptr;
static int g_trace_block = 0;
static int g_dbg = 0;

// gemm_ts.cu
int launch_gemm_ts(const GemmParams& p, int nts, int num_sms, cudaStream_t st);
int launch_pack_weight_ts(const float* w, int K, int c_in, int c_out, int nts, void* packed, cudaStream_t st);

static inline int pick_ntile(int c_out) {
  if (c_out <= 32) return 32;
  if (c_out <= 64) return 64;
  if (c_out <= 96) return 96;
  if (c_out <= 128) return 128;
  if (c_out <= 160) return 160;
  // wide dense GEMMs (encoder): 128-column tiles keep 2 CTAs/SM (measured 7 % faster end to end than 256)
  return 128;
}

// ---------------------------------------------------------------- weight packing
// Two images of the same size (N_TILE x 128 bytes per (n_tile, k, chunk)):
//  * stacked (N_TILE <= 128, kStackedB): 2 N_TILE rows of 64 bytes -- row n = bf16 hi of weight row n (32 channels),
//    row N_TILE + n = its bf16 lo -- 16-byte piece j stored at j ^ ((row >> 1) & 3) (Swizzle<2,4,3>, 64B rows).  One
//    MMA of N = 2 N_TILE then forms A_hi . [W_hi | W_lo] in one pass over the A tile (see the MMA issuer);
//  * side by side (N_TILE = 160): [N_TILE rows][8 x 16B], row n holds hi(32 ch) | lo(32 ch), 16-byte
//    piece j stored at physical position j ^ (n & 7)  (Swizzle<3,4,3>, 128B rows)
__host__ __device__ constexpr bool stacked_b(int n_tile) { return n_tile <= 128; }
__global__ void pack_weight_kernel(const float* __restrict__ w, int K, int c_in, int c_out, int n_tile_sz, int n_tiles,
                                   int n_chunks, uint4* __restrict__ out) {
  long long total = (long long)n_tiles * K * n_chunks * n_tile_sz * 8;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int j = (int)(t & 7);
    long long r = t >> 3;
    int n = (int)(r % n_tile_sz);
    r /= n_tile_sz;
    int c = (int)(r % n_chunks);
    r /= n_chunks;
    int k = (int)(r % K);
    int nt = (int)(r / K);
    int ng = nt * n_tile_sz + n;
    int ch0 = c * kChunk + (j & 3) * 8;
    uint32_t v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float x[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int ch = ch0 + e * 2 + h;
        x[h] = (ng < c_out && ch < c_in) ? w[((size_t)ng * K + k) * c_in + ch] : 0.f;
      }
      uint32_t hi, lo;
      split_bf16x2(x[0], x[1], hi, lo);
      v[e] = (j < 4) ? hi : lo;
    }
    size_t tile = ((size_t)nt * K + k) * n_chunks + c;
    if (stacked_b(n_tile_sz)) {
      const int row = (j < 4) ? n : n_tile_sz + n;
      out[tile * n_tile_sz * 8 + (size_t)row * 4 + ((j & 3) ^ ((row >> 1) & 3))] = make_uint4(v[0], v[1], v[2], v[3]);
    } else {
      out[tile * n_tile_sz * 8 + (size_t)n * 8 + (j ^ (n & 7))] = make_uint4(v[0], v[1], v[2], v[3]);
    }
  }
}

// position of tile row r inside an offset's 128-entry slice of the smem rulebook: lane group g = r & 3 owns the 32
// consecutive entries of rows g, g + 4, g + 8, ... (a producer lane reads its 32 source rows with 8 x LDS.128)
__device__ __forceinline__ int tbl_pos(int r) { return ((r & 3) << 5) + (r >> 2); }

constexpr int kProducerWarps = 8;                 // warps 0..7 gather A (warps 0..3 also run the epilogue)
constexpr int kWarpB = 8;                         // warp 8: weight-tile bulk copies
// warp 9: tcgen05.mma issue
constexpr int kThreadsTc = 32 * 10;
// One CTA per row tile.  (A persistent variant of the same loop -- gridDim.x = 2 CTAs per SM striding over the tiles --
// measured 10 % slower: the second resident CTA already overlaps prologue / epilogue, and the loop-carried state
// costs registers.)
constexpr bool kPersistent = false;

// Ring depth S and copies-in-flight D per instantiation.  A stage cycles fill (L2 latency L) -> MMA round
// trip (M: a_full arrive -> issue -> tcgen05.commit -> empty) -> refill, so steady-state step time per CTA is
// bounded by  D * T >= L  and  (S - D - 1) * T >= M  (a stage published at iteration t + D must be free again at
// t + S).  Traced on B200: L ~ M ~ 1000 cycles, so D = 1 with S = 4 (T >= M / 2) beats D = 2 (T >= M) and
// D = 3 (every step pays the full round trip).  S is the largest depth that keeps 2 CTAs/SM (1 for N_TILE >= 160).
template <int N_TILE> struct TcCfg {
  static constexpr int kStages = N_TILE <= 32 ? 3 : N_TILE <= 64 ? 4 : N_TILE <= 128 ? 3 : N_TILE == 160 ? 5 : 4;
  // resident CTAs per SM the kernel is compiled for (register cap).  (3 CTAs x 3 stages for N_TILE = 32 and 3 CTAs x 2
  // stages for N_TILE = 64 measured the same as 2 x 4: the main loop is bound by the gather latency x the stages in flight
  // per SM, not by the number of CTAs; a 6-stage, 1 CTA/SM ring for the split-K launches lost more in waves than it won.)
  static constexpr int kMinCtas = N_TILE <= 32 ? 3 : N_TILE <= 128 ? 2 : 1;
};

// Raw gathered operand of one K-step for this thread: 2 rows x 8 channels
struct GatherRegs {
  float4 v[2][2];
  int ok[2];
};

// Persistent: gridDim.x CTAs (2 per SM) walk the row tiles with stride gridDim.x; barriers, TMEM and the smem ring
// are set up once per CTA and the ring state (stage, use count) runs on across tiles.  (One CTA per tile cost ~3 us of
// launch / TMEM-allocation turnaround per tile: a third of the level-1 convolutions' time.)
// debug timeline (ud3d_debug_set_trace(buf, -2)): globaltimer of phase `slot` of every CTA (thread 0), 8 slots per CTA
#define UD3D_TL(slot)                                                                                        \
  do {                                                                                                       \
    if (UD3D_TRACE_BUF(p) && p.trace_block == -2 && tid == 0 && blockIdx.y == 0) {                                     \
      unsigned long long gt__;                                                                               \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt__));                                               \
      p.trace[8 * (blockIdx.x * gridDim.z + blockIdx.z) + (slot)] = (long long)gt__;                         \
    }                                                                                                        \
  } while (0)

template <int N_TILE>
__global__ void __launch_bounds__(kThreadsTc, TcCfg<N_TILE>::kMinCtas) gather_gemm_tc_kernel(const GemmParams p) {
  constexpr int STAGES = TcCfg<N_TILE>::kStages;
  constexpr int A_BYTES = kTileM * 128;
  constexpr int B_BYTES = N_TILE * 128;
  // stacked weight tile: the accumulator has 2 N_TILE columns -- [0, N) = hi.hi + lo.hi, [N, 2N) = hi.lo -- summed in
  // the epilogue
  constexpr bool STACKED = stacked_b(N_TILE);
  constexpr int ACC_COLS = STACKED ? 2 * N_TILE : N_TILE;
  constexpr uint32_t TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : 256;
  constexpr uint32_t IDESC = umma_idesc_bf16_m128(N_TILE);
  constexpr uint32_t IDESC2 = umma_idesc_bf16_m128(2 * N_TILE <= 256 ? 2 * N_TILE : 256);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // (the dynamic shared memory window starts 1024-byte aligned -- the kernel has no static shared memory -- and the
  // launch does not pay a kilobyte of slack for it: three CTAs of the N_TILE = 32 instantiation fit one SM)
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
  uint8_t* smem = smem_raw;
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint8_t* tail = sB + STAGES * B_BYTES;
  uint64_t* a_full = (uint64_t*)tail;            // [STAGES] count = producer arrivals + 1 (weight copy, +tx bytes)
  uint64_t* empty = a_full + 2 * STAGES;         // [STAGES] count = 1 (tcgen05.commit)
  uint64_t* acc_full = empty + STAGES;           // [1]
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);
  uint32_t* s_mask = tmem_slot + 1;
  uint8_t* s_actk = (uint8_t*)(tail + 128);      // [32] active kernel offsets, ascending
  float* s_scale = (float*)(tail + 192);
  const ud3d_gemm_args& a = p.a;
  const int c_in_pad = p.n_chunks * kChunk;
  float* s_shift = s_scale + c_in_pad;
  int32_t* s_tbl = (int32_t*)(s_shift + c_in_pad);   // [K][128] table slice of the current tile

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (UD3D_TRACE_BUF(p) && (int)blockIdx.x == p.trace_block && tid == 0) p.trace[1014] = clock64();
  const int nt = blockIdx.y;
  const int n0 = nt * N_TILE;
  const bool has_table = a.table != nullptr;
  const int n_row_tiles = (a.n_out + kTileM - 1) / kTileM;

  // ------------------------------------------------------------ once per CTA
  UD3D_TL(0);
  if (UD3D_TRACE_BUF(p) && p.trace_block == -2 && tid == 0 && blockIdx.y == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    p.trace[8 * (blockIdx.x * gridDim.z + blockIdx.z) + 2] = smid;
  }
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      // operand-form input: every producer thread's async copies arrive by themselves (cp.async.mbarrier.arrive.noinc)
      mbar_init(&a_full[s], (a.in_split ? 32 * (kProducerWarps / STAGES) : kProducerWarps) + 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
    *s_mask = 0u;
  }
  if (a.in_scale) {
    for (int c = tid; c < c_in_pad; c += kThreadsTc) {
      s_scale[c] = c < a.c_in ? a.in_scale[c] : 0.f;
      s_shift[c] = c < a.c_in ? a.in_shift[c] : 0.f;
    }
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (UD3D_TRACE_BUF(p) && (int)blockIdx.x == p.trace_block && tid == 0) p.trace[1015] = clock64();
  const uint8_t* wp = (const uint8_t*)a.w_packed + (size_t)nt * a.K * p.n_chunks * B_BYTES;

  // ring state: every role walks the same (stage, use) sequence, across tiles
  int rs = 0;
  uint32_t ruse = 0;
  uint32_t acc_phase = 0;

  for (int tile = blockIdx.x, it = 0; tile < n_row_tiles; tile += gridDim.x, ++it) {
    const int m0 = tile * kTileM;
    const bool traced = UD3D_TRACE_BUF(p) && (int)blockIdx.x == p.trace_block && it == p.trace_iter;
    if (traced && tid == 0) p.trace[1019] = clock64();
    if (traced && tid == 0) p.trace[1016] = p.trace[1015];
    // ---------------------------------------------------------- per tile: rulebook slice -> smem, active offsets
    // (one coalesced pass; removes the dependent table -> row load chain from the mainloop)
    const uint32_t tmask = (has_table && a.tile_mask) ? __ldg(a.tile_mask + tile) : 0u;
    uint32_t mybits = 0;
    if (has_table && !kPersistent && blockIdx.y == 0 && blockIdx.z == 0) {
      // pull the rulebook slice (and output-row list) of the tile that will run on this SM about one CTA life from now
      // towards L2: the slice is the first thing a CTA waits for, and it comes from DRAM on every launch (the table is
      // larger than what stays cached between convs)
      const int ahead = tile + 3 * 148;
      if (ahead < n_row_tiles) {
        const int lines = a.K * 4;                                     // 128 rows x 4 B = 4 lines of 128 B per offset
        if (tid < lines)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a.table + (size_t)(tid >> 2) * a.n_out + (size_t)ahead * kTileM + (tid & 3) * 32));
        else if (a.row_perm && tid < lines + 4)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a.row_perm + (size_t)ahead * kTileM + (tid - lines) * 32));
      }
    }
    if (has_table) {
      // all loads of a thread are issued back to back (one L2 round trip for the whole slice)
      constexpr int kPer = (32 * kTileM + kThreadsTc - 1) / kThreadsTc;   // 13
      int vals[kPer];
      const int total = a.K * kTileM;
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const int i = tid + j * kThreadsTc;
        const int k = i >> 7, row = m0 + (i & 127);
        vals[j] = (i < total && row < a.n_out) ? __ldg(a.table + (size_t)k * a.n_out + row) : -1;
      }
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const int i = tid + j * kThreadsTc;
        if (i < total) {
          s_tbl[(i & ~127) + tbl_pos(i & 127)] = vals[j];
          if (vals[j] >= 0) mybits |= 1u << (i >> 7);
        }
      }
    } else if (a.in_split) {
      if (tid < kTileM) s_tbl[tbl_pos(tid)] = (m0 + tid < a.n_out) ? m0 + tid : -1;   // identity gather as a one-offset table
    }
    uint32_t mask;
    if (!has_table) {
      mask = 1u;
    } else if (a.tile_mask) {
      mask = tmask;
    } else {
      mybits = __reduce_or_sync(0xffffffffu, mybits);
      if (lane == 0 && mybits) atomicOr(s_mask, mybits);
      __syncthreads();
      mask = *s_mask;
    }
    if (traced && tid == 0) p.trace[1017] = clock64();
    if (tid < 32 && ((mask >> tid) & 1u)) s_actk[__popc(mask & ((1u << tid) - 1u))] = (uint8_t)tid;
    __syncthreads();
    if (traced && tid == 0) p.trace[1018] = clock64();
    UD3D_TL(3);      // prologue done (rulebook slice staged)
    const int nact = __popc(mask);
    const int nsteps_all = nact * p.n_chunks;
    // split-K: this CTA handles steps [t_begin, t_end) of the tile's active (offset, chunk) sequence
    const int t_begin = gridDim.z == 1 ? 0 : (int)((unsigned)nsteps_all * blockIdx.z / gridDim.z);
    const int t_end = gridDim.z == 1 ? nsteps_all : (int)((unsigned)nsteps_all * (blockIdx.z + 1) / gridDim.z);
    const int nsteps = UD3D_DBG(p, 512) ? 0 : t_end - t_begin;
    // (kslot, chunk) of this CTA's first step: the only integer division of the tile
    const int kslot0 = t_begin / p.n_chunks;
    const int chunk0 = t_begin - kslot0 * p.n_chunks;
    // epilogue row of this thread (warps 0..7: TMEM lane quadrant warp & 3), loaded now so that the epilogue does not
    // start with a dependent global round trip.  Regrouped rows (ud3d_subm3_tile_order): position m0 + row of the
    // table stands for output row row_perm[m0 + row]
    const int epi_row = (warp & 3) * 32 + lane;
    const bool epi_row_ok = warp < 8 && m0 + epi_row < a.n_out;
    const int epi_grow = (epi_row_ok && a.row_perm) ? __ldg(a.row_perm + m0 + epi_row) : m0 + epi_row;

    if (warp < kProducerWarps) {
      // ========================================================= A producers
      if (a.in_split) {
        // ---- operand-form input: pure async copies (LDGSTS), 16 B per lane, 8 lanes per 128-byte row-chunk,
        //      zero-fill for missing neighbours, no ALU work.  The producers never wait for their own copies: each
        //      thread's cp.async.mbarrier.arrive.noinc makes the stage's full barrier count its arrival when its copies
        //      have landed, so the ring depth hides the L2 latency: wait empty -> 4 LDGSTS -> arrive.noinc.
        const uint32_t row_bytes = (uint32_t)a.ld_in * 4u;
        // The producer warps are bound to ring stages: stage s is filled, for every K-step that maps to it, by the same
        // WPG = 8 / STAGES warps (each copies 128 / WPG rows: NI = 32 / WPG LDGSTS per lane).  One pass through the
        // control code (empty wait, arrive) per NI copies instead of per 4 -- the loop is bound by the LSU (~8 cycles per
        // LDGSTS instruction), not by each warp's serial instruction latency -- the stages fill concurrently, and
        // every waiter sees all phases of its barrier in order (a parity wait cannot tell phases two apart).
        // Lane (g = lane >> 3, j = lane & 7) copies chunk j of rows base + g, base + g + 4, ...: instruction i covers
        // the 4 consecutive rows base + 4 i .. base + 4 i + 3.
        constexpr int WPG = kProducerWarps / STAGES;
        constexpr int NI = 32 / WPG;
        const int grp = warp / WPG, half = warp - grp * WPG;
        const int g = lane >> 3, j = lane & 7;
        // lane j copies piece j of the row-chunk to chunk j of the K-major tile row (opf_mem_piece is the identity)
        const uint32_t sw0 = (uint32_t)((j ^ g) << 4), sw1 = (uint32_t)(((j ^ g) ^ 4) << 4);   // (4 i + g) & 7 = 4 (i & 1) + g
        const uint8_t* src_base = (const uint8_t*)a.in + opf_mem_piece(j) * 16;
        const uint32_t sA_lane = smem_u32(sA) + (uint32_t)((half * (kTileM / WPG) + g) * 128);
        long long* tr = (traced && tid == 0) ? UD3D_TRACE_BUF(p) : nullptr;
        if (tr) { tr[1023] = nsteps; tr[1022] = clock64(); }
        if (grp < STAGES) {
          for (int t = (grp - rs + STAGES) % STAGES; t < nsteps; t += STAGES) {     // the steps that land in stage grp
            if (tr && t < 64) tr[t * 8 + 0] = clock64();
            const int s = grp;
            const uint32_t use = ruse + (uint32_t)((rs + t) / STAGES);
            const int tt = t_begin + t;
            const int kslot = tt / p.n_chunks;
            const int c = tt - kslot * p.n_chunks;
            const int k = has_table ? (int)s_actk[kslot] : 0;
            int idx[NI];
            {
              const int4* tp = (const int4*)(s_tbl + k * kTileM + g * 32 + half * NI);
#pragma unroll
              for (int q = 0; q < NI / 4; ++q) {
                const int4 v = tp[q];
                idx[4 * q] = v.x; idx[4 * q + 1] = v.y; idx[4 * q + 2] = v.z; idx[4 * q + 3] = v.w;
              }
            }
            if (use) mbar_wait(&empty[s], (use & 1u) ^ 1u);
            if (tr && t < 64) tr[t * 8 + 1] = clock64();
            const uint8_t* sb = src_base + c * 128;
            const uint32_t as_addr = sA_lane + s * A_BYTES;
            // branch-free: a missing neighbour is a zero-fill copy (src-size 0: no global read)
            if UD3D_DBG(p, 128) {
#pragma unroll
              for (int i = 0; i < NI; ++i) {
                const int ix = idx[i];
                cp_async_16_zfill_ca(as_addr + i * 512 + ((i & 1) ? sw1 : sw0),
                                     sb + (size_t)((uint32_t)(ix < 0 ? 0 : ix) * (uint64_t)row_bytes), ix < 0 ? 0u : 16u);
              }
            } else if (!UD3D_DBG(p, 4)) {
#pragma unroll
              for (int i = 0; i < NI; ++i) {
                const int ix = UD3D_DBG(p, 32) ? -1 : idx[i];
                cp_async_16_zfill(as_addr + i * 512 + ((i & 1) ? sw1 : sw0),
                                  sb + (size_t)((uint32_t)(ix < 0 ? 0 : ix) * (uint64_t)row_bytes), ix < 0 ? 0u : 16u);
              }
            }
            cp_async_mbar_arrive_noinc(&a_full[s]);
            if (tr && t < 64) tr[t * 8 + 2] = tr[t * 8 + 3] = clock64();
          }
        }
      } else {
        // ---- fp32 input: gather + folded BN/ReLU + bf16 hi/lo split in registers (4 lanes per row, 8 channels each)
        const int q = tid & 3;
        const int rl = tid >> 2;
        const bool affine = a.in_scale != nullptr;
        const bool relu = a.in_relu != 0;
        int ident[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) ident[h] = (m0 + rl + h * 64 < a.n_out) ? m0 + rl + h * 64 : -1;

        auto issue_loads = [&](int kslot, int c, GatherRegs& g) {
          const int32_t* trow = s_tbl + (int)s_actk[kslot] * kTileM;
          const int ch0 = c * kChunk + q * 8;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int idx = has_table ? trow[tbl_pos(rl + h * 64)] : ident[h];
            g.ok[h] = idx >= 0 && ch0 < a.c_in;
            if (g.ok[h]) {
              const float* src = a.in + (size_t)idx * a.ld_in + ch0;
              if (p.vec_ok) {
                g.v[h][0] = __ldg((const float4*)src);
                g.v[h][1] = __ldg((const float4*)src + 1);
              } else {
                float e[8];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) e[jj] = (ch0 + jj < a.c_in) ? __ldg(src + jj) : 0.f;
                g.v[h][0] = make_float4(e[0], e[1], e[2], e[3]);
                g.v[h][1] = make_float4(e[4], e[5], e[6], e[7]);
              }
            }
          }
        };
        auto store_step = [&](int s, uint32_t use, int c, const GatherRegs& g) {
          const int ch0 = c * kChunk + q * 8;
          if (use) {
            if (lane == 0) mbar_wait(&empty[s], (use & 1u) ^ 1u);
            __syncwarp();
          }
          uint8_t* As = sA + s * A_BYTES;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float v[8];
            if (g.ok[h]) {
              v[0] = g.v[h][0].x; v[1] = g.v[h][0].y; v[2] = g.v[h][0].z; v[3] = g.v[h][0].w;
              v[4] = g.v[h][1].x; v[5] = g.v[h][1].y; v[6] = g.v[h][1].z; v[7] = g.v[h][1].w;
              if (affine) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e], s_scale[ch0 + e], s_shift[ch0 + e]);
              }
              if (relu) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
              }
              if (!p.vec_ok) {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                  if (ch0 + e >= a.c_in) v[e] = 0.f;     // padded channels stay exactly zero
              }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = 0.f;
            }
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_bf16x2(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
            const int r = rl + h * 64;
            uint8_t* arow = As + r * 128;
            *(uint4*)(arow + ((q ^ (r & 7)) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *(uint4*)(arow + (((4 + q) ^ (r & 7)) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_full[s]);
        };

        // software pipeline: the global loads of step t+1 are in flight while step t is converted/stored
        GatherRegs g0, g1;
        int lk = kslot0, lc = chunk0;        // (kslot, chunk) of the next loads to issue
        int sc = chunk0;                     // chunk of the next step to store
        int s = rs;
        uint32_t use = ruse;
        auto adv_load = [&]() { if (++lc == p.n_chunks) { lc = 0; ++lk; } };
        auto adv_store = [&]() { if (++sc == p.n_chunks) sc = 0; if (++s == STAGES) { s = 0; ++use; } };
        if (nsteps > 0) { issue_loads(lk, lc, g0); adv_load(); }
        for (int t = 0; t < nsteps; t += 2) {
          if (t + 1 < nsteps) { issue_loads(lk, lc, g1); adv_load(); }
          store_step(s, use, sc, g0); adv_store();
          if (t + 1 < nsteps) {
            if (t + 2 < nsteps) { issue_loads(lk, lc, g0); adv_load(); }
            store_step(s, use, sc, g1); adv_store();
          }
        }
      }
    } else if (warp == kWarpB) {
      // ========================================================= B producer: one bulk copy (TMA engine) per step; the
      // warp stays converged and an elected lane issues (uniform-register operands)
      int kslot = kslot0, c = chunk0, s = rs;
      uint32_t use = ruse;
      for (int t = 0; t < nsteps; ++t) {
        if (use) mbar_wait(&empty[s], (use & 1u) ^ 1u);
        const int k = s_actk[kslot];
        if (elect_one_sync()) {
          if UD3D_DBG(p, 2) {
            mbar_arrive(&a_full[s]);
          } else {
            mbar_arrive_expect_tx(&a_full[s], B_BYTES);
            bulk_copy_g2s(sB + s * B_BYTES, wp + ((size_t)k * p.n_chunks + c) * B_BYTES, B_BYTES, &a_full[s]);
          }
        }
        if (++c == p.n_chunks) { c = 0; ++kslot; }
        if (++s == STAGES) { s = 0; ++use; }
      }
      __syncwarp();
    } else {
      // ========================================================= MMA issuer.  The whole warp runs the loop converged and
      // one elected lane issues each tcgen05 instruction (elect.sync): operands stay in uniform registers.  Under a
      // `lane == 0` branch the compiler wraps every UTCHMMA in an R2UR / vote loop (~80 cycles of issue per MMA).
      const uint64_t adesc0 = umma_desc_sw128(smem_u32(sA));
      const uint64_t bdesc0 = STACKED ? umma_desc_sw64(smem_u32(sB)) : umma_desc_sw128(smem_u32(sB));
      long long* tr = (traced && lane == 0) ? UD3D_TRACE_BUF(p) : nullptr;
      int s = rs;
      uint32_t use = ruse;
      for (int t = 0; t < nsteps; ++t) {
        mbar_wait(&a_full[s], use & 1u);   // A rows (producer threads) + weight tile (bulk copy tx bytes)
        if (tr && t < 64) tr[t * 8 + 4] = tr[t * 8 + 5] = clock64();
        tc_fence_after_sync();
        if UD3D_DBG(p, 1) {
          if (lane == 0) mbar_arrive(&empty[s]);
        } else {
          // the start-address field is in 16-byte units: stage s, then +2 = 32 B (second K=16 slice), +4 = lo half,
          // +6 = lo second slice
          const uint64_t ad = adesc0 + (uint64_t)(s * (A_BYTES >> 4)), bd = bdesc0 + (uint64_t)(s * (B_BYTES >> 4));
          if constexpr (STACKED) {
            // 4 MMAs: A_hi x [W_hi | W_lo] (N = 2 N_TILE: columns [0, N) += hi.hi, [N, 2N) += hi.lo) and A_lo x W_hi
            // (N = N_TILE, the first N rows of the same weight tile) per K = 16 slice.  The A_hi tile is read from shared
            // memory once instead of twice: the main loop is bound by the shared-memory port (operand reads of the
            // tensor core + the gather's writes), not by the tensor pipe
            umma_bf16_elect(tmem_base, ad + 0, bd + 0, IDESC2, t > 0);
            umma_bf16_elect(tmem_base, ad + 2, bd + 2, IDESC2, 1);
            umma_bf16_elect(tmem_base, ad + 4, bd + 0, IDESC, 1);
            umma_bf16_elect(tmem_base, ad + 6, bd + 2, IDESC, 1);
          } else {
            umma_bf16_elect(tmem_base, ad + 0, bd + 0, IDESC, t > 0);
            umma_bf16_elect(tmem_base, ad + 2, bd + 2, IDESC, 1);
            umma_bf16_elect(tmem_base, ad + 4, bd + 0, IDESC, 1);
            umma_bf16_elect(tmem_base, ad + 6, bd + 2, IDESC, 1);
            umma_bf16_elect(tmem_base, ad + 0, bd + 4, IDESC, 1);
            umma_bf16_elect(tmem_base, ad + 2, bd + 6, IDESC, 1);
          }
          umma_commit_elect(&empty[s]);      // stage s reusable once these MMAs have read it
        }
        if (tr && t < 64) tr[t * 8 + 6] = clock64();
        if (++s == STAGES) { s = 0; ++use; }
      }
      if (nsteps > 0) {
        if UD3D_DBG(p, 1) {
          if (lane == 0) mbar_arrive(acc_full);
        } else {
          umma_commit_elect(acc_full);
        }
      }
      __syncwarp();
    }
    // every role advanced the ring by nsteps
    {
      const int adv = rs + nsteps;
      ruse += (uint32_t)(adv / STAGES);
      rs = adv % STAGES;
    }

    // ---------------------------------------------------------- epilogue (warps 0..3, one row per thread)
    const bool split = gridDim.z > 1;
    if (!split) {
      // warps 0-3 and (tiles wider than 32 columns) the otherwise idle producer warps 4-7: warp w reads TMEM lanes
      // 32 * (w % 4) .. + 31 (= its 32 tile rows) and the two warps of a lane quadrant take alternate 32-column chunks.
      // One row per thread; the per-chunk chain (tcgen05.ld -> bias / residual loads -> activation -> hi/lo split ->
      // 128-byte row stores) is latency-bound, so doubling the warps nearly halves the epilogue of the wide GEMMs.
      constexpr int kEpiWarps = N_TILE > 32 ? 8 : 4;
      if (warp < kEpiWarps) {
        // the producers run ahead of the MMAs by the ring depth: while the last stages drain, pull this thread's residual
        // row segments towards L2 (a DRAM round trip otherwise paid after the accumulator is complete)
        if (a.residual && epi_row_ok) {
#pragma unroll 1
          for (int c0 = (warp >> 2) * 32; c0 < N_TILE && n0 + c0 < a.c_out; c0 += 32 * (kEpiWarps / 4))
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.residual + (size_t)epi_grow * a.ld_res + n0 + c0));
        }
        if (nsteps > 0) {
          mbar_wait(acc_full, acc_phase);
          tc_fence_after_sync();
        }
        if (traced && tid == 0) p.trace[1021] = clock64();
        UD3D_TL(4);  // main loop done
        const int quad = warp & 3;
        const bool row_ok = epi_row_ok;
        const int grow = epi_grow;
#pragma unroll 1
        for (int c0 = (warp >> 2) * 32; c0 < N_TILE; c0 += 32 * (kEpiWarps / 4)) {
          uint32_t r[32];
          if (nsteps > 0) {
            tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, r);
            if constexpr (STACKED) {
              uint32_t r2[32];
              tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(N_TILE + c0), r2);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
            } else {
              tmem_ld_wait();
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = 0u;
          }
          epilogue_store_chunk(p, r, grow, n0 + c0, row_ok && !UD3D_DBG(p, 16), smem_u32(sA) + (uint32_t)warp * 4096u, lane);
        }
      }
    } else {
      // ---- split-K: the gridDim.z CTAs of one (1, 1, z) thread-block cluster hold partial sums of the same output tile.
      //      Each parks its fp32 partial in its own (now idle) operand ring, then CTA `rank` sums rows
      //      [rank * rp, (rank + 1) * rp) over all peers through distributed shared memory, in rank order (deterministic),
      //      and runs the normal epilogue on them: no atomics, no pre-zeroed output, no second pass.
      //      Partial tile: [128 rows][N_TILE] fp32, 16-byte chunks XOR-swizzled with (row & 7) (conflict-free stores).
      const uint32_t sP = smem_u32(sA);
      if (warp < 4) {
        if (nsteps > 0) {
          mbar_wait(acc_full, acc_phase);
          tc_fence_after_sync();
        }
        UD3D_TL(4);  // main loop done
        const int row = warp * 32 + lane;
        const uint32_t prow = sP + (uint32_t)(row * N_TILE * 4);
#pragma unroll 1
        for (int c0 = 0; c0 < N_TILE; c0 += 32) {
          uint32_t r[32];
          if (nsteps > 0) {
            tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
            if constexpr (STACKED) {
              uint32_t r2[32];
              tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(N_TILE + c0), r2);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
            } else {
              tmem_ld_wait();
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = 0u;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + (uint32_t)(c0 * 4) + (uint32_t)((j ^ (row & 7)) << 4)),
                         "r"(r[4 * j]), "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                         : "memory");
        }
      }
      UD3D_TL(5);    // partial parked
      cluster_arrive_release();
      cluster_wait_acquire();
      UD3D_TL(6);    // all partials visible
      constexpr int NC16 = N_TILE / 4;
      constexpr int NCH = N_TILE / 32;
      const int z = (int)gridDim.z;
      uint32_t rank;
      asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
      const int rp = (kTileM + z - 1) / z;
      const int r_begin = (int)rank * rp;
      const int r_end = r_begin + rp < kTileM ? r_begin + rp : kTileM;
      const int n_rows = r_end > r_begin ? r_end - r_begin : 0;
      {
        // phase 1 (all warps): thread = (row, 16-byte chunk): one float4 per peer, all loads independent (the peers are
        // visited starting at this CTA's own rank, so the cluster's CTAs do not all hit the same SM at once); the sum is
        // written back in place -- rows [r_begin, r_end) of this CTA's own partial tile are read by nobody else
        for (int item = tid; item < n_rows * NC16; item += kThreadsTc) {
          const int row = r_begin + item / NC16;
          const uint32_t off = (uint32_t)(row * N_TILE * 4) + (uint32_t)((item - (item / NC16) * NC16) << 4);   // (swizzled slot: same in every CTA)
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          int q = (int)rank;
          for (int q0 = 0; q0 < z; q0 += 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              if (q0 + u < z) v[u] = ld_shared_cluster_f4(cluster_map_shared(sP, (uint32_t)q) + off);
              if (++q == z) q = 0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              if (q0 + u < z) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
            }
          }
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sP + off), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
        }
      }
      __syncthreads();
      if (warp < 4) {
        // phase 2: thread = (row, 32-column chunk) of the reduced rows: the normal (warp-collective) epilogue
        const uint32_t stage = sP + (uint32_t)(kTileM * N_TILE * 4) + (uint32_t)warp * 4096u;
        const int n_items = n_rows * NCH;
        for (int base = warp * 32; base < n_items; base += 128) {
          const int item = base + lane;
          const bool valid_item = item < n_items;
          const int row = r_begin + (valid_item ? item / NCH : 0);
          const int cc = valid_item ? item - (item / NCH) * NCH : 0;
          const bool valid = valid_item && m0 + row < a.n_out;
          const int grow = (valid && a.row_perm) ? __ldg(a.row_perm + m0 + row) : m0 + row;
          const uint32_t src = sP + (uint32_t)(row * N_TILE * 4 + cc * 128);
          uint32_t r[32];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(r[4 * j]), "=r"(r[4 * j + 1]), "=r"(r[4 * j + 2]), "=r"(r[4 * j + 3])
                         : "r"(src + (uint32_t)((j ^ (row & 7)) << 4))
                         : "memory");
          epilogue_store_chunk(p, r, grow, n0 + cc * 32, valid, stage, lane);
        }
      }
      UD3D_TL(7);    // reduced + stored
      // no CTA may exit (or reuse its ring) while a peer still reads its partial tile
      cluster_arrive_release();
      cluster_wait_acquire();
    }
    if (nsteps > 0) acc_phase ^= 1u;
    if (traced && tid == 0) p.trace[1020] = clock64();
    // the next tile overwrites the rulebook slice, the active-offset list and the TMEM accumulator
    if (tid == 0) *s_mask = 0u;
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (!kPersistent) break;
  }
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
  UD3D_TL(1);
}

// ---------------------------------------------------------------- fp32 -> operand form (one warp per 4 row-chunks)
__global__ void __launch_bounds__(256) act_split_kernel(const float* __restrict__ raw, int ld_raw, int n, int c,
                                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                                        int relu, float* __restrict__ out, int ld_out, int vec) {
  // thread -> (row, chunk, 8-channel segment q): reads 32 B, writes 16 B hi + 16 B lo.  c need not be a multiple of 32:
  // the channels of the last chunk beyond c are written as zeros (e.g. the 6-channel voxel features -> one 32-ch chunk)
  const int chunks = (c + kChunk - 1) / kChunk;
  const long long total = (long long)n * chunks * 4;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(t & 3);
    const long long rc = t >> 2;
    const int ch = (int)(rc % chunks);
    const int row = (int)(rc / chunks);
    const int ch0 = ch * kChunk + q * 8;
    const float* src = raw + (size_t)row * ld_raw + ch0;
    float v[8];
    if (vec) {
      float4 x0 = *(const float4*)src, x1 = *((const float4*)src + 1);
      v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w; v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = (ch0 + e < c) ? src[e] : 0.f;
    }
    if (scale) {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (ch0 + e < c) v[e] = fmaf(v[e], __ldg(scale + ch0 + e), __ldg(shift + ch0 + e));
    }
    if (relu) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_bf16x2(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
    uint4* dst = (uint4*)((uint8_t*)(out + (size_t)row * ld_out) + (size_t)ch * 128);
    dst[opf_mem_piece(q)] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    dst[opf_mem_piece(4 + q)] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

static int launch_act_split(const float* raw, int ld_raw, int n, int c, const float* scale, const float* shift, int relu,
                            float* out, int ld_out, cudaStream_t st) {
  long long total = (long long)n * cdiv(c, kChunk) * 4;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  int vec = (c % 8 == 0) && (ld_raw % 4 == 0) && (((uintptr_t)raw & 15) == 0);
  act_split_kernel<<<blocks, 256, 0, st>>>(raw, ld_raw, n, c, scale, shift, relu, out, ld_out, vec);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

// ---------------------------------------------------------------- fp32 CUDA-core cross-check kernel
// block (32 cols, 8 rows)
__global__ void gather_gemm_simt_kernel(const ud3d_gemm_args a, const float* __restrict__ w) {
  int col = blockIdx.y * 32 + threadIdx.x;
  int row = blockIdx.x * 8 + threadIdx.y;
  if (row >= a.n_out || col >= a.c_out) return;
  float acc = 0.f;
  for (int k = 0; k < a.K; ++k) {
    int idx = a.table ? a.table[(size_t)k * a.n_out + row] : row;
    if (idx < 0) continue;
    const float* src = a.in + (size_t)idx * a.ld_in;
    const float* wk = w + ((size_t)col * a.K + k) * a.c_in;
    for (int ci = 0; ci < a.c_in; ++ci) {
      float v = src[ci];
      if (a.in_scale) v = fmaf(v, a.in_scale[ci], a.in_shift[ci]);
      if (a.in_relu) v = fmaxf(v, 0.f);
      acc = fmaf(v, wk[ci], acc);
    }
  }
  if (a.bias) acc += a.bias[col];
  acc = apply_act(acc, a.act);
  const int orow = a.row_perm ? a.row_perm[row] : row;
  if (a.residual) acc += a.residual[(size_t)orow * a.ld_res + col];
  a.out[(size_t)orow * a.ld_out + col] = acc;
}

template <int N_TILE>
static size_t tc_smem_bytes(int n_chunks, int K, bool has_table) {
  return (size_t)TcCfg<N_TILE>::kStages * (kTileM * 128 + N_TILE * 128) + 192 + (size_t)n_chunks * kChunk * 4 * 2 +
         (has_table ? (size_t)K * kTileM * 4 : 0);
}

template <int N_TILE>
static int launch_tc(const GemmParams& p, int n_tiles, int splits, cudaStream_t st) {
  size_t smem = tc_smem_bytes<N_TILE>(p.n_chunks, p.a.K, p.a.table != nullptr || p.a.in_split);
#ifdef UD3D_DEBUG_HOOKS
  smem += (size_t)((p.dbg >> 24) & 255) * 1024;       // extra shared memory (lowers the CTAs per SM)
#endif
  DeviceCtx* ctx = device_ctx();
  if (!ctx) return UD3D_ECUDA;
  const int g_num_sms = ctx_sm_count(ctx);
  int ctas_per_sm = N_TILE <= 128 ? 2 : 1;
#ifdef UD3D_DEBUG_HOOKS
  const int carve = ((p.dbg >> 16) & 255) ? ((p.dbg >> 16) & 255) : (int)cudaSharedmemCarveoutMaxShared;   // percent
#else
  const int carve = (int)cudaSharedmemCarveoutMaxShared;
#endif
  // function attributes are per device: the context remembers the largest size / the carve-out set for this kernel
  {
    CtxGuard guard(ctx);
    if (ctx_needs_config(ctx, (const void*)gather_gemm_tc_kernel<N_TILE>, smem, carve)) {
    UD3D_CUDA(cudaFuncSetAttribute(gather_gemm_tc_kernel<N_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    UD3D_CUDA(cudaFuncSetAttribute(gather_gemm_tc_kernel<N_TILE>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    }
  }
  // persistent over the row tiles: as many CTAs as are resident at once (the CTAs of the other grid dimensions
  // share the same SMs)
  const int row_tiles = cdiv(p.a.n_out, kTileM);
  int gx = row_tiles;
  if (kPersistent) {
    const int want = N_TILE <= 128 ? 2 : 1;     // (the occupancy query under-reports the 2 resident CTAs)
    gx = (g_num_sms * (ctas_per_sm > want ? ctas_per_sm : want)) / (n_tiles * splits);
    if (gx < 1) gx = 1;
    if (gx > row_tiles) gx = row_tiles;
  }
  dim3 grid(gx, n_tiles, splits);
  if UD3D_DBG(p, 2048) {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, gather_gemm_tc_kernel<N_TILE>);
    int occ = -1;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gather_gemm_tc_kernel<N_TILE>, kThreadsTc, smem);
    int smem_sm = 0, regs_sm = 0;
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, 0);
    cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, 0);
    printf("launch_tc<%d>: grid %d x %d x %d, ctas/SM %d (now %d, err %d), smem %zu, sms %d; regs %d static smem %zu maxdyn %d carveout %d; SM smem %d regs %d\n",
           N_TILE, gx, n_tiles, splits, ctas_per_sm, occ, (int)e, smem, g_num_sms, fa.numRegs, fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes,
           fa.preferredShmemCarveout, smem_sm, regs_sm);
  }
  if (splits > 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kThreadsTc);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = (unsigned)splits;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    UD3D_CUDA(cudaLaunchKernelEx(&cfg, gather_gemm_tc_kernel<N_TILE>, p));
  } else {
    gather_gemm_tc_kernel<N_TILE><<<grid, kThreadsTc, smem, st>>>(p);
  }
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

static int check_args(const ud3d_gemm_args* a, const char* who) {
  UD3D_CHECK_ARG(a && a->in && a->out, "%s: NULL in/out", who);
  UD3D_CHECK_ARG(a->c_in > 0 && a->c_out > 0 && a->K > 0 && a->n_out >= 0, "%s: bad sizes", who);
  UD3D_CHECK_ARG(a->ld_in >= a->c_in && a->ld_out >= a->c_out, "%s: leading dimension smaller than channel count", who);
  UD3D_CHECK_ARG(a->table || a->K == 1, "%s: identity gather requires K == 1", who);
  UD3D_CHECK_ARG(!a->residual || a->ld_res >= a->c_out, "%s: bad ld_res", who);
  UD3D_CHECK_ARG((a->in_scale == nullptr) == (a->in_shift == nullptr), "%s: in_scale/in_shift must both be set", who);
  UD3D_CHECK_ARG(!a->table || a->K <= 32, "%s: a gather table supports at most 32 kernel offsets", who);
  UD3D_CHECK_ARG(!a->row_perm || a->table, "%s: row_perm requires a gather table", who);
  if (a->in_split) {
    UD3D_CHECK_ARG(!a->in_scale && !a->in_relu, "%s: in_split input is already activated (no in_scale / in_relu)", who);
    UD3D_CHECK_ARG(a->c_in % 32 == 0 && a->ld_in % 32 == 0 && ((uintptr_t)a->in & 15) == 0,
                   "%s: in_split needs c_in %% 32 == 0, ld_in %% 32 == 0, 16-byte aligned", who);
  }
  for (int i = 0; i < 2; ++i) {
    if (!a->out_act[i]) continue;
    UD3D_CHECK_ARG((a->act_scale[i] == nullptr) == (a->act_shift[i] == nullptr), "%s: act_scale / act_shift must both be set or both NULL", who);
    UD3D_CHECK_ARG(a->c_out % 32 == 0 && a->ld_act[i] % 4 == 0 && a->ld_act[i] >= a->c_out &&
                       ((uintptr_t)a->out_act[i] & 15) == 0,
                   "%s: out_act needs c_out %% 32 == 0 and a 16-byte aligned buffer", who);
  }
  return UD3D_OK;
}

}  // namespace ud3d

using namespace ud3d;

extern "C" {

size_t ud3d_gemm_packed_weight_bytes(int K, int c_in, int c_out) {
  if (K <= 0 || c_in <= 0 || c_out <= 0) return 0;
  int nts = pick_ntile(c_out);
  int n_tiles = cdiv(c_out, nts), n_chunks = cdiv(c_in, kChunk);
  return (size_t)n_tiles * K * n_chunks * nts * 128;
}

int ud3d_gemm_pack_weight(const float* w, int K, int c_in, int c_out, void* packed, void* stream) {
  UD3D_CHECK_ARG(w && packed && K > 0 && c_in > 0 && c_out > 0, "ud3d_gemm_pack_weight: bad argument");
  UD3D_CHECK_ARG(((uintptr_t)packed & 127) == 0, "ud3d_gemm_pack_weight: packed buffer must be 128-byte aligned");
  int nts = pick_ntile(c_out);
  int n_tiles = cdiv(c_out, nts), n_chunks = cdiv(c_in, kChunk);
  long long total = (long long)n_tiles * K * n_chunks * nts * 8;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  pack_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, K, c_in, c_out, nts, n_tiles, n_chunks, (uint4*)packed);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

int ud3d_gemm_pack_weight_ts(const float* w, int K, int c_in, int c_out, void* packed, void* stream) {
  UD3D_CHECK_ARG(w && packed && K > 0 && c_in > 0 && c_out > 0, "ud3d_gemm_pack_weight_ts: bad argument");
  UD3D_CHECK_ARG(((uintptr_t)packed & 127) == 0, "ud3d_gemm_pack_weight_ts: packed buffer must be 128-byte aligned");
  return launch_pack_weight_ts(w, K, c_in, c_out, pick_ntile(c_out), packed, (cudaStream_t)stream);
}

int ud3d_gemm_fwd(const ud3d_gemm_args* args, void* stream) {
  int rc = check_args(args, "ud3d_gemm_fwd");
  if (rc) return rc;
  UD3D_CHECK_ARG(args->w_packed && ((uintptr_t)args->w_packed & 127) == 0, "ud3d_gemm_fwd: w_packed NULL or misaligned");
  if (args->n_out == 0) return UD3D_OK;
  GemmParams p;
  p.a = *args;
  p.n_chunks = cdiv(args->c_in, kChunk);
  p.trace = g_trace;
  p.trace_block = g_trace_block;
  p.trace_iter = kPersistent ? 1 : 0;
  p.dbg = g_dbg;
  p.vec_ok = (args->ld_in % 4 == 0) && (((uintptr_t)args->in & 15) == 0) && (args->c_in % 8 == 0);
  p.out_vec_ok = (args->ld_out % 4 == 0) && (((uintptr_t)args->out & 15) == 0) &&
                 (!args->bias || ((uintptr_t)args->bias & 15) == 0) &&
                 (!args->act_scale[0] || (((uintptr_t)args->act_scale[0] | (uintptr_t)args->act_shift[0]) & 15) == 0) &&
                 (!args->act_scale[1] || (((uintptr_t)args->act_scale[1] | (uintptr_t)args->act_shift[1]) & 15) == 0) &&
                 (!args->residual || ((args->ld_res % 4 == 0) && (((uintptr_t)args->residual & 15) == 0)));
  int nts = pick_ntile(args->c_out);
  int n_tiles = cdiv(args->c_out, nts);
  cudaStream_t st = (cudaStream_t)stream;
  // split-K for launches that cannot fill the 148 SMs (deep U-Net levels: 5..91 row tiles x 100+ K-steps): the
  // (offset, chunk) sequence is cut across the gridDim.z CTAs of a (1, 1, z) thread-block cluster, which reduce their
  // partial tiles through distributed shared memory (see the kernel's epilogue)
  int splits = 1;
  {
    long long ctas = (long long)cdiv(args->n_out, kTileM) * n_tiles;
    int max_steps = args->K * p.n_chunks;
    if (ctas < 120 && max_steps >= 16) {
      splits = (int)(296 / ctas);
      if (splits > max_steps / 6) splits = max_steps / 6;
      if (splits > 8) splits = 8;      // portable cluster size (14..16-CTA clusters scheduled in two waves: slower)
      if (splits < 1) splits = 1;
    }
  }
  // operand-form inputs of launches that fill the GPU: the A operand goes global -> registers -> TMEM (gemm_ts.cu)
  // dense GEMMs on operand-form rows with 128-column weight tiles (the encoder's Linear layers): the persistent TMA kernel
  // The experimental TMEM-operand kernel (gemm_ts.cu) takes its input in the INTERLEAVED operand form (in_split == 2,
  // ops.operand_form_interleave); the product path does not use it: measured on B200 it wins with warm caches
  // (graph-replayed level-1 SubM3 32->32: 127 vs 182 us, 64->64: 61 vs 92 us) but not inside the real step, where every
  // launch starts on cold inputs (ncu launch list: 112 vs 103 us, 65 vs 53 us) -- see DESIGN.md section 4.
  if (args->in_split == 2) {
    UD3D_CHECK_ARG(args->w_packed_ts && splits == 1 && nts <= 160 && (!args->table || args->tile_mask),
                   "ud3d_gemm_fwd: the interleaved operand form (in_split == 2) is the input of the TMEM-operand kernel only: it needs "
                   "w_packed_ts, a tile mask with a table, c_out <= 160 per tile and a launch that fills the GPU");
    UD3D_CHECK_ARG(((uintptr_t)args->w_packed_ts & 127) == 0, "ud3d_gemm_fwd: w_packed_ts misaligned");
    UD3D_CHECK_ARG(((uintptr_t)args->in & 31) == 0 && args->ld_in % 8 == 0, "ud3d_gemm_fwd: operand-form input must be 32-byte aligned");
    DeviceCtx* ctx = device_ctx();
    if (!ctx) return UD3D_ECUDA;
    return launch_gemm_ts(p, nts, ctx_sm_count(ctx), st);
  }
  int rc2;
  switch (nts) {
    case 32: rc2 = launch_tc<32>(p, n_tiles, splits, st); break;
    case 64: rc2 = launch_tc<64>(p, n_tiles, splits, st); break;
    case 96: rc2 = launch_tc<96>(p, n_tiles, splits, st); break;
    case 128: rc2 = launch_tc<128>(p, n_tiles, splits, st); break;
    case 160: rc2 = launch_tc<160>(p, n_tiles, splits, st); break;
    default: rc2 = launch_tc<256>(p, n_tiles, splits, st); break;
  }
  return rc2;
}

int ud3d_act_split(const float* raw, int ld_raw, int n, int c, const float* scale, const float* shift, int relu,
                   float* out_split, int ld_out, void* stream) {
  UD3D_CHECK_ARG(raw && out_split && n >= 0 && c > 0, "ud3d_act_split: bad argument");
  UD3D_CHECK_ARG(ld_out % 4 == 0 && ld_raw >= c && ld_out >= cdiv(c, 32) * 32 && ((uintptr_t)out_split & 15) == 0,
                 "ud3d_act_split: out_split needs ld_out >= roundup(c, 32) and 16-byte aligned rows");
  UD3D_CHECK_ARG((scale == nullptr) == (shift == nullptr), "ud3d_act_split: scale/shift must both be set");
  if (n == 0) return UD3D_OK;
  return launch_act_split(raw, ld_raw, n, c, scale, shift, relu, out_split, ld_out, (cudaStream_t)stream);
}

/* debug only (not declared in the public header; effective only in -DUD3D_DEBUG_HOOKS builds): record clock64
 * timestamps of one CTA of subsequent ud3d_gemm_fwd launches into `buf` (device, >= 1024 int64), or stop with buf = NULL */
int ud3d_debug_set_trace(long long* buf, int block) {
#ifndef UD3D_DEBUG_HOOKS
  UD3D_CHECK_ARG(buf == nullptr, "ud3d_debug_set_trace: this library was built without -DUD3D_DEBUG_HOOKS");
#endif
  g_trace = buf;
  g_trace_block = block;
  return UD3D_OK;
}

int ud3d_debug_set_flags(int flags) {
#ifndef UD3D_DEBUG_HOOKS
  UD3D_CHECK_ARG(flags == 0, "ud3d_debug_set_flags: this library was built without -DUD3D_DEBUG_HOOKS");
#endif
  g_dbg = flags;
  return UD3D_OK;
}

int ud3d_gemm_fwd_simt(const ud3d_gemm_args* args, const float* w, void* stream) {
  int rc = check_args(args, "ud3d_gemm_fwd_simt");
  if (rc) return rc;
  UD3D_CHECK_ARG(w, "ud3d_gemm_fwd_simt: NULL weight");
  if (args->n_out == 0) return UD3D_OK;
  dim3 grid(cdiv(args->n_out, 8), cdiv(args->c_out, 32)), block(32, 8);
  gather_gemm_simt_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(*args, w);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}

}  // extern "C"
