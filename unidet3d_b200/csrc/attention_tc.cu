// Varlen multi-head self-attention on the 5th-gen tensor cores (head_dim 32, operand-form q|k|v), sm_100a.
// Reference: nn.MultiheadAttention inside SelfAttentionLayer.forward, unidet3d/encoder.py:24-41 (per scene, no mask).
//
// CTA = 128 queries of one (scene, head); K/V tiles of 64 keys; 2 CTAs per SM (256 TMEM columns each):
//   S[128 x 64]  = Q K^T  : A = Q tile, B = K tile, both K-major SW128 rows of 128 B (bf16 hi | lo); 6 MMAs (hi.hi,
//                           lo.hi, hi.lo: logits need the precision, an error in S is an exponent error in P), N = 64;
//                           DOUBLE-BUFFERED: S of tile j+2 is issued right after P V of tile j, so the softmax warps
//                           find the next S ready and run back to back (their phase is the critical path)
//   softmax               : 4 warps, ONE THREAD PER QUERY ROW (TMEM lane): tcgen05.ld S (each chunk's load in flight while
//                           the previous chunk is exponentiated), row max without any cross-thread exchange,
//                           P = 2^(S c - m) split into bf16 hi + lo and written with tcgen05.st straight into TMEM, over
//                           the S columns just consumed, as the A operand of the second GEMM (no shared-memory round trip)
//   O'[128 x 64] += P V'  : A = P_hi, P_lo from TMEM, B = V tile used MN-major (row = key, 128 B = [V_hi(32) | V_lo(32)] =
//                           N 64): all four hi/lo products; the accumulator STAYS in TMEM over all KV tiles,
//                           O = O'[:, :32] + O'[:, 32:].  (The precision ladder -- tools/attn_precision_ladder.py -- shows
//                           that with a peaked softmax NO term of either GEMM can be dropped: P as a single bf16 term
//                           costs 1.4e-2 on the encoder's final logits, a single-term Q K^T 0.16.)
//   L[128 x 16]  += P 1   : the row sums come from the tensor core too (P_hi, P_lo times a tile of ones): numerator and
//                           denominator of softmax(S) V see exactly the same P, at no ALU cost
//   rescaling             : lazy and EXACT -- the reference maximum of a row is an integer (log2 domain) and only moves when
//                           a tile's maximum exceeds it by more than 8; then that row's O' and sum are rescaled in TMEM
//                           by the row's own thread, by an exact power of two.  In practice: the first tile or two.
//   roles                 : warps 0-3 softmax / output, warp 4 K/V loader, warp 5 MMA issue; mbarrier hand-offs only.
//   loads                 : TMA tensor-map copies (cp.async.bulk.tensor.2d, 128B hardware swizzle): q|k|v is described
//                           as a 2-D byte tensor [tokens, 3 d 4 B], a Q / K / V tile = a 128 B x 64-row box -- ONE
//                           instruction per tile instead of 1024 16-byte cp.async (the loader's address arithmetic was
//                           the largest share of the instruction stream of its scheduler partition)
// Bounds per 128 x 64 tile: MUFU 8192 ex2 at 16 per cycle and SM = 512 cycles; TMEM read of S (32 KB at ~64 B per cycle)
// = 512 cycles (S is read exactly once); MMAs ~320 cycles.
#include "common.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>

namespace ud3d {

constexpr int kAQ = 128, kAK = 64, kAStages = 4;
constexpr int kASoftmaxWarps = 4;
constexpr int kAThreads = 32 * (kASoftmaxWarps + 2);
// TMEM columns: S buffers at 0..63 and 64..127; P (32 columns of bf16 pairs) ALIASES the first half of the S buffer it was
// computed from, written by the row's own thread after it has read that buffer (the buffer's next S = Q K^T is issued
// after this tile's P V on the in-order tensor pipe, so it cannot overwrite P early); O' at 128..191; the row sums L at
// 192..207 (P times a tile of ones: the tensor core sums exactly the rounded P it multiplies into O').
constexpr uint32_t kAColS = 0, kAColO = 128, kAColL = 192, kACols = 256;
constexpr float kARescaleThr = 8.f;          // log2 domain

// UMMA descriptor for an MN-major operand tile with 128B swizzle: rows (K index) of 128 bytes, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128_mn_a(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_m128_bmn_a(uint32_t N) {   // B operand MN-major
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((N >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ float ex2_approx_a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ uint32_t tmem_ld_32x32_x1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}
__device__ __forceinline__ void tmem_st_32x32_x1(uint32_t taddr, uint32_t r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(r) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// 2-D tiled TMA load: box (128 bytes x 64 rows) at byte column x, row y of the q|k|v tensor -> 128B-swizzled smem tile
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar))
               : "memory");
}
// 32 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]),
      "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait_a() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(kAThreads, 2) attention_tc_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                   const int32_t* __restrict__ cu, int num_heads,
                                                                   uint8_t* __restrict__ out) {
  const int b = blockIdx.z, h = blockIdx.y;
  const int t0 = cu[b];
  const int T = cu[b + 1] - t0;
  const int q0 = blockIdx.x * kAQ;
  if (q0 >= T) return;
  const int d_model = num_heads * 32;
  const size_t ldo = (size_t)d_model * 4;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                          // 128 x 128 B
  uint8_t* sK = sQ + 16384;                    // [kAStages] 64 x 128 B
  uint8_t* sV = sK + kAStages * 8192;          // [kAStages] 64 x 128 B
  uint8_t* sOnes = sV + kAStages * 8192;       // 1 KB of bf16 1.0: the B operand of the row-sum MMAs
  uint64_t* bars = (uint64_t*)(sOnes + 1024);
  uint64_t* q_full = bars;                     // loader: expect-tx + the TMA copies' bytes
  uint64_t* kv_full = bars + 1;                // [kAStages] loader: expect-tx + bytes
  uint64_t* kv_empty = kv_full + kAStages;     // [kAStages] tcgen05.commit
  uint64_t* s_full = kv_empty + kAStages;      // [2] tcgen05.commit (per S buffer)
  uint64_t* p_full = s_full + 2;               // [2] 4 softmax warps (per S buffer: a waiter is never two phases behind)
  uint64_t* o_done = p_full + 2;               // [2] tcgen05.commit (P V of the tiles with this parity)
  uint32_t* tmem_slot = (uint32_t*)(o_done + 2);

  if (tid == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kAStages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], kASoftmaxWarps); mbar_init(&o_done[i], 1); }
    fence_mbar_init();
  }
  for (int i = tid; i < 256; i += kAThreads) ((uint32_t*)sOnes)[i] = 0x3F803F80u;
  fence_proxy_async_smem();
  if (warp == 0) {
    tmem_alloc(tmem_slot, kACols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tiles = (T + kAK - 1) / kAK;

  if (warp < kASoftmaxWarps) {
    // =========================================================== softmax / output: thread = query row = TMEM lane
    const int row = warp * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float qscale = 1.44269504088896340736f * 0.17677669529663688110f;   // log2(e) / sqrt(32)
    float m_ref = -INFINITY;
    // P chunk (32 keys) = 2^(S c - m_ref) split into bf16 hi | lo: 16 + 16 packed registers
    auto exp_split = [&](const uint32_t (&r)[32], int nvc, uint32_t (&pk)[32]) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float p0, p1;
        p0 = ex2_approx_a(fmaf(__uint_as_float(r[i]), qscale, -m_ref));
        p1 = ex2_approx_a(fmaf(__uint_as_float(r[i + 1]), qscale, -m_ref));
        if (nvc < 32) {
          if (i >= nvc) p0 = 0.f;
          if (i + 1 >= nvc) p1 = 0.f;
        }
        split_bf16x2(p0, p1, pk[i >> 1], pk[16 + (i >> 1)]);
      }
    };
    auto chunk_max = [&](const uint32_t (&r)[32], int nvc) {
      float m = -INFINITY;
      if (nvc >= 32) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) m = fmax3(m, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < nvc) m = fmaxf(m, __uint_as_float(r[i]));
      }
      return m;
    };
    uint32_t ra[32], rb[32], pk[32];
    mbar_wait(&s_full[0], 0u);
    tc_fence_after_sync();
    tmem_ld_32x32(t_lane + kAColS, ra);                 // chunk a of tile 0
    for (int j = 0; j < n_tiles; ++j) {
      const int buf = j & 1;
      const int nv = T - j * kAK;                       // valid keys of this tile (>= 1)
      const uint32_t tSb = t_lane + kAColS + (uint32_t)(buf * 64);
      // S is read from TMEM exactly ONCE (TMEM reads run at ~64 B per cycle and SM: a tile's 32 KB cost as much as its
      // 8192 exponentials), and every load is in flight while the other chunk is exponentiated.
      tmem_ld_wait();                                   // chunk a (requested during the previous tile)
      tmem_ld_32x32(tSb + 32u, rb);                     // chunk b
      const float cmax_a = chunk_max(ra, nv);
      if (j > 0) exp_split(ra, nv, pk);                 // optimistic: with the current reference maximum
      tmem_ld_wait();
      const float hmax = fmaxf(cmax_a, chunk_max(rb, nv - 32)) * qscale;
      // ---- lazy, exact rescale: the reference maximum is an integer and only moves when this tile exceeds it by > 8
      const bool need = hmax > m_ref + kARescaleThr;    // (always true for tile 0: m_ref = -inf)
      if (__any_sync(0xffffffffu, need)) {
        float factor = 1.f;
        if (need) {
          const float m_new = ceilf(hmax);
          if (j > 0) {                                  // 2^(m_ref - m_new), built exactly (both are integers)
            const float d = m_ref - m_new;
            factor = d < -126.f ? 0.f : __int_as_float((127 + (int)d) << 23);
          }
          m_ref = m_new;
        }
        if (j > 0) {
          // O' and L are complete up to tile j-1 once its MMAs have completed (those of tile j wait for p_full below)
          mbar_wait(&o_done[(j - 1) & 1], (uint32_t)((j - 1) >> 1) & 1u);
          tc_fence_after_sync();
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(t_lane + kAColO + (uint32_t)(c * 32), r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * factor);
            tmem_st_32x32(t_lane + kAColO + (uint32_t)(c * 32), r);
          }
          const uint32_t lv = tmem_ld_32x32_x1(t_lane + kAColL);
          tmem_ld_wait();
          tmem_st_32x32_x1(t_lane + kAColL, __float_as_uint(__uint_as_float(lv) * factor));
        }
        exp_split(ra, nv, pk);                          // chunk a again, with the new reference (ra still holds raw S)
      }
      // P of chunk a over the S columns it came from: hi at +0..15, lo at +16..31
      tmem_st_32x32_x16(tSb, pk);
      tmem_st_32x32_x16(tSb + 16u, pk + 16);
      if (j + 1 < n_tiles) {
        mbar_wait(&s_full[buf ^ 1], (uint32_t)((j + 1) >> 1) & 1u);
        tc_fence_after_sync();
        tmem_ld_32x32(t_lane + kAColS + (uint32_t)((buf ^ 1) * 64), ra);      // chunk a of the next tile
      }
      exp_split(rb, nv - 32, pk);
      tmem_st_32x32_x16(tSb + 32u, pk);
      tmem_st_32x32_x16(tSb + 48u, pk + 16);
      tmem_st_wait_a();
      tc_fence_before_sync();           // P (and a rescaled O' / L) are written, S has been read
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[buf]);
    }
    // ---- output: O = (O'[:, :32] + O'[:, 32:]) / L, operand form (head h == 32-channel chunk h)
    mbar_wait(&o_done[(n_tiles - 1) & 1], (uint32_t)((n_tiles - 1) >> 1) & 1u);
    tc_fence_after_sync();
    uint32_t a[32], c2[32];
    tmem_ld_32x32(t_lane + kAColO, a);
    tmem_ld_32x32(t_lane + kAColO + 32u, c2);
    const uint32_t lraw = tmem_ld_32x32_x1(t_lane + kAColL);
    tmem_ld_wait();
    const int grow = q0 + row;
    if (grow < T) {
      const float inv = 1.f / __uint_as_float(lraw);
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2)
        split_bf16x2((__uint_as_float(a[i]) + __uint_as_float(c2[i])) * inv, (__uint_as_float(a[i + 1]) + __uint_as_float(c2[i + 1])) * inv,
                     hi[i >> 1], lo[i >> 1]);
      uint4* dst = (uint4*)(out + (size_t)(t0 + grow) * ldo + (size_t)h * 128);
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        dst[opf_mem_piece(c4)] = make_uint4(hi[4 * c4], hi[4 * c4 + 1], hi[4 * c4 + 2], hi[4 * c4 + 3]);
        dst[opf_mem_piece(4 + c4)] = make_uint4(lo[4 * c4], lo[4 * c4 + 1], lo[4 * c4 + 2], lo[4 * c4 + 3]);
      }
    }
  } else if (warp == kASoftmaxWarps) {
    // =========================================================== loader: Q once, then K/V tiles through the ring (TMA)
    // (rows past the scene's last token belong to the next scene or are out of bounds (zero-filled): their scores are
    //  masked in the softmax and their P is exactly 0, so whatever finite values they hold never reach the output)
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 16384);
      tma_load_2d(sQ, &tmap, h * 128, t0 + q0, q_full);
      tma_load_2d(sQ + 8192, &tmap, h * 128, t0 + q0 + 64, q_full);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % kAStages;
        if (j >= kAStages) mbar_wait(&kv_empty[st], (uint32_t)(j / kAStages - 1) & 1u);
        mbar_arrive_expect_tx(&kv_full[st], 16384);
        tma_load_2d(sK + st * 8192, &tmap, (num_heads + h) * 128, t0 + j * kAK, &kv_full[st]);
        tma_load_2d(sV + st * 8192, &tmap, (2 * num_heads + h) * 128, t0 + j * kAK, &kv_full[st]);
      }
    }
    __syncwarp();
  } else {
    // =========================================================== MMA issue (one elected lane; the warp stays converged)
    constexpr uint32_t IDESC_S = umma_idesc_bf16_m128(kAK);
    // P V: B (= V) MN-major, N = 64; row sums: B (= ones) K-major, N = 16.  (P as fp16 with a bf16 V -- three more mantissa
    // bits for the softmax weights -- is not possible: kind::f16 with different A / B formats is an illegal instruction.)
    constexpr uint32_t IDESC_O = umma_idesc_bf16_m128_bmn_a(64);
    constexpr uint32_t IDESC_L = umma_idesc_bf16_m128(16);
    const uint64_t qd = umma_desc_sw128(smem_u32(sQ));
    // ones tile: no swizzle, K-major, core matrices 128 B apart along K and 256 B along N -- every element is 1.0, so any
    // in-range addressing reads ones
    const uint64_t od = (uint64_t)((smem_u32(sOnes) & 0x3FFFFu) >> 4) | (8ull << 16) | (16ull << 32) | (1ull << 46);
    const uint32_t tS = tmem_base + kAColS, tO = tmem_base + kAColO, tL = tmem_base + kAColL;
    auto issue_s = [&](int j) {
      const int st = j % kAStages;
      mbar_wait(&kv_full[st], (uint32_t)(j / kAStages) & 1u);
      tc_fence_after_sync();
      if (elect_one_sync()) {
        const uint64_t kd = umma_desc_sw128(smem_u32(sK + st * 8192));
        const uint32_t tSb = tS + (uint32_t)((j & 1) * 64);
        umma_bf16(tSb, qd + 0, kd + 0, IDESC_S, 0);
        umma_bf16(tSb, qd + 2, kd + 2, IDESC_S, 1);
        umma_bf16(tSb, qd + 4, kd + 0, IDESC_S, 1);
        umma_bf16(tSb, qd + 6, kd + 2, IDESC_S, 1);
        umma_bf16(tSb, qd + 0, kd + 4, IDESC_S, 1);
        umma_bf16(tSb, qd + 2, kd + 6, IDESC_S, 1);
        umma_commit(&s_full[j & 1]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0u);
    issue_s(0);
    if (n_tiles > 1) issue_s(1);
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j % kAStages;
      mbar_wait(&p_full[j & 1], (uint32_t)(j >> 1) & 1u);     // P(j) written, S(j) consumed, O' / L rescaled if needed
      tc_fence_after_sync();
      if (elect_one_sync()) {
        const uint64_t vd = umma_desc_sw128_mn_a(smem_u32(sV + st * 8192));
        const uint32_t tP = tS + (uint32_t)((j & 1) * 64);
        // K-step i covers keys 16 i .. 16 i + 15 = 8 columns of P_hi and 8 of P_lo (per 32-key chunk: hi at +0..15, lo at
        // +16..31); V rows advance by 16 x 128 B
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t ph = tP + (uint32_t)((i >> 1) * 32 + (i & 1) * 8), pl = ph + 16u;
          const uint64_t vdi = vd + (uint64_t)(i * 128);
          umma_bf16_ts(tO, ph, vdi, IDESC_O, (j > 0 || i > 0) ? 1u : 0u);
          umma_bf16_ts(tO, pl, vdi, IDESC_O, 1u);
          umma_bf16_ts(tL, ph, od, IDESC_L, (j > 0 || i > 0) ? 1u : 0u);
          umma_bf16_ts(tL, pl, od, IDESC_L, 1u);
        }
        umma_commit(&o_done[j & 1]);
        umma_commit(&kv_empty[st]);
      }
      __syncwarp();
      if (j + 2 < n_tiles) issue_s(j + 2);              // into the S buffer P V (j) has just finished with (in-order pipe)
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 0) tmem_dealloc(tmem_base, kACols);
}

}  // namespace ud3d

using namespace ud3d;

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
static PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (PFN_cuTensorMapEncodeTiled_v12000)p;
  }();
  return fn;
}

extern "C" int ud3d_attention_fwd_tc(const float* qkv_split, const int32_t* cu_seqlens, int B, int max_T, int total_T,
                                     int num_heads, float* out_split, void* stream) {
  UD3D_CHECK_ARG(qkv_split && cu_seqlens && out_split, "ud3d_attention_fwd_tc: NULL argument");
  UD3D_CHECK_ARG(B > 0 && num_heads > 0 && max_T >= 0 && total_T >= max_T, "ud3d_attention_fwd_tc: bad sizes");
  UD3D_CHECK_ARG((((uintptr_t)qkv_split | (uintptr_t)out_split) & 15) == 0, "ud3d_attention_fwd_tc: pointers must be 16-byte aligned");
  if (max_T == 0) return UD3D_OK;
  auto encode = tensor_map_encoder();
  if (!encode) {
    set_error("ud3d_attention_fwd_tc: cuTensorMapEncodeTiled is not available from this driver");
    return UD3D_ECUDA;
  }
  // q|k|v as a 2-D byte tensor [total_T rows, 3 * d_model * 4 bytes]; box = 128 bytes x 64 rows, 128B swizzle, zero OOB fill
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)3 * num_heads * 128, (cuuint64_t)total_T};
  const cuuint64_t gstride[1] = {(cuuint64_t)3 * num_heads * 128};
  const cuuint32_t box[2] = {128, 64};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)qkv_split, gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("ud3d_attention_fwd_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return UD3D_ECUDA;
  }
  const size_t smem = 1024 + 16384 + 2 * kAStages * 8192 + 1024 + 256;
  DeviceCtx* ctx = device_ctx();
  if (!ctx) return UD3D_ECUDA;
  {
    CtxGuard guard(ctx);
    if (ctx_needs_config(ctx, (const void*)attention_tc_kernel, smem)) {
    UD3D_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    UD3D_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    }
  }
  dim3 grid(cdiv(max_T, kAQ), num_heads, B);
  attention_tc_kernel<<<grid, kAThreads, smem, (cudaStream_t)stream>>>(tmap, cu_seqlens, num_heads, (uint8_t*)out_split);
  UD3D_LAUNCH_CHECK();
  return UD3D_OK;
}
