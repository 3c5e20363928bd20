// Shared helpers for the unidet3d_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/unidet3d_b200.h"

namespace ud3d {

// ---------------------------------------------------------------- errors / bookkeeping
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define UD3D_CHECK_ARG(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      ::ud3d::set_error(__VA_ARGS__);             \
      return UD3D_EINVAL;                         \
    }                                             \
  } while (0)

#define UD3D_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      ::ud3d::set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return UD3D_ECUDA;                                                                  \
    }                                                                                     \
  } while (0)

#define UD3D_LAUNCH_CHECK()                                                               \
  do {                                                                                    \
    ::ud3d::count_launch();                                                               \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess) {                                                             \
      ::ud3d::set_error("%s:%d kernel launch: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return UD3D_ECUDA;                                                                  \
    }                                                                                     \
  } while (0)

// ---------------------------------------------------------------- per-device context ("no global state
// except an opaque ud3d_ctx* per device").  Everything the library caches belongs to one device: the SM count and, per
// kernel, the largest dynamic shared-memory size / carve-out its function attributes were set to.  Contexts are created
// lazily for the calling thread's current device and guarded by a mutex (thread-safe; launches themselves are not
// serialised).  Defined in grid.cu.
struct DeviceCtx;
DeviceCtx* device_ctx(int* device_out = nullptr);          // current device; nullptr + set_error on failure
int ctx_sm_count(const DeviceCtx* c);
// Returns true when kernel `key` has to be (re)configured on this device for `smem` bytes / `carve` percent, i.e. when
// the call asks for more than what was configured so far; the caller then issues cudaFuncSetAttribute.  The caller must
// hold the context's lock (CtxGuard) from this call until the attributes ARE set: another host thread that finds the
// kernel "configured" launches at once (a launch asking for more dynamic shared memory than the attribute allows fails
// with "invalid argument").
bool ctx_needs_config(DeviceCtx* c, const void* key, size_t smem, int carve = -1);
void ctx_lock(DeviceCtx* c);
void ctx_unlock(DeviceCtx* c);
struct CtxGuard {
  DeviceCtx* c;
  explicit CtxGuard(DeviceCtx* ctx) : c(ctx) { ctx_lock(c); }
  ~CtxGuard() { ctx_unlock(c); }
  CtxGuard(const CtxGuard&) = delete;
  CtxGuard& operator=(const CtxGuard&) = delete;
};

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------- occupancy grid layout
struct GridDims {
  int B, X, Y, Z, Zw;          // Zw = words per z-column
  long long nwords;
};
static inline GridDims make_grid_dims(const int32_t d[4]) {
  GridDims g;
  g.B = d[0]; g.X = d[1]; g.Y = d[2]; g.Z = d[3];
  g.Zw = (g.Z + 31) / 32;
  g.nwords = (long long)g.B * g.X * g.Y * g.Zw;
  return g;
}
// workspace carve-up (all 256B aligned): words | prefix | block sums | count
struct GridView {
  uint32_t* words;
  uint32_t* prefix;
  uint32_t* bsum;
  int nblocks;
};
constexpr int kScanBlockWords = 4096;   // words scanned per CTA
static inline size_t grid_ws_bytes(const GridDims& g) {
  size_t nb = (size_t)((g.nwords + kScanBlockWords - 1) / kScanBlockWords);
  return align_up((size_t)g.nwords * 4, 256) * 2 + align_up((nb + 1) * 4, 256);
}
static inline GridView grid_view(const GridDims& g, const void* ws) {
  GridView v;
  char* p = (char*)ws;
  v.words = (uint32_t*)p;
  p += align_up((size_t)g.nwords * 4, 256);
  v.prefix = (uint32_t*)p;
  p += align_up((size_t)g.nwords * 4, 256);
  v.bsum = (uint32_t*)p;
  v.nblocks = (int)((g.nwords + kScanBlockWords - 1) / kScanBlockWords);
  return v;
}

#ifdef __CUDACC__
__device__ __forceinline__ long long grid_word_index(const GridDims& g, int b, int x, int y, int z) {
  return (((long long)b * g.X + x) * g.Y + y) * g.Zw + (z >> 5);
}
// canonical rank of cell (b,x,y,z) or -1
__device__ __forceinline__ int grid_rank(const GridDims& g, const uint32_t* __restrict__ words,
                                         const uint32_t* __restrict__ prefix, int b, int x, int y, int z) {
  if ((unsigned)b >= (unsigned)g.B || (unsigned)x >= (unsigned)g.X || (unsigned)y >= (unsigned)g.Y ||
      (unsigned)z >= (unsigned)g.Z)
    return -1;
  long long w = grid_word_index(g, b, x, y, z);
  uint32_t bits = __ldg(words + w);
  uint32_t bit = 1u << (z & 31);
  if (!(bits & bit)) return -1;
  return (int)(__ldg(prefix + w) + __popc(bits & (bit - 1)));
}

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------- PTX wrappers (sm_100a)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)     // suspend-time hint: sleep in hardware instead of spinning
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps instead of hanging the GPU box (slow path kept out of line)
static __device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) {
      printf("ud3d: mbarrier wait timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}
// try_wait without a suspend-time hint (the default, short, system-dependent time limit), in a bounded loop
__device__ __forceinline__ void mbar_wait_nohint(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (++spins > (1u << 24)) __trap();
  }
}
// spin on test_wait (no hardware suspend): lowest wake-up latency, for single-thread roles
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (++spins > (1u << 26)) __trap();
  }
}
// generic-proxy smem writes -> visible to the async proxy (tensor core / bulk copy engine)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP), completion on an mbarrier
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// 16-byte async copy global -> shared (LDGSTS, L2-only caching); src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async_16_zfill(uint32_t smem_dst, const void* gmem_src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_16_zfill_ca(uint32_t smem_dst, const void* gmem_src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void st_shared_zero16(uint32_t smem_dst) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(smem_dst), "r"(0) : "memory");
}
// the mbarrier receives one arrival (counted in its expected count) when all prior cp.async of this thread have landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate, single-CTA
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// warp-converged variants: one elected lane issues (keeps the operands in uniform registers)
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void umma_bf16_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate) {
  if (elect_one_sync()) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  if (elect_one_sync()) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
  }
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i = TMEM lane base+i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// ---- thread-block clusters: barrier over all CTAs of the cluster, distributed shared memory reads
__device__ __forceinline__ void cluster_arrive_release() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of this CTA's shared-memory location `smem_addr` in the CTA of rank `rank` of the cluster
__device__ __forceinline__ uint32_t cluster_map_shared(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 ld_shared_cluster_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// UMMA shared-memory descriptor: K-major operand tile, 128B swizzle, rows of 128 bytes,
// 8-row atoms 1024 B apart (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30) (=1, unused
// for swizzled K-major), SBO>>4 [32,46) = 64, version [46,48) = 1, layout_type [61,64) = 2 SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// same, 64B swizzle: rows of 64 bytes, 8-row atoms 512 B apart (layout_type 4 = SWIZZLE_64B)
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
// UMMA instruction descriptor: kind::f16, A=B=BF16, D=F32, both K-major, M=128, N
__host__ __device__ constexpr uint32_t umma_idesc_bf16_m128(uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((128u >> 4) << 24);
}

// ---- operand form of a feature map: per row, per 32-channel chunk, 128 bytes = 8 pieces of
// 16 bytes: pieces 0..3 = bf16 hi of channels 0..31, pieces 4..7 = bf16 lo -- byte for byte the 128-byte row of the
// K-major shared-memory tile, so that lane j of a gather copies piece j to chunk j (^ swizzle).  (An interleaved layout
// -- hi / lo of 8 channels next to each other, one 32-byte load = one thread's share of a tcgen05.st into a TMEM A
// operand, the input form of the experimental kernel in gemm_ts.cu -- was measured 35 % slower for the cp.async gather:
// any non-XOR permutation between lane order and address order inside the 128-byte row costs LDGSTS efficiency.)
// Every producer / consumer addresses the layout through these helpers.
__host__ __device__ __forceinline__ constexpr int opf_mem_piece(int L) { return L; }        // logical piece -> memory piece
// byte offset of the bf16 hi of channel ch (0..31) inside a 128-byte row-chunk; its lo part is kOpfLo bytes further
__host__ __device__ __forceinline__ constexpr int opf_hi_off(int ch) { return ch << 1; }
constexpr int kOpfLo = 64;

// fp32 -> bf16 hi + bf16 lo (x ~= hi + lo, |err| <~ 2^-17 |x|)
// Packed conversions (one F2FP.BF16.PACK_AB per pair, FMA-pipe) instead of scalar __float2bfloat16_rn (F2F, quarter-rate
// XU pipe: it co-limited the attention kernel, 96 F2F next to 96 HMMA per K/V tile); same round-to-nearest-even results.
__device__ __forceinline__ uint32_t cvt_bf16x2_rn(float lo16, float hi16) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi16), "f"(lo16));
  return d;
}
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = cvt_bf16x2_rn(a, b);
  const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xffff0000u);
  lo = cvt_bf16x2_rn(a - ah, b - bh);
}
#endif  // __CUDACC__

}  // namespace ud3d
