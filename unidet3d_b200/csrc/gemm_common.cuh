// Pieces shared by the two gather-GEMM kernels (gemm.cu: operands through shared memory; gemm_ts.cu: A operand gathered
// through registers into TMEM): launch parameters, activation, and the warp-collective epilogue.
#pragma once
#include "common.cuh"

namespace ud3d {

constexpr int kTileM = UD3D_TILE_M;   // 128
constexpr int kChunk = 32;            // input channels per K-step

struct GemmParams {
  ud3d_gemm_args a;
  int n_chunks;
  int vec_ok;      // 16-byte vector gather allowed
  int out_vec_ok;  // 16-byte vector epilogue allowed
  long long* trace;   // debug: clock64 timestamps of one CTA (ud3d_debug_set_trace), else nullptr
  int trace_block;
  int trace_iter;     // which of the CTA's tiles is traced
  int dbg;            // debug: feature-disable bits for timing breakdowns (ud3d_debug_set_flags), normally 0
};

// Timing-breakdown hooks (ud3d_debug_set_flags / ud3d_debug_set_trace, used by tools/*_trace.py and tools/*_probe.py)
// exist only in builds with -DUD3D_DEBUG_HOOKS (UD3D_NVCC_EXTRA=-DUD3D_DEBUG_HOOKS python -m unidet3d_b200.build --force);
// in the release library the conditions below are compile-time false and the branches disappear from the kernels.
#ifdef UD3D_DEBUG_HOOKS
#define UD3D_DBG(p, bits) (((p).dbg & (bits)) != 0)
#define UD3D_TRACE_BUF(p) ((p).trace)
#else
#define UD3D_DBG(p, bits) (false)
#define UD3D_TRACE_BUF(p) ((long long*)nullptr)
#endif

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  // (an Abramowitz-Stegun 7.1.26 erf -- rcp + ex2 + 5 FMA -- measured slower than libdevice's erff here)
  if (act == 2) return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
  return v;
}

// ---- warp-collective row-segment transposition through shared memory.
// In the epilogue every lane owns one output row and produces 128-byte row segments (32 fp32 columns, or one operand-form
// row-chunk).  Stored directly, one st.global.v4 instruction touches 32 different rows (32 partial-line writes of 16 B,
// 8 instructions per segment).  Staged through 4 KB of shared memory per warp (16-byte pieces XOR-swizzled with the
// row: conflict-free both ways), lanes 8r..8r+7 write the 8 pieces of one row: every instruction stores 4 complete
// 128-byte lines.  `dst` / `src` == nullptr: this lane's row is skipped.  All 32 lanes must call.
__device__ __forceinline__ void st_global_v4_na(void* p, const uint4& v) {     // streaming store: do not allocate in L1
  asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void warp_store_rows(uint32_t stage, const uint4 (&pc)[8], uint8_t* dst, int lane, bool na = false) {
  const uint32_t mine = stage + (uint32_t)lane * 128u;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(mine + (uint32_t)((j ^ (lane & 7)) << 4)), "r"(pc[j].x),
                 "r"(pc[j].y), "r"(pc[j].z), "r"(pc[j].w)
                 : "memory");
  __syncwarp();
  const int jj = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = (lane >> 3) + 4 * i;
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(stage + (uint32_t)(rr * 128) + (uint32_t)((jj ^ (rr & 7)) << 4))
                 : "memory");
    uint8_t* ptr = (uint8_t*)__shfl_sync(0xffffffffu, (unsigned long long)dst, rr);
    if (ptr) {
      if (na) st_global_v4_na(ptr + jj * 16, v);
      else *(uint4*)(ptr + jj * 16) = v;
    }
  }
  __syncwarp();
}
__device__ __forceinline__ void warp_load_rows(uint32_t stage, const uint8_t* src, uint4 (&pc)[8], int lane) {
  const int jj = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = (lane >> 3) + 4 * i;
    const uint8_t* ptr = (const uint8_t*)__shfl_sync(0xffffffffu, (unsigned long long)src, rr);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (ptr) v = __ldg((const uint4*)(ptr + jj * 16));
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + (uint32_t)(rr * 128) + (uint32_t)((jj ^ (rr & 7)) << 4)),
                 "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
  }
  __syncwarp();
  const uint32_t mine = stage + (uint32_t)lane * 128u;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(pc[j].x), "=r"(pc[j].y), "=r"(pc[j].z), "=r"(pc[j].w)
                 : "r"(mine + (uint32_t)((j ^ (lane & 7)) << 4))
                 : "memory");
  __syncwarp();
}

// One thread's 32 consecutive output columns [col0, col0+32) of row `grow`: bias / activation / residual, fp32 store
// and/or operand-form stores.  r = raw fp32 accumulator bits.  WARP-COLLECTIVE: all 32 lanes call it (`valid` = this
// lane has a row to write); `stage` = 4 KB of shared memory owned by the warp.
__device__ __forceinline__ void epilogue_store_chunk(const GemmParams& p, const uint32_t (&r)[32], int grow, int col0, bool valid,
                                                     uint32_t stage, int lane) {
  const ud3d_gemm_args& a = p.a;
  valid = valid && col0 < a.c_out;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  // (launch-uniform) vector path: every 32-column chunk is complete and all row pointers are 16-byte aligned
  if (p.out_vec_ok && (a.c_out & 31) == 0) {
    if (a.bias) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 b = __ldg((const float4*)(a.bias + (valid ? col0 : 0) + j));
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    }
    if (a.act) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], a.act);
    }
    if (a.residual) {
      uint4 t[8];
      warp_load_rows(stage, valid ? (const uint8_t*)(a.residual + (size_t)grow * a.ld_res + col0) : nullptr, t, lane);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[4 * j] += __uint_as_float(t[j].x); v[4 * j + 1] += __uint_as_float(t[j].y);
        v[4 * j + 2] += __uint_as_float(t[j].z); v[4 * j + 3] += __uint_as_float(t[j].w);
      }
    }
    if (!a.no_raw) {
      uint4 t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        t[j] = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                          __float_as_uint(v[4 * j + 3]));
      warp_store_rows(stage, t, valid ? (uint8_t*)(a.out + (size_t)grow * a.ld_out + col0) : nullptr, lane, UD3D_DBG(p, 64));
    }
#pragma unroll
    for (int oi = 0; oi < 2; ++oi) {
      if (!a.out_act[oi]) continue;
      const float* sc = a.act_scale[oi] ? a.act_scale[oi] + (valid ? col0 : 0) : nullptr;
      const float* sh = a.act_scale[oi] ? a.act_shift[oi] + (valid ? col0 : 0) : nullptr;
      const bool do_relu = !((a.act_norelu >> oi) & 1);
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float x[4] = {v[j], v[j + 1], v[j + 2], v[j + 3]};
        if (sc) {      // (16-byte aligned: part of out_vec_ok)
          const float4 s4 = __ldg((const float4*)(sc + j)), h4 = __ldg((const float4*)(sh + j));
          x[0] = fmaf(x[0], s4.x, h4.x); x[1] = fmaf(x[1], s4.y, h4.y);
          x[2] = fmaf(x[2], s4.z, h4.z); x[3] = fmaf(x[3], s4.w, h4.w);
        }
        if (do_relu) {
#pragma unroll
          for (int e = 0; e < 4; ++e) x[e] = fmaxf(x[e], 0.f);
        }
        split_bf16x2(x[0], x[1], hi[j >> 1], lo[j >> 1]);
        split_bf16x2(x[2], x[3], hi[(j >> 1) + 1], lo[(j >> 1) + 1]);
      }
      uint4 t[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        t[opf_mem_piece(j)] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
        t[opf_mem_piece(4 + j)] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
      }
      warp_store_rows(stage, t, valid ? (uint8_t*)(a.out_act[oi] + (size_t)grow * a.ld_act[oi]) + (size_t)col0 * 4 : nullptr, lane,
                      UD3D_DBG(p, 64));
    }
    return;
  }
  // generic path (ragged channel counts / unaligned views): per-thread scalar stores
  if (!valid) return;
  float* orow = a.out + (size_t)grow * a.ld_out;
  const float* rrow = a.residual ? a.residual + (size_t)grow * a.ld_res : nullptr;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    int col = col0 + j;
    if (col < a.c_out) {
      if (a.bias) v[j] += __ldg(a.bias + col);
      v[j] = apply_act(v[j], a.act);
      if (rrow) v[j] += __ldg(rrow + col);
      if (!a.no_raw) orow[col] = v[j];
    }
  }
  // operand-form outputs (c_out % 32 == 0 and 16-byte aligned rows are enforced on the host): direct 16-byte stores
#pragma unroll
  for (int oi = 0; oi < 2; ++oi) {
    if (!a.out_act[oi]) continue;
    const float* sc = a.act_scale[oi] ? a.act_scale[oi] + col0 : nullptr;
    const float* sh = a.act_scale[oi] ? a.act_shift[oi] + col0 : nullptr;
    const bool do_relu = !((a.act_norelu >> oi) & 1);
    uint4* dst = (uint4*)((uint8_t*)(a.out_act[oi] + (size_t)grow * a.ld_act[oi]) + (size_t)col0 * 4);
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      float x0 = v[j], x1 = v[j + 1];
      if (sc) {
        x0 = fmaf(x0, __ldg(sc + j), __ldg(sh + j));
        x1 = fmaf(x1, __ldg(sc + j + 1), __ldg(sh + j + 1));
      }
      if (do_relu) {
        x0 = fmaxf(x0, 0.f);
        x1 = fmaxf(x1, 0.f);
      }
      split_bf16x2(x0, x1, hi[j >> 1], lo[j >> 1]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dst[opf_mem_piece(j)] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
      dst[opf_mem_piece(4 + j)] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
    }
  }
}


}  // namespace ud3d
