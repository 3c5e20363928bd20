// Micro-benchmark: how fast can one SM gather 128-byte rows from L2 into shared memory?
// (the A-operand producer of the gather-GEMM; no MMA, no epilogue).  Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/gather_probe tools/gather_probe.cu && tools/gather_probe
// Modes: 0 LDGSTS.cg 16 B/lane (8 lanes per row)   1 LDGSTS.ca   2 LDG.128 + STS.128   3 cp.async.bulk 128 B per row
//        (one thread per row)   4 LDGSTS.cg on a compacted (dst,src) list (all lanes busy)   5 bulk, one warp issues
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kK = 27, kTile = 128, kStages = 6;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    if (++spins > (1u << 24)) __trap();
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int MODE, int D, int kThreads>
__global__ void __launch_bounds__(kThreads) probe(const uint8_t* __restrict__ in, const int32_t* __restrict__ table, int n_rows,
                                                  int n_tiles, unsigned long long* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  int32_t* s_tbl = (int32_t*)(sA + kStages * kTile * 128);   // [27][128]
  int32_t* s_cnt = s_tbl + kK * kTile;                       // [27]
  uint64_t* bars = (uint64_t*)(s_cnt + 32);                  // [kStages]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j = tid & 7, rbase = tid >> 3;
  if (tid == 0)
    for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], MODE == 5 ? 32 : kTile);
  __syncthreads();
  uint32_t phase_bits = 0;   // per-stage parity
  unsigned long long acc = 0;
  long long t_pro = 0, t_main = 0, t_drain = 0, n_t = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int m0 = tile * kTile;
    __syncthreads();
    const long long c0 = clock64();
    {
      constexpr int kPer = (kK * kTile + kThreads - 1) / kThreads;
      int vals[kPer];
#pragma unroll
      for (int q = 0; q < kPer; ++q) {
        const int i = tid + q * kThreads;
        const int k = i >> 7, r = i & 127;
        vals[q] = (i < kK * kTile && m0 + r < n_rows) ? __ldg(table + (size_t)k * n_rows + m0 + r) : -1;
      }
#pragma unroll
      for (int q = 0; q < kPer; ++q) {
        const int i = tid + q * kThreads;
        if (i < kK * kTile) s_tbl[i] = vals[q];
      }
    }
    __syncthreads();
    if (MODE == 4) {
      // compact each offset's valid entries: s_tbl[k][e] = dst_row | src << 7
      for (int k = warp; k < kK; k += kThreads / 32) {
        int vals[4], base = 0;
        for (int i = 0; i < 4; ++i) vals[i] = s_tbl[k * kTile + i * 32 + lane];
        __syncwarp();
        for (int i = 0; i < 4; ++i) {
          const unsigned b = __ballot_sync(0xffffffffu, vals[i] >= 0);
          if (vals[i] >= 0) s_tbl[k * kTile + base + __popc(b & ((1u << lane) - 1))] = (i * 32 + lane) | (vals[i] << 7);
          base += __popc(b);
        }
        if (lane == 0) s_cnt[k] = base;
      }
      __syncthreads();
    }
    const long long c1 = clock64();
    for (int k = 0; k < kK; ++k) {
      const int s = k % kStages;
      const uint32_t as = smem_u32(sA + s * kTile * 128);
      const int32_t* trow = s_tbl + k * kTile;
      if (MODE == 0 || MODE == 1) {
#pragma unroll
        for (int i = 0; i < 1024 / kThreads; ++i) {
          const int r = rbase + (kThreads / 8) * i;
          const int idx = trow[r];
          if (idx >= 0) {
            const uint32_t dst = as + r * 128 + ((j ^ (r & 7)) << 4);
            const uint8_t* src = in + (size_t)idx * 128 + j * 16;
            if (MODE == 0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            else asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" ::"n"(D) : "memory");
      } else if (MODE == 6) {
#pragma unroll
        for (int i = 0; i < 1024 / kThreads; ++i) {
          const int r = rbase + (kThreads / 8) * i;
          const int idx = trow[r];
          const uint32_t dst = as + r * 128 + ((j ^ (r & 7)) << 4);
          const uint8_t* src = in + (size_t)(idx < 0 ? 0 : idx) * 128 + j * 16;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(idx < 0 ? 0u : 16u) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" ::"n"(D) : "memory");
      } else if (MODE == 4) {
        const int cnt = s_cnt[k];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int e = rbase + 32 * i;
          if (e < cnt) {
            const int ent = trow[e];
            const int r = ent & 127, idx = ent >> 7;
            const uint32_t dst = as + r * 128 + ((j ^ (r & 7)) << 4);
            const uint8_t* src = in + (size_t)idx * 128 + j * 16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" ::"n"(D) : "memory");
      } else if (MODE == 2) {
        uint4 v[4];
        int ok[4];
#pragma unroll
        for (int i = 0; i < 1024 / kThreads; ++i) {
          const int r = rbase + (kThreads / 8) * i;
          const int idx = trow[r];
          ok[i] = idx >= 0;
          if (ok[i]) v[i] = __ldg((const uint4*)(in + (size_t)idx * 128 + j * 16));
        }
#pragma unroll
        for (int i = 0; i < 1024 / kThreads; ++i) {
          const int r = rbase + (kThreads / 8) * i;
          if (ok[i]) *(uint4*)(sA + s * kTile * 128 + r * 128 + ((j ^ (r & 7)) << 4)) = v[i];
        }
      } else if (MODE == 3) {
        if (tid < kTile) {
          const int idx = trow[tid];
          if (idx >= 0) {
            mbar_arrive_tx(&bars[s], 128);
            bulk_g2s(as + tid * 128, in + (size_t)idx * 128, 128, &bars[s]);
          } else {
            mbar_arrive(&bars[s]);
          }
        }
        // D = 1: wait for the previous step's rows
        if (k > 0 && tid < kTile) {
          const int ps = (k - 1) % kStages;
          mbar_wait(&bars[ps], (phase_bits >> ps) & 1u);
          phase_bits ^= 1u << ps;
        }
      } else if (MODE == 5) {
        if (warp == 0) {
          int nb = 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = lane + 32 * i;
            const int idx = trow[r];
            if (idx >= 0) { bulk_g2s(as + r * 128, in + (size_t)idx * 128, 128, &bars[s]); nb += 128; }
          }
          if (nb) mbar_arrive_tx(&bars[s], nb); else mbar_arrive(&bars[s]);
        }
        if (k > 0 && warp == 0) {
          const int ps = (k - 1) % kStages;
          mbar_wait(&bars[ps], (phase_bits >> ps) & 1u);
          phase_bits ^= 1u << ps;
        }
      }
    }
    const long long c2 = clock64();
    if (MODE == 0 || MODE == 1 || MODE == 4 || MODE == 6) asm volatile("cp.async.wait_group 0;" ::: "memory");
    if ((MODE == 3 && tid < kTile) || (MODE == 5 && warp == 0)) {
      const int ps = (kK - 1) % kStages;
      mbar_wait(&bars[ps], (phase_bits >> ps) & 1u);
      phase_bits ^= 1u << ps;
    }
    __syncthreads();
    const long long c3 = clock64();
    t_pro += c1 - c0; t_main += c2 - c1; t_drain += c3 - c2; ++n_t;
    acc += *(const uint32_t*)(sA + (tid * 16) % (kStages * kTile * 128));
  }
  if (blockIdx.x == 7 && tid == 0) { sink[1] = t_pro / n_t; sink[2] = t_main / n_t; sink[3] = t_drain / n_t; }
  if (acc == 0x123456789ull) *sink = acc;
}

static uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int MODE, int D, int kThreads = 256>
static void run(const char* name, const uint8_t* d_in, const int32_t* d_tbl, int n_rows, int ctas_per_sm, double valid_rows,
                unsigned long long* sink) {
  const int n_tiles = (n_rows + kTile - 1) / kTile;
  size_t smem = 1024 + kStages * kTile * 128 + kK * kTile * 4 + 256;
  CK(cudaFuncSetAttribute(probe<MODE, D, kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e9f;
  for (int it = 0; it < 5; ++it) {
    CK(cudaEventRecord(e0));
    probe<MODE, D, kThreads><<<148 * ctas_per_sm, kThreads, smem>>>(d_in, d_tbl, n_rows, n_tiles, sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (it > 0 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  const double steps_per_sm = (double)n_tiles * kK / 148.0;
  const double cyc = best * 1e-3 * 1.965e9;
  unsigned long long hs[4];
  CK(cudaMemcpy(hs, sink, 32, cudaMemcpyDeviceToHost));
  printf("[tile: prologue %5llu main %6llu drain %5llu cyc] ", hs[1], hs[2], hs[3]);
  printf("%-28s thr %4d ctas/SM %d: %8.1f us  %7.0f cyc/step/SM  %6.1f B/clk/SM gathered\n", name, kThreads, ctas_per_sm, best * 1e3,
         cyc / steps_per_sm, valid_rows * 128.0 / 148.0 / cyc);
}

int main(int argc, char** argv) {
  const int n_rows = 261120;
  uint8_t* d_in; int32_t* d_tbl; unsigned long long* sink;
  CK(cudaMalloc(&d_in, (size_t)n_rows * 128));
  CK(cudaMemset(d_in, 1, (size_t)n_rows * 128));
  CK(cudaMalloc(&d_tbl, (size_t)kK * n_rows * 4));
  CK(cudaMalloc(&sink, 64));
  for (int pi = 0; pi < 2; ++pi) {
    const double p = pi == 0 ? 0.4 : 1.0;
    std::vector<int32_t> tbl((size_t)kK * n_rows);
    double valid = 0;
    for (int k = 0; k < kK; ++k)
      for (int r = 0; r < n_rows; ++r) {
        // neighbour of row r for offset k: a nearby row (runs of consecutive rows, like z-fastest voxel order)
        const uint32_t h = hash32((uint32_t)(r / 4) * 31u + k * 7919u);
        const bool ok = (h & 0xffff) < (uint32_t)(p * 65536.0);
        long long src = (long long)r + (k - 13) * 997;
        if (src < 0) src += n_rows;
        if (src >= n_rows) src -= n_rows;
        tbl[(size_t)k * n_rows + r] = ok ? (int32_t)src : -1;
        valid += ok;
      }
    CK(cudaMemcpy(d_tbl, tbl.data(), tbl.size() * 4, cudaMemcpyHostToDevice));
    printf("--- valid fraction %.2f (%.0f row gathers of 128 B)\n", p, valid);
    for (int c = 1; c <= 2; ++c) {
      run<0, 1, 256>("LDGSTS.cg D=1", d_in, d_tbl, n_rows, c, valid, sink);
      run<0, 3, 256>("LDGSTS.cg D=3", d_in, d_tbl, n_rows, c, valid, sink);
      run<0, 1, 512>("LDGSTS.cg D=1", d_in, d_tbl, n_rows, c, valid, sink);
      run<6, 1, 256>("LDGSTS.cg zfill D=1", d_in, d_tbl, n_rows, c, valid, sink);
      run<6, 3, 256>("LDGSTS.cg zfill D=3", d_in, d_tbl, n_rows, c, valid, sink);
      run<4, 1, 256>("LDGSTS.cg compact D=1", d_in, d_tbl, n_rows, c, valid, sink);
      run<4, 3, 256>("LDGSTS.cg compact D=3", d_in, d_tbl, n_rows, c, valid, sink);
    }
  }
  return 0;
}
