// Full-chip tcgen05.mma rate measurement (the denominator of the tensor roofline bench.py reports).
// One CTA per SM; the elected thread of warp 0 issues back-to-back MMAs of M = 128 (cta_group::1) or M = 256
// (cta_group::2, clusters of two CTAs) x N = 256 on operands that stay resident in shared memory, alternating between two
// TMEM accumulators.  No TMA and no epilogue: this is what the tensor pipe sustains, chip-wide, at the clocks it gets.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "umma.cuh"

namespace peakk {

struct PeakResult { double tflops, seconds, flop_per_clk_sm, sm_mhz; int launches; };

__device__ __forceinline__ void mma_any(bool f16, bool pair, uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
  if (f16) {
    if (pair)
      asm volatile("{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
                   "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc) : "memory");
    else
      asm volatile("{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
                   "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc) : "memory");
  } else {
    if (pair)
      asm volatile("{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
                   "tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %4, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc) : "memory");
    else
      asm volatile("{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
                   "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc) : "memory");
  }
}

constexpr int kMmaPerIter = 8;      // 4 k-steps of one 128-byte operand row x 2 accumulators

template <bool F16, bool PAIR>
__global__ void __launch_bounds__(128, 1) peak_kernel(int n_iter, long long* __restrict__ clocks) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = umma::smem_u32(smem);
  // operands: 16 KB of A (128 rows x 128 B) + 32 KB of B (256 rows x 128 B), pseudo-random finite values of modest range
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    uint32_t w;
    if (F16) w = (h & 0x83ff83ffu) | 0x38003800u;          // two halves with |x| in [0.5, 1), random sign and mantissa
    else w = (h & 0x807fe000u) | 0x3f000000u;               // tf32-exact float with |x| in [0.5, 1)
    reinterpret_cast<uint32_t*>(smem)[i] = w;
  }
  const uint32_t rank = PAIR ? umma::cluster_ctarank() : 0;
  if (threadIdx.x == 0) { umma::mbar_init(umma::smem_u32(&bar), 1); umma::fence_mbar_init(); }
  if (threadIdx.x < 32) {
    if (PAIR) { umma::tmem_alloc_pair(umma::smem_u32(&slot), 512); umma::tmem_relinquish_pair(); }
    else { umma::tmem_alloc(umma::smem_u32(&slot), 512); umma::tmem_relinquish(); }
  }
  umma::fence_proxy_async();
  umma::tc_fence_before(); __syncthreads();
  if (PAIR) umma::cluster_sync();
  umma::tc_fence_after();
  const uint32_t tmem = slot;
  constexpr uint64_t desc = umma::make_desc_base(16, 1024, umma::LAYOUT_SW128);
  const uint32_t hi = (uint32_t)(desc >> 32), lbo = (uint32_t)(desc & 0xffff0000u);
  // kind::f16: a/b format 0 (fp16); kind::tf32: format 2; fp32 accumulate; K-major; N = 256; M = 128 or 256
  const uint32_t idesc = (1u << 4) | ((F16 ? 0u : 2u) << 7) | ((F16 ? 0u : 2u) << 10) | ((256u >> 3) << 17) | (((PAIR ? 256u : 128u) >> 4) << 24);
  if (threadIdx.x < 32) {
    const long long t0 = clock64();
    if (!PAIR || rank == 0) {
      const uint32_t a0 = ((sbase >> 4) & 0x3fffu) | lbo, b0 = (((sbase + 16384) >> 4) & 0x3fffu) | lbo;
      for (int it = 0; it < n_iter; ++it) {
        if (umma::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            mma_any(F16, PAIR, tmem, a0 + 2 * k, b0 + 2 * k, hi, idesc);
            mma_any(F16, PAIR, tmem + 256, a0 + 2 * k, b0 + 2 * k, hi, idesc);
          }
        }
        __syncwarp();
      }
      if (umma::elect_one()) { if (PAIR) umma::mma_commit_pair(umma::smem_u32(&bar)); else umma::mma_commit(umma::smem_u32(&bar)); }
      __syncwarp();
    }
    for (uint32_t i = 0; i < (1u << 26); ++i) if (umma::mbar_try_wait(umma::smem_u32(&bar), 0)) break;   // bounded: a probe must not hang the box
    const long long t1 = clock64();
    if (threadIdx.x == 0 && clocks) clocks[blockIdx.x] = t1 - t0;
  }
  umma::tc_fence_before(); __syncthreads();
  if (PAIR) umma::cluster_sync();
  if (threadIdx.x < 32) { if (PAIR) umma::tmem_dealloc_pair(tmem, 512); else umma::tmem_dealloc(tmem, 512); }
}

template <bool F16, bool PAIR>
static inline cudaError_t launch(int grid, int n_iter, long long* clocks, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(peak_kernel<F16, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 50 * 1024; cfg.stream = st;
  cudaLaunchAttribute a[1];
  a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = PAIR ? 2 : 1; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
  cfg.attrs = a; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, peak_kernel<F16, PAIR>, n_iter, clocks);
}
static inline cudaError_t launch_any(int f16, int pair, int grid, int n_iter, long long* clocks, cudaStream_t st) {
  if (f16) return pair ? launch<true, true>(grid, n_iter, clocks, st) : launch<true, false>(grid, n_iter, clocks, st);
  return pair ? launch<false, true>(grid, n_iter, clocks, st) : launch<false, false>(grid, n_iter, clocks, st);
}

// Runs launches of ~8 ms for `seconds`; returns the sustained rate over all of them (CUDA events around the whole series).
static inline int measure(int f16, int pair, int sms, double seconds, cudaStream_t st, PeakResult* out) {
  const int grid = pair ? sms - sms % 2 : sms;
  long long* d_clk = nullptr;
  if (cudaMalloc(&d_clk, sizeof(long long) * grid) != cudaSuccess) return -1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timed = [&](int n_iter, int reps, float* ms) -> cudaError_t {
    cudaEventRecord(e0, st);
    for (int r = 0; r < reps; ++r) { cudaError_t e = launch_any(f16, pair, grid, n_iter, d_clk, st); if (e != cudaSuccess) return e; }
    cudaEventRecord(e1, st);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) return e;
    cudaEventElapsedTime(ms, e0, e1);
    return cudaGetLastError();
  };
  float ms = 0;
  int rc = 0;
  if (timed(2000, 1, &ms) != cudaSuccess) rc = -2;            // warm-up / calibration
  if (!rc && timed(2000, 1, &ms) != cudaSuccess) rc = -2;
  if (!rc) {
    const int n_iter = (int)(2000.0 * 8.0 / (ms > 0.01f ? ms : 0.01f));      // ~8 ms per launch
    const int reps = (int)(seconds * 1000.0 / 8.0) + 1;
    if (timed(n_iter, reps, &ms) != cudaSuccess) rc = -3;
    else {
      std::vector<long long> h(grid);
      cudaMemcpy(h.data(), d_clk, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
      double clk = 0; for (long long c : h) clk += (double)c / grid;
      const double m = pair ? 256.0 : 128.0, k = f16 ? 16.0 : 8.0;
      const double flop_launch = 2.0 * m * 256.0 * k * kMmaPerIter * (double)n_iter * (pair ? grid / 2 : grid);
      out->tflops = flop_launch * reps / (ms * 1e-3) / 1e12;
      out->seconds = ms * 1e-3; out->launches = reps;
      out->flop_per_clk_sm = flop_launch / grid / clk;
      out->sm_mhz = clk / (ms * 1e-3 / reps) / 1e6;
    }
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_clk);
  return rc;
}

}  // namespace peakk
