// Hardware probe #2 (developer tool): sustained tcgen05.mma kind::tf32 rate of ONE CTA issuing back-to-back MMAs on
// shared-memory-resident operands, per operand layout.  Answers: what is the per-SM ceiling of the 3xTF32 conv /
// wgrad inner loops, and does the swizzle mode (SW128 vs SW64 K-major, SW128/32B-atom MN-major) change it?
#include <stdio.h>
#include <stdlib.h>
#include "../umma.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2);} } while (0)

struct RateParams { uint64_t a_desc, b_desc; uint32_t idesc; int n_mma; uint32_t a_stride, b_stride; int n_slots; int m_tiles; uint32_t tile_stride; int N; };

__global__ void __launch_bounds__(128, 1) rate_kernel(const __grid_constant__ RateParams p, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = umma::smem_u32(smem);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (threadIdx.x == 0) { umma::mbar_init(umma::smem_u32(&bar), 1); umma::fence_mbar_init(); }
  if (threadIdx.x < 32) { umma::tmem_alloc(umma::smem_u32(&slot), 512); umma::tmem_relinquish(); }
  umma::fence_proxy_async();
  umma::tc_fence_before(); __syncthreads(); umma::tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32) {
    long long t0 = clock64();
    for (int i = 0; i < p.n_mma; ++i) {
      const int s = i % p.n_slots, t = (i / 3) % p.m_tiles;
      const uint64_t ad = umma::desc_at(p.a_desc, sbase + s * p.a_stride + t * p.tile_stride);
      const uint64_t bd = umma::desc_at(p.b_desc, sbase + 131072 + s * p.b_stride);
      if (umma::elect_one()) umma::mma_tf32_ss(tmem + t * p.N, ad, bd, p.idesc, i >= p.m_tiles * 3);
    }
    if (umma::elect_one()) umma::mma_commit(umma::smem_u32(&bar));
    long long t1 = clock64();
    umma::mbar_wait(umma::smem_u32(&bar), 0);
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  umma::tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) umma::tmem_dealloc(tmem, 512);
}


// Variant 2: loop structure of the conv kernel.  Outer loop over B stages (runtime ring index), inner 12 MMAs fully
// unrolled with compile-time offsets from two per-stage base addresses.
template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel2(int n_outer, int n_stages, uint32_t a_stage_bytes, uint32_t b_stage_bytes, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = umma::smem_u32(smem);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (threadIdx.x == 0) { umma::mbar_init(umma::smem_u32(&bar), 1); umma::fence_mbar_init(); }
  if (threadIdx.x < 32) { umma::tmem_alloc(umma::smem_u32(&slot), 512); umma::tmem_relinquish(); }
  umma::fence_proxy_async();
  umma::tc_fence_before(); __syncthreads(); umma::tc_fence_after();
  const uint32_t tmem = slot;
  constexpr uint64_t desc = umma::make_desc_base(16, 512, umma::LAYOUT_SW64);
  constexpr uint32_t idesc = umma::make_idesc_tf32(128, N, 0, 0);
  if (threadIdx.x < 32) {
    long long t0 = clock64();
    int stage = 0;
    for (int it = 0; it < n_outer; ++it) {
      const uint32_t av = sbase + stage * a_stage_bytes, al = av + 16384 * 2;
      const uint32_t bv = sbase + 131072 + stage * b_stage_bytes, bl = bv + 16384;
      if (umma::elect_one()) {
#pragma unroll
        for (int tile = 0; tile < 2; ++tile)
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint32_t ao = tile * 8192 + k * 32, bo = k * 32, d = tmem + tile * N;
            umma::mma_tf32_ss(d, umma::desc_at(desc, al + ao), umma::desc_at(desc, bv + bo), idesc, (it | k) != 0);
            umma::mma_tf32_ss(d, umma::desc_at(desc, av + ao), umma::desc_at(desc, bl + bo), idesc, 1);
            umma::mma_tf32_ss(d, umma::desc_at(desc, av + ao), umma::desc_at(desc, bv + bo), idesc, 1);
          }
      }
      __syncwarp();
      if (++stage == n_stages) stage = 0;
    }
    if (umma::elect_one()) umma::mma_commit(umma::smem_u32(&bar));
    long long t1 = clock64();
    umma::mbar_wait(umma::smem_u32(&bar), 0);
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  umma::tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) umma::tmem_dealloc(tmem, 512);
}

template <int N> void run2(long long* d) {
  CK(cudaFuncSetAttribute(rate_kernel2<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  rate_kernel2<N><<<1, 128, 210 * 1024>>>(256, 3, 1024, 2048, d);
  CK(cudaDeviceSynchronize());
  long long h[2]; CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
  const double ideal = 128.0 * N * 8 * 2 / 4096.0, n = 256 * 12;
  printf("variant2 (unrolled 12 MMAs/stage, elect) N=%d: issue %.1f clk/MMA, complete %.1f clk/MMA, floor %.1f -> %.0f%%\n", N, h[0] / n, h[1] / n, ideal,
         100.0 * ideal / (h[1] / n));
}

int main() {
  { long long* d2; CK(cudaMalloc(&d2, 16)); run2<96>(d2); run2<48>(d2); run2<144>(d2); }
  CK(cudaSetDevice(0));
  long long* d; CK(cudaMalloc(&d, 16));
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  struct Cfg { const char* name; uint64_t ad, bd; int am, bm; int N; uint32_t as, bs; };
  const uint64_t k128 = umma::make_desc_base(16, 1024, umma::LAYOUT_SW128), k64 = umma::make_desc_base(16, 512, umma::LAYOUT_SW64);
  const uint64_t mn = umma::make_desc_base(4096, 512, 1);
  Cfg cfgs[] = {
      {"K-major SW128 A,B  N=96 ", k128, k128, 0, 0, 96, 32, 32},
      {"K-major SW64  A,B  N=96 ", k64, k64, 0, 0, 96, 32, 32},
      {"K-major SW64  A,B  N=48 ", k64, k64, 0, 0, 48, 32, 32},
      {"K-major SW64  A,B  N=144", k64, k64, 0, 0, 144, 32, 32},
      {"K-major SW128 A,B  N=192", k128, k128, 0, 0, 192, 32, 32},
      {"K-major SW128 A,B  N=256", k128, k128, 0, 0, 256, 32, 32},
      {"MN-major 32B-atom  N=96 ", mn, mn, 1, 1, 96, 1024, 1024},
      {"MN-major 32B-atom  N=48 ", mn, mn, 1, 1, 48, 1024, 1024},
  };
  for (auto& c : cfgs) for (int tiles : {1, 2}) {
    if (tiles * c.N > 512) continue;
    RateParams p{}; p.a_desc = c.ad; p.b_desc = c.bd; p.idesc = umma::make_idesc_tf32(128, c.N, c.am, c.bm); p.n_mma = 3072;
    p.a_stride = c.as; p.b_stride = c.bs; p.n_slots = 2; p.m_tiles = tiles; p.tile_stride = 16384; p.N = c.N;
    rate_kernel<<<1, 128, 210 * 1024>>>(p, d);
    CK(cudaDeviceSynchronize());
    long long h[2]; CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    const double ideal = 128.0 * c.N * 8 * 2 / 4096.0;
    printf("%s tiles=%d : issue %.1f clk/MMA, complete %.1f clk/MMA (math floor %.1f clk at 4096 tf32 FLOP/clk/SM) -> %.0f%% of floor\n", c.name, tiles,
           (double)h[0] / p.n_mma, (double)h[1] / p.n_mma, ideal, 100.0 * ideal / ((double)h[1] / p.n_mma));
  }
  return 0;
}
