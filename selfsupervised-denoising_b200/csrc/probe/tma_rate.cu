// Hardware probe #3 (developer tool): sustained TMA (cp.async.bulk.tensor.2d) load throughput per SM as a function of
// the box shape, with all SMs active.  One elected thread per CTA keeps `depth` loads in flight into a smem ring.
#include <stdio.h>
#include <stdlib.h>
#include "../umma.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2);} } while (0)

__global__ void __launch_bounds__(64, 1) tma_kernel(const __grid_constant__ CUtensorMap map, int box_cols, int box_rows, int n_loads, int depth,
                                                    int rows_total, int same_region, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[16];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = umma::smem_u32(smem);
  const uint32_t bytes = box_cols * 4 * box_rows;
  const uint32_t slot = (bytes + 1023) / 1024 * 1024;
  if (threadIdx.x == 0) { for (int i = 0; i < depth; ++i) umma::mbar_init(umma::smem_u32(&bars[i]), 1); umma::fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int region = rows_total - box_rows;
    uint32_t r = same_region ? 0 : (blockIdx.x * 7919u) % region;
    long long t0 = clock64();
    for (int i = 0; i < n_loads + depth; ++i) {
      const int s = i % depth;
      if (i >= depth) umma::mbar_wait(umma::smem_u32(&bars[s]), ((i / depth) - 1) & 1);
      if (i < n_loads && umma::elect_one()) {
        umma::mbar_expect_tx(umma::smem_u32(&bars[s]), bytes);
        umma::tma_load_2d(sbase + s * slot, &map, umma::smem_u32(&bars[s]), 0, (int)r);
      }
      __syncwarp();
      r = (r + box_rows) % region;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
}

int main() {
  CK(cudaSetDevice(0));
  const int rows_total = 1 << 18;                     // 262144 rows x 128 B = 32 MB: L2 resident
  float* d; CK(cudaMalloc(&d, (size_t)rows_total * 32 * 4)); CK(cudaMemset(d, 0, (size_t)rows_total * 32 * 4));
  long long* o; CK(cudaMalloc(&o, 148 * 8));
  CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  struct Cfg { int cols, rows; CUtensorMapSwizzle sw; const char* name; };
  Cfg cfgs[] = {{16, 200, CU_TENSOR_MAP_SWIZZLE_64B, "16ch x 200 rows SW64 (conv A box)"}, {16, 96, CU_TENSOR_MAP_SWIZZLE_64B, "16ch x 96 rows SW64 (conv B box)"},
                {32, 100, CU_TENSOR_MAP_SWIZZLE_128B, "32ch x 100 rows SW128"}, {32, 256, CU_TENSOR_MAP_SWIZZLE_128B, "32ch x 256 rows SW128"},
                {32, 40, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, "32ch x 40 rows (wgrad X box)"}, {32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, "32ch x 32 rows (wgrad dZ box)"},
                {8, 256, CU_TENSOR_MAP_SWIZZLE_32B, "8ch x 256 rows SW32"}};
  for (auto& c : cfgs) for (int same : {0, 1}) for (int depth : {2, 8}) {
    CUtensorMap m; uint64_t dims[2] = {32, (uint64_t)rows_total}; uint64_t str[1] = {128}; uint32_t box[2] = {(uint32_t)c.cols, (uint32_t)c.rows};
    if (umma::encode_f32(&m, d, 2, dims, str, box, c.sw)) { printf("encode failed\n"); return 1; }
    const int n_loads = 400;
    const size_t slot = ((size_t)c.cols * 4 * c.rows + 1023) / 1024 * 1024;
    tma_kernel<<<148, 64, depth * slot + 1024>>>(m, c.cols, c.rows, n_loads, depth, rows_total, same, o);
    CK(cudaDeviceSynchronize());
    long long h[148]; CK(cudaMemcpy(h, o, sizeof(h), cudaMemcpyDeviceToHost));
    double mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const double bytes = (double)n_loads * c.cols * 4 * c.rows;
    printf("%-36s %s depth %d: %6.1f B/clk/SM  (%.0f clk per load of %d B, %.1f clk per box row)\n", c.name, same ? "same region " : "spread      ", depth,
           bytes / mx, mx / n_loads, c.cols * 4 * c.rows, mx / n_loads / c.rows);
  }
  return 0;
}
