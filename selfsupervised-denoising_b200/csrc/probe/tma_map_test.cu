// Probe #4: does cuTensorMapEncodeTiled accept (a) a 4-D map whose channel-block dimension has a SMALLER stride than the
// row dimension, (b) a plane dimension, and does the box land in smem as [plane][block][row][32 ch]?
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../umma.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2);} } while (0)
__global__ void k(const __grid_constant__ CUtensorMap map, float* out, int nfloats, int r0) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (threadIdx.x == 0) { umma::mbar_init(umma::smem_u32(&bar), 1); umma::fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    umma::mbar_expect_tx(umma::smem_u32(&bar), nfloats * 4);
    umma::tma_load_4d(umma::smem_u32(smem), &map, umma::smem_u32(&bar), 0, r0, 0, 0);
    umma::mbar_wait(umma::smem_u32(&bar), 0);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}
int main() {
  const int rows = 64, C = 96, pitch = 100;         // channels-last [plane][row][pitch]
  std::vector<float> h(2 * rows * pitch);
  for (int p = 0; p < 2; ++p) for (int r = 0; r < rows; ++r) for (int c = 0; c < pitch; ++c) h[(p * rows + r) * pitch + c] = p * 100000 + r * 100 + c;
  float* d; CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  auto fn = umma::get_encode_fn();
  CUtensorMap m;
  cuuint64_t dims[4] = {32, (cuuint64_t)rows, 3, 2};                       // c_in, row, channel block, plane
  cuuint64_t str[3] = {(cuuint64_t)pitch * 4, 128, (cuuint64_t)rows * pitch * 4};
  cuuint32_t box[4] = {32, 8, 3, 2}, es[4] = {1, 1, 1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode (non-monotonic strides, 4-D) -> %d\n", (int)r);
  if (r != CUDA_SUCCESS) return 0;
  const int nf = 32 * 8 * 3 * 2;
  float* o; CK(cudaMalloc(&o, nf * 4));
  k<<<1, 128, 32768>>>(m, o, nf, 5);
  CK(cudaDeviceSynchronize());
  std::vector<float> g(nf); CK(cudaMemcpy(g.data(), o, nf * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int p = 0; p < 2; ++p) for (int b = 0; b < 3; ++b) for (int rr = 0; rr < 8; ++rr) for (int c = 0; c < 32; ++c) {
    const int ch = b * 32 + c;
    // NOTE: dims say 3 blocks x 32 = 96 channels; pitch is 100, so nothing is out of bounds here
    float want = p * 100000 + (5 + rr) * 100 + ch;
    float got = g[((p * 3 + b) * 8 + rr) * 32 + c];
    if (want != got) { if (bad < 5) printf("mismatch p%d b%d r%d c%d: got %g want %g\n", p, b, rr, c, got, want); ++bad; }
  }
  printf("smem layout [plane][block][row][32]: %s (%d mismatches)\n", bad ? "WRONG" : "OK", bad);
  return 0;
}
