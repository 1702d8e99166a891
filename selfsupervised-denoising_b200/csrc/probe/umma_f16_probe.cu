// Hardware probe #3 (developer tool for the NEXT step of DESIGN.md section 8, item 1; not part of the product path and
// NOT yet run on a GPU): tcgen05.mma kind::f16 with fp16 operand planes, as a two-term split of fp32 data.
//   F1  K-major SWIZZLE_128B, 64 halves per row (4 MMAs of K = 16 per row, +32 B each)
//   F2  K-major SWIZZLE_64B, 32-channel chunks (2 MMAs per chunk; F2': half chunk by TMA zero fill)   F3  K-major SWIZZLE_32B, 16-channel chunks
//   F4  row-shifted A start address for F2 / F3 layouts (the implicit-GEMM halo reuse), base_offset 0 vs (r & 7)
//   F5  fp16x2 split of fp32 data with a power-of-two scale: hi*hi + lo*hi + hi*lo, against fp64 and against 3xTF32's error
//   F6  MN-major A / B (the weight-gradient operands), SWIZZLE_128B with 64-half atoms, both LBO / SBO conventions
//   F7  sustained issue rate of kind::f16 M = 128 MMAs per N (expected: the clocks of a kind::tf32 MMA at twice the K)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_f16_probe umma_f16_probe.cu
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../umma.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2);} } while (0)

namespace f16p {
// Instruction descriptor for kind::f16, fp16 A and B (format 0), fp32 accumulate (c_format 1); same field positions as
// make_idesc_tf32 (cute/arch/mma_sm100_desc.hpp: a_format [7,10), b_format [10,13), a_major 15, b_major 16, N>>3 [17,23), M>>4 [24,29)).
__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t M, uint32_t N, uint32_t a_mn_major, uint32_t b_mn_major) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((a_mn_major & 1u) << 15) | ((b_mn_major & 1u) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
inline int encode_f16(CUtensorMap* m, void* base, uint64_t inner, uint64_t outer, uint32_t box_in, uint32_t box_out, CUtensorMapSwizzle sw) {
  auto fn = umma::get_encode_fn();
  if (!fn) return -1;
  cuuint64_t d[2] = {inner, outer}, s[1] = {inner * 2};
  cuuint32_t b[2] = {box_in, box_out}, e[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}
}  // namespace f16p

struct ProbeLoad { int map; uint32_t smem_off; int c0, c1; uint32_t bytes; };
struct ProbeMma { uint32_t a_off, b_off; };
struct ProbeParams {
  int n_loads; ProbeLoad loads[16];
  int n_mma;   ProbeMma mma[48];
  uint64_t a_desc, b_desc;
  uint32_t idesc; int N; int f16;       // f16 = 0 runs the same schedule with kind::tf32 (comparison rows of F5)
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ ProbeParams p,
             float* __restrict__ out, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_base_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = umma::smem_u32(smem);
  const uint32_t bar_ld = umma::smem_u32(&bars[0]), bar_mma = umma::smem_u32(&bars[1]);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    umma::mbar_init(bar_ld, 1);
    umma::mbar_init(bar_mma, 1);
    umma::fence_mbar_init();
  }
  if (warp == 1) {
    umma::tmem_alloc(umma::smem_u32(&tmem_base_slot), 128);
    umma::tmem_relinquish();
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  if (threadIdx.x == 0) {
    uint32_t total = 0;
    for (int i = 0; i < p.n_loads; ++i) total += p.loads[i].bytes;
    umma::mbar_expect_tx(bar_ld, total);
    for (int i = 0; i < p.n_loads; ++i) {
      const ProbeLoad& l = p.loads[i];
      umma::tma_load_2d(sbase + l.smem_off, l.map == 0 ? &mapA : &mapB, bar_ld, l.c0, l.c1);
    }
    if (umma::mbar_wait(bar_ld, 0)) {
      umma::tc_fence_after();
      for (int i = 0; i < p.n_mma; ++i) {
        const uint64_t ad = umma::desc_at(p.a_desc, sbase + p.mma[i].a_off), bd = umma::desc_at(p.b_desc, sbase + p.mma[i].b_off);
        if (p.f16) f16p::mma_f16_ss(tmem, ad, bd, p.idesc, i > 0);
        else umma::mma_tf32_ss(tmem, ad, bd, p.idesc, i > 0);
      }
      umma::mma_commit(bar_mma);
    } else {
      *status = 1;
    }
  }
  __syncthreads();
  if (*reinterpret_cast<volatile int*>(status) == 0) {
    if (!umma::mbar_wait(bar_mma, 0)) { if ((threadIdx.x & 31) == 0) *status = 2; }
    umma::tc_fence_after();
    const int row = warp * 32 + (threadIdx.x & 31);
    for (int c = 0; c < p.N; c += 16) {
      uint32_t r[16];
      umma::tmem_ld16(tmem + (uint32_t(warp * 32) << 16) + c, r);
      umma::tmem_ld_wait();
      for (int j = 0; j < 16; ++j) out[row * p.N + c + j] = __uint_as_float(r[j]);
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem, 128);
}

// F7: one warp issues n_mma back-to-back MMAs on resident operands (two operand slots, one accumulator tile).
__global__ void __launch_bounds__(128, 1) rate_kernel(uint64_t desc, uint32_t idesc, int n_mma, int f16, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = umma::smem_u32(smem);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;   // halves 1.0 (tf32: small floats)
  if (threadIdx.x == 0) { umma::mbar_init(umma::smem_u32(&bar), 1); umma::fence_mbar_init(); }
  if (threadIdx.x < 32) { umma::tmem_alloc(umma::smem_u32(&slot), 512); umma::tmem_relinquish(); }
  umma::fence_proxy_async();
  umma::tc_fence_before(); __syncthreads(); umma::tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32) {
    const long long t0 = clock64();
    if (umma::elect_one()) {
      for (int i = 0; i < n_mma; ++i) {
        const uint64_t ad = umma::desc_at(desc, sbase + (i & 1) * 32), bd = umma::desc_at(desc, sbase + 65536 + (i & 1) * 32);
        if (f16) f16p::mma_f16_ss(tmem, ad, bd, idesc, i > 0);
        else umma::mma_tf32_ss(tmem, ad, bd, idesc, i > 0);
      }
      umma::mma_commit(umma::smem_u32(&bar));
    }
    __syncwarp();
    umma::mbar_wait(umma::smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
  }
  umma::tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) umma::tmem_dealloc(tmem, 512);
}

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }
static float tf32_rn(float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x0FFFu + ((u >> 13) & 1u)) & 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

template <class T> static T* upload(const std::vector<T>& h) { T* d; CK(cudaMalloc(&d, h.size() * sizeof(T))); CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice)); return d; }
static CUtensorMap map_h(__half* base, uint64_t inner, uint64_t outer, uint32_t box_in, uint32_t box_out, CUtensorMapSwizzle sw) {
  CUtensorMap m; int r = f16p::encode_f16(&m, base, inner, outer, box_in, box_out, sw);
  if (r) { printf("fp16 tensor map encode failed %d (inner %llu box %u x %u)\n", r, (unsigned long long)inner, box_in, box_out); exit(3); }
  return m;
}
static CUtensorMap map_f(float* base, uint64_t inner, uint64_t outer, uint32_t box_in, uint32_t box_out, CUtensorMapSwizzle sw) {
  CUtensorMap m; uint64_t dims[2] = {inner, outer}; uint64_t str[1] = {inner * 4}; uint32_t box[2] = {box_in, box_out};
  if (umma::encode_f32(&m, base, 2, dims, str, box, sw)) { printf("fp32 tensor map encode failed\n"); exit(3); }
  return m;
}

static float* d_out; static int* d_status;

static double run(const char* name, const CUtensorMap& mA, const CUtensorMap& mB, const ProbeParams& p, const std::vector<double>& ref, size_t smem_bytes) {
  CK(cudaMemset(d_out, 0, 128 * 256 * 4)); CK(cudaMemset(d_status, 0, 4));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes + 1024));
  probe_kernel<<<1, 128, smem_bytes + 1024>>>(mA, mB, p, d_out, d_status);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-58s LAUNCH/EXEC ERROR %s\n", name, cudaGetErrorString(e)); exit(4); }
  int st; CK(cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost));
  std::vector<float> h(128 * p.N); CK(cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost));
  double maxref = 0, maxerr = 0, se = 0, sr = 0;
  for (size_t i = 0; i < h.size(); ++i) { maxref = fmax(maxref, fabs(ref[i])); maxerr = fmax(maxerr, fabs(h[i] - ref[i])); se += (h[i] - ref[i]) * (h[i] - ref[i]); sr += ref[i] * ref[i]; }
  printf("%-58s status=%d max_err/max_ref = %.3e  rel_l2 = %.3e  (d[0]=%.8g ref[0]=%.8g)\n", name, st, maxerr / (maxref + 1e-30), sqrt(se / (sr + 1e-300)), h[0], ref[0]);
  return maxerr / (maxref + 1e-30);
}

static std::vector<double> gemm_ref(const std::vector<float>& A, const std::vector<float>& B, int M, int N, int K, int a_row0 = 0) {
  std::vector<double> ref((size_t)M * N);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[(size_t)(m + a_row0) * K + k] * B[(size_t)n * K + k]; ref[(size_t)m * N + n] = s; }
  return ref;
}
static std::vector<__half> to_half(const std::vector<float>& v) { std::vector<__half> h(v.size()); for (size_t i = 0; i < v.size(); ++i) h[i] = __float2half_rn(v[i]); return h; }
static std::vector<float> half_exact(std::vector<float> v) { for (auto& x : v) x = __half2float(__float2half_rn(x)); return v; }

int main() {
  CK(cudaSetDevice(0));
  CK(cudaMalloc(&d_out, 128 * 256 * 4)); CK(cudaMalloc(&d_status, 4));
  srand(1);
  const int M = 128;

  // ---------- F1 / F2 / F3: K-major layouts, N = 96.  K elements per smem row: 64 (SW128), 32 (SW64), 16 (SW32)
  struct KL { const char* name; int krow; CUtensorMapSwizzle sw; uint32_t layout; uint32_t sbo; };
  const KL kls[] = {{"F1 K-major SW128 (64 halves / row)", 64, CU_TENSOR_MAP_SWIZZLE_128B, umma::LAYOUT_SW128, 1024},
                    {"F2 K-major SW64  (32 halves / row)", 32, CU_TENSOR_MAP_SWIZZLE_64B, umma::LAYOUT_SW64, 512},
                    {"F3 K-major SW32  (16 halves / row)", 16, CU_TENSOR_MAP_SWIZZLE_32B, umma::LAYOUT_SW32, 256}};
  for (const KL& kl : kls) {
    const int N = 96, K = kl.krow;
    std::vector<float> A = half_exact(std::vector<float>(M * K)), B = half_exact(std::vector<float>(N * K));
    for (auto& v : A) v = __half2float(__float2half_rn(frand()));
    for (auto& v : B) v = __half2float(__float2half_rn(frand()));
    std::vector<double> ref = gemm_ref(A, B, M, N, K);
    __half *dA = upload(to_half(A)), *dB = upload(to_half(B));
    CUtensorMap mA = map_h(dA, K, M, K, 128, kl.sw), mB = map_h(dB, K, N, K, 96, kl.sw);
    const uint32_t a_bytes = M * K * 2, b_off = 16384;
    ProbeParams p{}; p.f16 = 1; p.n_loads = 2; p.loads[0] = {0, 0, 0, 0, a_bytes}; p.loads[1] = {1, b_off, 0, 0, (uint32_t)N * K * 2};
    p.n_mma = K / 16; for (int k = 0; k < p.n_mma; ++k) p.mma[k] = {(uint32_t)k * 32, b_off + k * 32};
    p.a_desc = umma::make_desc_base(16, kl.sbo, kl.layout); p.b_desc = p.a_desc;
    p.idesc = f16p::make_idesc_f16(128, N, 0, 0); p.N = N;
    run(kl.name, mA, mB, p, ref, 32768);

    // ---------- F4: A staged with 136 rows, D rows = A rows [r, r + 128)  (tap shifts of the implicit GEMM)
    std::vector<float> A2(136 * K); for (auto& v : A2) v = __half2float(__float2half_rn(frand()));
    __half* dA2 = upload(to_half(A2));
    CUtensorMap mA2 = map_h(dA2, K, 136, K, 136, kl.sw);
    const uint32_t row_bytes = K * 2, b2 = 20480;
    for (int r : {8, 1, 3, 5}) for (int bo : {0, 1}) {
      std::vector<double> ref2 = gemm_ref(A2, B, M, N, K, r);
      ProbeParams q = p; q.loads[0] = {0, 0, 0, 0, 136u * row_bytes}; q.loads[1] = {1, b2, 0, 0, (uint32_t)N * K * 2};
      for (int k = 0; k < q.n_mma; ++k) q.mma[k] = {(uint32_t)(r * row_bytes + k * 32), b2 + k * 32};
      const uint32_t rows_per_atom = 8;      // every swizzle mode permutes within 8 rows
      q.a_desc = umma::make_desc_base(16, kl.sbo, kl.layout, bo ? (uint32_t)(r % rows_per_atom) : 0);
      char nm[128]; snprintf(nm, 128, "F4 %.16s row shift r=%d base_offset=%d", kl.name + 3, r, bo ? r % 8 : 0);
      run(nm, mA2, mB, q, ref2, 40960);
    }
  }

  // ---------- F2': half chunk - 16 valid channels loaded through a 32-wide SW64 box (TMA zero-fills the rest), ONE MMA per row
  {
    const int N = 96, K = 16;
    std::vector<float> A(M * K), B(N * K);
    for (auto& v : A) v = __half2float(__float2half_rn(frand()));
    for (auto& v : B) v = __half2float(__float2half_rn(frand()));
    std::vector<double> ref = gemm_ref(A, B, M, N, K);
    __half *dA = upload(to_half(A)), *dB = upload(to_half(B));
    CUtensorMap mA = map_h(dA, K, M, 32, 128, CU_TENSOR_MAP_SWIZZLE_64B), mB = map_h(dB, K, N, 32, 96, CU_TENSOR_MAP_SWIZZLE_64B);
    ProbeParams p{}; p.f16 = 1; p.n_loads = 2; p.loads[0] = {0, 0, 0, 0, (uint32_t)M * 32 * 2}; p.loads[1] = {1, 16384, 0, 0, (uint32_t)N * 32 * 2};
    p.n_mma = 1; p.mma[0] = {0, 16384};
    p.a_desc = umma::make_desc_base(16, 512, umma::LAYOUT_SW64); p.b_desc = p.a_desc;
    p.idesc = f16p::make_idesc_f16(128, N, 0, 0); p.N = N;
    run("F2' half chunk: 16 of 32 channels valid, one MMA", mA, mB, p, ref, 32768);
    p.n_mma = 2; p.mma[1] = {32, 16384 + 32};        // the second MMA must add exactly zero
    run("F2' half chunk, both MMAs (second reads the zero fill)", mA, mB, p, ref, 32768);
  }

  // ---------- F5: two-term fp16 split of fp32 data (wide dynamic range), scaled by 2^k, vs 3xTF32 on the same data
  {
    const int N = 96, K = 64;
    std::vector<float> A(M * K), B(N * K);
    for (auto& v : A) v = frand() * ldexpf(1.f, -(rand() % 12));            // activations: 12 binades
    for (auto& v : B) v = frand() * 1e-4f * ldexpf(1.f, -(rand() % 16));    // gradients: ~1e-4 and 16 binades below
    std::vector<double> ref = gemm_ref(A, B, M, N, K);
    auto max_abs = [](const std::vector<float>& v) { float m = 0; for (float x : v) m = fmaxf(m, fabsf(x)); return m; };
    for (int target : {14, 9, 6, 0}) {
      const float sa = ldexpf(1.f, target - (int)ceilf(log2f(max_abs(A)))), sb = ldexpf(1.f, target - (int)ceilf(log2f(max_abs(B))));
      std::vector<__half> Ahl(2 * M * K), Bhl(2 * N * K);
      for (int i = 0; i < M * K; ++i) { const float x = A[i] * sa; Ahl[i] = __float2half_rn(x); Ahl[M * K + i] = __float2half_rn(x - __half2float(Ahl[i])); }
      for (int i = 0; i < N * K; ++i) { const float x = B[i] * sb; Bhl[i] = __float2half_rn(x); Bhl[N * K + i] = __float2half_rn(x - __half2float(Bhl[i])); }
      __half *dA = upload(Ahl), *dB = upload(Bhl);
      CUtensorMap mA = map_h(dA, K, 2 * M, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B), mB = map_h(dB, K, 2 * N, 64, 96, CU_TENSOR_MAP_SWIZZLE_128B);
      ProbeParams p{}; p.f16 = 1; p.n_loads = 4;
      p.loads[0] = {0, 0, 0, 0, 16384}; p.loads[1] = {0, 16384, 0, M, 16384};
      p.loads[2] = {1, 32768, 0, 0, 12288}; p.loads[3] = {1, 49152, 0, N, 12288};
      p.a_desc = umma::make_desc_base(16, 1024, umma::LAYOUT_SW128); p.b_desc = p.a_desc;
      p.idesc = f16p::make_idesc_f16(128, N, 0, 0); p.N = N;
      std::vector<double> sref(ref.size()); for (size_t i = 0; i < ref.size(); ++i) sref[i] = ref[i] * (double)sa * (double)sb;
      char nm[128];
      p.n_mma = 4; for (int k = 0; k < 4; ++k) p.mma[k] = {(uint32_t)k * 32, 32768u + k * 32};
      snprintf(nm, 128, "F5 fp16 single plane, max -> 2^%d", target); run(nm, mA, mB, p, sref, 65536);
      p.n_mma = 12;
      for (int k = 0; k < 4; ++k) { p.mma[3 * k] = {16384u + k * 32, 32768u + k * 32}; p.mma[3 * k + 1] = {(uint32_t)k * 32, 49152u + k * 32}; p.mma[3 * k + 2] = {(uint32_t)k * 32, 32768u + k * 32}; }
      snprintf(nm, 128, "F5 fp16x2 (lo*hi + hi*lo + hi*hi), max -> 2^%d", target); run(nm, mA, mB, p, sref, 65536);
    }
    // the same data through today's 3xTF32 (round-to-nearest hi / lo planes, K = 64 as 2 x 32-wide SW128 rows)
    std::vector<float> Ahl(2 * M * K), Bhl(2 * N * K);
    for (int i = 0; i < M * K; ++i) { Ahl[i] = tf32_rn(A[i]); Ahl[M * K + i] = tf32_rn(A[i] - Ahl[i]); }
    for (int i = 0; i < N * K; ++i) { Bhl[i] = tf32_rn(B[i]); Bhl[N * K + i] = tf32_rn(B[i] - Bhl[i]); }
    float *dA = upload(Ahl), *dB = upload(Bhl);
    CUtensorMap mA = map_f(dA, K, 2 * M, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B), mB = map_f(dB, K, 2 * N, 32, 96, CU_TENSOR_MAP_SWIZZLE_128B);
    ProbeParams p{}; p.f16 = 0; p.n_loads = 8;
    // smem: A hi k0 | A hi k1 | A lo k0 | A lo k1 (16 KB each) | B hi k0 | B hi k1 | B lo k0 | B lo k1 (12 KB each)
    for (int pl = 0; pl < 2; ++pl) for (int kc = 0; kc < 2; ++kc) {
      p.loads[pl * 2 + kc] = {0, (uint32_t)(pl * 2 + kc) * 16384, kc * 32, pl * M, 16384};
      p.loads[4 + pl * 2 + kc] = {1, 65536u + (uint32_t)(pl * 2 + kc) * 12288, kc * 32, pl * N, 12288};
    }
    p.a_desc = umma::make_desc_base(16, 1024, umma::LAYOUT_SW128); p.b_desc = p.a_desc;
    p.idesc = umma::make_idesc_tf32(128, N, 0, 0); p.N = N;
    p.n_mma = 24;
    for (int k = 0; k < 8; ++k) {
      const uint32_t ah = (k / 4) * 16384 + (k % 4) * 32, al = 32768 + ah, bh = 65536 + (k / 4) * 12288 + (k % 4) * 32, bl = bh + 24576;
      p.mma[3 * k] = {al, bh}; p.mma[3 * k + 1] = {ah, bl}; p.mma[3 * k + 2] = {ah, bh};
    }
    run("F5 3xTF32 on the same data (reference point)", mA, mB, p, ref, 65536 + 49152);
  }

  // ---------- F6: MN-major A (stored [K][M]) and MN-major B (stored [K][N]); SW128 atoms of 64 halves x 8 K-rows
  {
    const int N = 96, K = 32;
    std::vector<float> At(K * M), B(N * K), A(M * K);
    for (auto& v : At) v = __half2float(__float2half_rn(frand()));
    for (auto& v : B) v = __half2float(__float2half_rn(frand()));
    for (int k = 0; k < K; ++k) for (int m = 0; m < M; ++m) A[m * K + k] = At[k * M + m];
    std::vector<double> ref = gemm_ref(A, B, M, N, K);
    __half *dA = upload(to_half(At)), *dB = upload(to_half(B));
    // A: 2 M-blocks of 64 halves (128 B) x 32 K-rows = 4 KB each; within a block 8 K-rows = 1024 B
    CUtensorMap mA = map_h(dA, M, K, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B), mB = map_h(dB, K, N, 32, 96, CU_TENSOR_MAP_SWIZZLE_64B);
    ProbeParams p{}; p.f16 = 1; p.n_loads = 3;
    for (int b = 0; b < 2; ++b) p.loads[b] = {0, (uint32_t)b * 4096, b * 64, 0, 4096};
    p.loads[2] = {1, 16384, 0, 0, (uint32_t)N * K * 2};
    p.n_mma = 2; for (int k = 0; k < 2; ++k) p.mma[k] = {(uint32_t)k * 2048, 16384u + k * 32};       // 16 K-rows = 2048 B per MMA
    p.b_desc = umma::make_desc_base(16, 512, umma::LAYOUT_SW64);
    p.idesc = f16p::make_idesc_f16(128, N, 1, 0); p.N = N;
    p.a_desc = umma::make_desc_base(4096, 1024, umma::LAYOUT_SW128);
    run("F6 MN-major A SW128 (LBO=4096 M-block, SBO=1024 8 K-rows)", mA, mB, p, ref, 32768);
    p.a_desc = umma::make_desc_base(1024, 4096, umma::LAYOUT_SW128);
    run("F6' MN-major A SW128 (LBO=1024, SBO=4096)", mA, mB, p, ref, 32768);

    for (int NN : {96, 48}) {
      std::vector<float> Ak(M * K), Bt(K * NN), Bk(NN * K);
      for (auto& v : Ak) v = __half2float(__float2half_rn(frand()));
      for (auto& v : Bt) v = __half2float(__float2half_rn(frand()));
      for (int k = 0; k < K; ++k) for (int n = 0; n < NN; ++n) Bk[n * K + k] = Bt[k * NN + n];
      std::vector<double> r2 = gemm_ref(Ak, Bk, M, NN, K);
      __half *dA2 = upload(to_half(Ak)), *dB2 = upload(to_half(Bt));
      CUtensorMap mA2 = map_h(dA2, K, M, 32, 128, CU_TENSOR_MAP_SWIZZLE_64B);
      // N-blocks of 64 columns; the last block runs out of bounds (zero fill): N = 96 -> 64 + 32, N = 48 -> 48 of 64
      CUtensorMap mB2 = map_h(dB2, NN, K, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B);
      const int nb = (NN + 63) / 64;
      ProbeParams q{}; q.f16 = 1; q.n_loads = 1 + nb; q.loads[0] = {0, 0, 0, 0, (uint32_t)M * K * 2};
      for (int b = 0; b < nb; ++b) q.loads[1 + b] = {1, 16384u + b * 4096, b * 64, 0, 4096};
      q.n_mma = 2; for (int k = 0; k < 2; ++k) q.mma[k] = {(uint32_t)k * 32, 16384u + k * 2048};
      q.a_desc = umma::make_desc_base(16, 512, umma::LAYOUT_SW64);
      q.idesc = f16p::make_idesc_f16(128, NN, 0, 1); q.N = NN;
      char nm[128];
      q.b_desc = umma::make_desc_base(4096, 1024, umma::LAYOUT_SW128);
      snprintf(nm, 128, "F6 MN-major B SW128 N=%d (LBO=4096, SBO=1024)", NN); run(nm, mA2, mB2, q, r2, 32768);
      q.b_desc = umma::make_desc_base(1024, 4096, umma::LAYOUT_SW128);
      snprintf(nm, 128, "F6' MN-major B SW128 N=%d (LBO=1024, SBO=4096)", NN); run(nm, mA2, mB2, q, r2, 32768);
    }
  }

  // ---------- F7: issue rate, kind::f16 vs kind::tf32, K-major SW128, one accumulator tile
  {
    long long* d; CK(cudaMalloc(&d, 8));
    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const uint64_t desc = umma::make_desc_base(16, 1024, umma::LAYOUT_SW128);
    for (int N : {48, 96, 144, 192, 256}) for (int f16 : {1, 0}) {
      const int n_mma = 4096;
      const uint32_t idesc = f16 ? f16p::make_idesc_f16(128, N, 0, 0) : umma::make_idesc_tf32(128, N, 0, 0);
      rate_kernel<<<1, 128, 170 * 1024>>>(desc, idesc, n_mma, f16, d);
      CK(cudaDeviceSynchronize());
      long long h; CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
      const double clk = (double)h / n_mma, flop = 2.0 * 128 * N * (f16 ? 16 : 8);
      printf("F7 %s M=128 N=%3d: %.1f clk/MMA -> %.0f FLOP/clk/SM (dense peak: %d)\n", f16 ? "kind::f16 " : "kind::tf32", N, clk, flop / clk, f16 ? 8192 : 4096);
    }
  }
  printf("f16 probe done\n");
  return 0;
}
