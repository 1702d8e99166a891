// Hardware probe (developer tool, not part of the product path): runs single-CTA tcgen05.mma
// kind::tf32 experiments whose operand tiles arrive by TMA, and compares against a CPU fp64
// product. It pins down the descriptor conventions the implicit-GEMM kernels rely on:
//   E1  K-major SWIZZLE_128B A and B            E2  how fp32 inputs become tf32 (truncate / round)
//   E3  3xTF32 hi/lo split accuracy             E4  MN-major A     E5/E6  MN-major B (N=96 / 48)
//   E7  K-major A whose start address is shifted by whole 128-byte rows (halo reuse)
//   E8  K-major SWIZZLE_64B (16-channel chunks)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../umma.cuh"

struct ProbeLoad { int map; uint32_t smem_off; int c0, c1; uint32_t bytes; };
struct ProbeMma { uint32_t a_off, b_off; };
struct ProbeParams {
  int n_loads; ProbeLoad loads[16];
  int n_mma;   ProbeMma mma[48];
  uint64_t a_desc, b_desc;
  uint32_t idesc; int N;
};

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2);} } while (0)

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
             const __grid_constant__ ProbeParams p, float* __restrict__ out, int* __restrict__ status) {
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
  bool ok = true;
  if (threadIdx.x == 0) {
    uint32_t total = 0;
    for (int i = 0; i < p.n_loads; ++i) total += p.loads[i].bytes;
    umma::mbar_expect_tx(bar_ld, total);
    for (int i = 0; i < p.n_loads; ++i) {
      const ProbeLoad& l = p.loads[i];
      umma::tma_load_2d(sbase + l.smem_off, l.map == 0 ? &mapA : &mapB, bar_ld, l.c0, l.c1);
    }
    ok = umma::mbar_wait(bar_ld, 0);
    if (ok) {
      umma::tc_fence_after();
      for (int i = 0; i < p.n_mma; ++i)
        umma::mma_tf32_ss(tmem, umma::desc_at(p.a_desc, sbase + p.mma[i].a_off),
                          umma::desc_at(p.b_desc, sbase + p.mma[i].b_off), p.idesc, i > 0);
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

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

struct Dev { float* p; size_t n; };
static Dev upload(const std::vector<float>& h) { Dev d; d.n = h.size(); CK(cudaMalloc(&d.p, d.n * 4)); CK(cudaMemcpy(d.p, h.data(), d.n * 4, cudaMemcpyHostToDevice)); return d; }

static CUtensorMap map2d(float* base, uint64_t inner, uint64_t outer, uint32_t box_in, uint32_t box_out, CUtensorMapSwizzle sw) {
  CUtensorMap m; uint64_t dims[2] = {inner, outer}; uint64_t str[1] = {inner * 4}; uint32_t box[2] = {box_in, box_out};
  int r = umma::encode_f32(&m, base, 2, dims, str, box, sw);
  if (r) { printf("tensor map encode failed %d\n", r); exit(3); }
  return m;
}

static float* d_out; static int* d_status;

// Runs the probe and returns max |D - ref| / max|ref| over an [M=128][N] result (ref given in double).
static double run(const char* name, const CUtensorMap& mA, const CUtensorMap& mB, const ProbeParams& p,
                  const std::vector<double>& ref, size_t smem_bytes, std::vector<float>* out_copy = nullptr) {
  CK(cudaMemset(d_out, 0, 128 * 256 * 4)); CK(cudaMemset(d_status, 0, 4));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes + 1024));
  probe_kernel<<<1, 128, smem_bytes + 1024>>>(mA, mB, p, d_out, d_status);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-46s LAUNCH/EXEC ERROR %s\n", name, cudaGetErrorString(e)); exit(4); }
  int st; CK(cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost));
  std::vector<float> h(128 * p.N); CK(cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost));
  double maxref = 0, maxerr = 0;
  for (size_t i = 0; i < h.size(); ++i) { maxref = fmax(maxref, fabs(ref[i])); maxerr = fmax(maxerr, fabs(h[i] - ref[i])); }
  printf("%-46s status=%d max_err/max_ref = %.3e  (d[0]=%.8g ref[0]=%.8g)\n", name, st, maxerr / (maxref + 1e-30), h[0], ref[0]);
  if (out_copy) *out_copy = h;
  return maxerr / (maxref + 1e-30);
}

int main() {
  CK(cudaSetDevice(0));
  CK(cudaMalloc(&d_out, 128 * 256 * 4)); CK(cudaMalloc(&d_status, 4));
  srand(1);
  const int M = 128, K = 32;

  // ---------- E1: K-major SW128, N=96
  {
    const int N = 96;
    std::vector<float> A(M * K), B(N * K);
    for (auto& v : A) v = tf32_trunc(frand());
    for (auto& v : B) v = tf32_trunc(frand());
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k]; ref[m * N + n] = s; }
    Dev dA = upload(A), dB = upload(B);
    CUtensorMap mA = map2d(dA.p, K, M, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B), mB = map2d(dB.p, K, N, 32, 96, CU_TENSOR_MAP_SWIZZLE_128B);
    ProbeParams p{}; p.n_loads = 2; p.loads[0] = {0, 0, 0, 0, (uint32_t)M * K * 4}; p.loads[1] = {1, 16384, 0, 0, (uint32_t)N * K * 4};
    p.n_mma = 4; for (int k = 0; k < 4; ++k) p.mma[k] = {(uint32_t)k * 32, 16384u + k * 32};
    p.a_desc = umma::make_desc_base(16, 1024, umma::LAYOUT_SW128); p.b_desc = p.a_desc;
    p.idesc = umma::make_idesc_tf32(128, N, 0, 0); p.N = N;
    run("E1 K-major SW128 A,B N=96", mA, mB, p, ref, 32768);

    // ---------- E7: row-shifted A start (A staged with 136 rows; D rows = A rows [r, r+128))
    std::vector<float> A2(136 * K); for (auto& v : A2) v = tf32_trunc(frand());
    Dev dA2 = upload(A2);
    CUtensorMap mA2 = map2d(dA2.p, K, 136, 32, 136, CU_TENSOR_MAP_SWIZZLE_128B);
    for (int r : {8, 1, 3}) for (int bo : {0, 1}) {
      std::vector<double> ref2(M * N);
      for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A2[(m + r) * K + k] * B[n * K + k]; ref2[m * N + n] = s; }
      ProbeParams q = p; q.loads[0] = {0, 0, 0, 0, 136u * K * 4}; q.loads[1] = {1, 20480, 0, 0, (uint32_t)N * K * 4};
      for (int k = 0; k < 4; ++k) q.mma[k] = {(uint32_t)(r * 128 + k * 32), 20480u + k * 32};
      q.a_desc = umma::make_desc_base(16, 1024, umma::LAYOUT_SW128, bo ? (r & 7) : 0);
      char nm[96]; snprintf(nm, 96, "E7 row-shift r=%d base_offset=%d", r, bo ? (r & 7) : 0);
      run(nm, mA2, mB, q, ref2, 40960);
    }
  }
  // ---------- E2: fp32 -> tf32 conversion semantics of the MMA datapath
  {
    const int N = 16;
    float a = 1.f + ldexpf(1.f, -11) + ldexpf(1.f, -12);
    std::vector<float> A(M * K, a), B(N * K, 1.f);
    std::vector<double> ref(M * N, 32.0);
    Dev dA = upload(A), dB = upload(B);
    CUtensorMap mA = map2d(dA.p, K, M, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B), mB = map2d(dB.p, K, N, 32, 16, CU_TENSOR_MAP_SWIZZLE_128B);
    ProbeParams p{}; p.n_loads = 2; p.loads[0] = {0, 0, 0, 0, (uint32_t)M * K * 4}; p.loads[1] = {1, 16384, 0, 0, (uint32_t)N * K * 4};
    p.n_mma = 4; for (int k = 0; k < 4; ++k) p.mma[k] = {(uint32_t)k * 32, 16384u + k * 32};
    p.a_desc = umma::make_desc_base(16, 1024, umma::LAYOUT_SW128); p.b_desc = p.a_desc;
    p.idesc = umma::make_idesc_tf32(128, N, 0, 0); p.N = N;
    std::vector<float> o;
    run("E2 conversion (32=truncate, 32.03125=round)", mA, mB, p, ref, 32768, &o);
    printf("   E2 value: %.8f  (truncate -> 32.0, round-to-nearest -> %.8f, exact fp32 -> %.8f)\n", o[0], 32.0 * (1 + ldexp(1., -10)), 32.0 * a);
  }
  // ---------- E3: 3xTF32 split accuracy on plain fp32 data
  {
    const int N = 96;
    std::vector<float> A(M * K), B(N * K), Ahl(2 * M * K), Bhl(2 * N * K);
    for (auto& v : A) v = frand();
    for (auto& v : B) v = frand();
    for (int i = 0; i < M * K; ++i) { Ahl[i] = tf32_trunc(A[i]); Ahl[M * K + i] = A[i] - Ahl[i]; }
    for (int i = 0; i < N * K; ++i) { Bhl[i] = tf32_trunc(B[i]); Bhl[N * K + i] = B[i] - Bhl[i]; }
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k]; ref[m * N + n] = s; }
    Dev dA = upload(Ahl), dB = upload(Bhl);
    CUtensorMap mA = map2d(dA.p, K, 2 * M, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B), mB = map2d(dB.p, K, 2 * N, 32, 96, CU_TENSOR_MAP_SWIZZLE_128B);
    ProbeParams p{}; p.n_loads = 4;
    p.loads[0] = {0, 0, 0, 0, 16384}; p.loads[1] = {0, 16384, 0, M, 16384};
    p.loads[2] = {1, 32768, 0, 0, 12288}; p.loads[3] = {1, 49152, 0, N, 12288};
    p.a_desc = umma::make_desc_base(16, 1024, umma::LAYOUT_SW128); p.b_desc = p.a_desc;
    p.idesc = umma::make_idesc_tf32(128, N, 0, 0); p.N = N;
    p.n_mma = 4; for (int k = 0; k < 4; ++k) p.mma[k] = {(uint32_t)k * 32, 32768u + k * 32};
    run("E3a single-pass tf32 (hi*hi only)", mA, mB, p, ref, 65536);
    p.n_mma = 12;
    for (int k = 0; k < 4; ++k) { p.mma[3 * k] = {16384u + k * 32, 32768u + k * 32}; p.mma[3 * k + 1] = {(uint32_t)k * 32, 49152u + k * 32}; p.mma[3 * k + 2] = {(uint32_t)k * 32, 32768u + k * 32}; }
    run("E3b 3xTF32 (lo*hi + hi*lo + hi*hi)", mA, mB, p, ref, 65536);
  }
  // ---------- E4: MN-major A (A stored [K][M]), K-major B
  {
    const int N = 96;
    std::vector<float> At(K * M), B(N * K);
    for (auto& v : At) v = tf32_trunc(frand());
    for (auto& v : B) v = tf32_trunc(frand());
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)At[k * M + m] * B[n * K + k]; ref[m * N + n] = s; }
    Dev dA = upload(At), dB = upload(B);
    // box = 32 M-elements (128 B) x 32 K-rows = 4 KB per M-block; 4 M-blocks, LBO = 4096; 8 K-rows = 1024 B (SBO)
    CUtensorMap mA = map2d(dA.p, M, K, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B), mB = map2d(dB.p, K, N, 32, 96, CU_TENSOR_MAP_SWIZZLE_128B);
    ProbeParams p{}; p.n_loads = 5;
    for (int b = 0; b < 4; ++b) p.loads[b] = {0, (uint32_t)b * 4096, b * 32, 0, 4096};
    p.loads[4] = {1, 16384, 0, 0, 12288};
    p.n_mma = 4; for (int k = 0; k < 4; ++k) p.mma[k] = {(uint32_t)k * 1024, 16384u + k * 32};
    p.a_desc = umma::make_desc_base(4096, 1024, umma::LAYOUT_SW128);
    p.b_desc = umma::make_desc_base(16, 1024, umma::LAYOUT_SW128);
    p.idesc = umma::make_idesc_tf32(128, N, 1, 0); p.N = N;
    run("E4 MN-major A SW128 (LBO=4096,SBO=1024)", mA, mB, p, ref, 32768);
    p.a_desc = umma::make_desc_base(1024, 4096, umma::LAYOUT_SW128);
    run("E4' MN-major A SW128 (LBO=1024,SBO=4096)", mA, mB, p, ref, 32768);
  }
  // ---------- E5/E6: K-major A, MN-major B (B stored [K][N]) with N = 96 and N = 48
  for (int N : {96, 48}) {
    std::vector<float> A(M * K), Bt(K * N);
    for (auto& v : A) v = tf32_trunc(frand());
    for (auto& v : Bt) v = tf32_trunc(frand());
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * Bt[k * N + n]; ref[m * N + n] = s; }
    Dev dA = upload(A), dB = upload(Bt);
    CUtensorMap mA = map2d(dA.p, K, M, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    // N-blocks of 32 columns; the last block of N=48 is loaded with a 32-wide box that runs out of bounds (zero fill)
    CUtensorMap mB = map2d(dB.p, N, K, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
    const int nb = (N + 31) / 32;
    ProbeParams p{}; p.n_loads = 1 + nb; p.loads[0] = {0, 0, 0, 0, 16384};
    for (int b = 0; b < nb; ++b) p.loads[1 + b] = {1, 16384u + b * 4096, b * 32, 0, 4096};
    p.n_mma = 4; for (int k = 0; k < 4; ++k) p.mma[k] = {(uint32_t)k * 32, 16384u + k * 1024};
    p.a_desc = umma::make_desc_base(16, 1024, umma::LAYOUT_SW128);
    p.b_desc = umma::make_desc_base(4096, 1024, umma::LAYOUT_SW128);
    p.idesc = umma::make_idesc_tf32(128, N, 0, 1); p.N = N;
    char nm[96]; snprintf(nm, 96, "E5/6 MN-major B SW128 N=%d", N);
    run(nm, mA, mB, p, ref, 32768);
  }
  // ---------- E8: K-major SW64 (16-element K chunks), N = 48
  {
    const int N = 48, K16 = 16;
    std::vector<float> A(M * K16), B(N * K16);
    for (auto& v : A) v = tf32_trunc(frand());
    for (auto& v : B) v = tf32_trunc(frand());
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K16; ++k) s += (double)A[m * K16 + k] * B[n * K16 + k]; ref[m * N + n] = s; }
    Dev dA = upload(A), dB = upload(B);
    CUtensorMap mA = map2d(dA.p, K16, M, 16, 128, CU_TENSOR_MAP_SWIZZLE_64B), mB = map2d(dB.p, K16, N, 16, 48, CU_TENSOR_MAP_SWIZZLE_64B);
    ProbeParams p{}; p.n_loads = 2; p.loads[0] = {0, 0, 0, 0, (uint32_t)M * K16 * 4}; p.loads[1] = {1, 8192, 0, 0, (uint32_t)N * K16 * 4};
    p.n_mma = 2; for (int k = 0; k < 2; ++k) p.mma[k] = {(uint32_t)k * 32, 8192u + k * 32};
    p.a_desc = umma::make_desc_base(16, 512, umma::LAYOUT_SW64); p.b_desc = p.a_desc;
    p.idesc = umma::make_idesc_tf32(128, N, 0, 0); p.N = N;
    run("E8 K-major SW64 K=16 N=48", mA, mB, p, ref, 16384);
  }
  // ---------- E9: K-major SW128 with a 16-wide inner box on a 48-channel tensor (OOB zero fill to 32)
  {
    const int N = 48, C = 48;   // A: [M][48]; chunk 1 = channels 32..63 (16 valid + 16 OOB zeros)
    std::vector<float> A(M * C), B(N * C);
    for (auto& v : A) v = tf32_trunc(frand());
    for (auto& v : B) v = tf32_trunc(frand());
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < C; ++k) s += (double)A[m * C + k] * B[n * C + k]; ref[m * N + n] = s; }
    Dev dA = upload(A), dB = upload(B);
    CUtensorMap mA = map2d(dA.p, C, M, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B), mB = map2d(dB.p, C, N, 32, 48, CU_TENSOR_MAP_SWIZZLE_128B);
    ProbeParams p{}; p.n_loads = 4;
    p.loads[0] = {0, 0, 0, 0, 16384}; p.loads[1] = {0, 16384, 32, 0, 16384};
    p.loads[2] = {1, 32768, 0, 0, 6144}; p.loads[3] = {1, 40960, 32, 0, 6144};
    p.n_mma = 8; for (int k = 0; k < 8; ++k) p.mma[k] = {(uint32_t)((k / 4) * 16384 + (k % 4) * 32), (uint32_t)(32768 + (k / 4) * 8192 + (k % 4) * 32)};
    p.a_desc = umma::make_desc_base(16, 1024, umma::LAYOUT_SW128); p.b_desc = p.a_desc;
    p.idesc = umma::make_idesc_tf32(128, N, 0, 0); p.N = N;
    run("E9 C=48 as 2x32 chunks w/ OOB zero fill", mA, mB, p, ref, 49152);
  }
  // ---------- E10: MN-major operands with the 128B swizzle whose atom is 32 bytes (tf32's only MN-major mode)
  {
    const int N = 96;
    std::vector<float> At(K * M), B(N * K);
    for (auto& v : At) v = tf32_trunc(frand());
    for (auto& v : B) v = tf32_trunc(frand());
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)At[k * M + m] * B[n * K + k]; ref[m * N + n] = s; }
    Dev dA = upload(At), dB = upload(B);
    CUtensorMap mA = map2d(dA.p, M, K, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), mB = map2d(dB.p, K, N, 32, 96, CU_TENSOR_MAP_SWIZZLE_128B);
    ProbeParams p{}; p.n_loads = 5;
    for (int b = 0; b < 4; ++b) p.loads[b] = {0, (uint32_t)b * 4096, b * 32, 0, 4096};
    p.loads[4] = {1, 16384, 0, 0, 12288};
    p.n_mma = 4; for (int k = 0; k < 4; ++k) p.mma[k] = {(uint32_t)k * 1024, 16384u + k * 32};
    p.a_desc = umma::make_desc_base(4096, 512, 1);
    p.b_desc = umma::make_desc_base(16, 1024, umma::LAYOUT_SW128);
    p.idesc = umma::make_idesc_tf32(128, N, 1, 0); p.N = N;
    run("E10 MN-major A SW128/32B (LBO=4096,SBO=512)", mA, mB, p, ref, 32768);
    p.a_desc = umma::make_desc_base(512, 4096, 1);
    run("E10' MN-major A SW128/32B (LBO=512,SBO=4096)", mA, mB, p, ref, 32768);
  }
  for (int N : {96, 48}) for (int r : {0, 1, 2, 4}) {
    const int KR = 40;   // B staged with 40 K-rows; the MMAs read rows [r, r+32)
    std::vector<float> A(M * K), Bt(KR * N);
    for (auto& v : A) v = tf32_trunc(frand());
    for (auto& v : Bt) v = tf32_trunc(frand());
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * Bt[(k + r) * N + n]; ref[m * N + n] = s; }
    Dev dA = upload(A), dB = upload(Bt);
    CUtensorMap mA = map2d(dA.p, K, M, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    CUtensorMap mB = map2d(dB.p, N, KR, 32, KR, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    const int nb = (N + 31) / 32; const uint32_t blk = KR * 128;   // 5120 B per N-block
    ProbeParams p{}; p.n_loads = 1 + nb; p.loads[0] = {0, 0, 0, 0, 16384};
    for (int b = 0; b < nb; ++b) p.loads[1 + b] = {1, 16384u + b * blk, b * 32, 0, blk};
    p.n_mma = 4; for (int k = 0; k < 4; ++k) p.mma[k] = {(uint32_t)k * 32, 16384u + (uint32_t)(r + k * 8) * 128};
    p.a_desc = umma::make_desc_base(16, 1024, umma::LAYOUT_SW128);
    p.b_desc = umma::make_desc_base(blk, 512, 1);
    p.idesc = umma::make_idesc_tf32(128, N, 0, 1); p.N = N;
    char nm[96]; snprintf(nm, 96, "E11 MN-major B SW128/32B N=%d K-row shift r=%d", N, r);
    run(nm, mA, mB, p, ref, 16384 + 3 * 5120);
  }
  printf("probe done\n");
  return 0;
}
