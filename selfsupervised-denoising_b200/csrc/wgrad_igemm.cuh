// Weight-gradient implicit GEMM on tcgen05 tensor cores (sm_100a).
//
//   dW[tap][co][ci] = sum over flat pixels j of  dZ[j][co] * X[j + off_tap][ci]
//
// Both operands are padded-flat NHWC tensors (common.cuh), so the reduction (K) dimension is the
// flat pixel index and both operands are "MN-major" for the tensor core: a smem row is one pixel,
// 64 fp16 channels (128 bytes) wide, SWIZZLE_128B atoms of 8 pixel rows (LBO = distance between 64-channel blocks,
// SBO = 1024; verified by csrc/probe/umma_f16_probe.cu F6), with start addresses shifted by whole pixel rows -
// which is how the taps of one stencil row share a single staged window of X.
//   unit  = (128-wide co tile) x (ci block, N <= 144 / 192) x (tap group = stencil row) x (K split)
//   stage = KC flat pixels of dZ (<= 2 blocks of 64 co) + KC+halo pixels of X (<= 3 blocks of 64 ci), two TMA ops
// Accumulators [co lane][tap][ci] live in TMEM for the whole K range of the unit; partial results (float4 rows, padded
// to 16 bytes) go to a workspace and a fixed-order parallel reduction produces dW (deterministic, no atomics).
// Unit order: the tap groups / ci blocks / co tiles of one K range are adjacent, so concurrently running CTAs share
// their dZ and X rows through L2; the K split gives one unit per SM (a whole wave, and the fewest partials to reduce).  Layers with <= 64 output
// channels issue M = 64 MMAs (half the dZ operand bytes).  The first conv (3 input channels) runs here too, as an N = 16 tile.
// Precision: two-term fp16 split as in the forward kernel (K = 16 pixels per MMA), the accumulator scaled back by
// 2^-(k_dz + k_x), with the same accumulator-truncation compensation per K split.
#pragma once
#include "common.cuh"
#include "umma.cuh"

struct WgradGroup { int row_off; int ntaps; int tap_rel[9]; int tap_id[9]; };

struct WgradParams {
  long long k_total;
  int KC, n_kchunks, ksplit, chunks_per_split;
  int n_co_tiles, cout, cin;
  float acc_beta;      // SSDN_ACC_BETA or 0: truncation-bias compensation of the accumulators (common.cuh)
  int m64;             // cout <= 64: MMAs of M = 64 read half the dZ bytes from shared memory (same tensor time, the kernel is
                       // operand-bandwidth-bound); accumulator row m then lives in TMEM lane (m % 16) + 32 * (m / 16)
  int cin_pitch;       // floats per (tap, co) row of the partial buffer: cin rounded up to 4 so that every row is 16-byte aligned
  int n_ci_blocks, ci_start[4], ci_n[4];
  int nba, nbx;        // 64-channel blocks actually loaded per stage for dZ (<= 2) and X (<= 3)
  const int* k_dz; const int* k_x;   // scale exponents of the two operand tensors (device; null = 0)
  int n_groups, ntaps_total;
  int rows_per_group, row_stride;   // a group holds rows_per_group stencil rows of 3 taps (consecutive pixels) each, row_stride pixels apart
  WgradGroup groups[9];
  int b_rows, stages;
  uint32_t a_plane_bytes, b_plane_bytes;   // a_plane_bytes = nba * KC * 128 (loaded part; the MMA may address up to 2 blocks)
  float* partial;     // [ksplit][ntaps][cout][cin]
  int* error_flag;
  unsigned long long* stats;   // developer instrumentation (SSDN_CONV_STATS=1)
};

struct WgradPlan {
  WgradParams p;
  double flops = 0, bytes = 0;
  CUtensorMap dz, x;   // 4-D fp16 maps (64 channels, flat pixel, channel block, plane): ONE TMA op per operand per stage
  int grid; size_t smem;
};

namespace wgradk {

constexpr int kThreads = 192;   // warp 0: TMA producer, warp 1: MMA issuer + TMEM alloc, warps 2..5: epilogue
constexpr int kMaxStages = 6;

__global__ void __launch_bounds__(kThreads, 1)
wgrad_igemm_kernel(const __grid_constant__ CUtensorMap map_dz, const __grid_constant__ CUtensorMap map_x,
                   const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 2];
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t abort_word;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = umma::smem_u32(smem);
  const uint32_t stage_bytes = 2 * (p.a_plane_bytes + p.b_plane_bytes);
  auto full = [&](int s) { return umma::smem_u32(&bars[s]); };
  auto empty = [&](int s) { return umma::smem_u32(&bars[kMaxStages + s]); };
  const uint32_t acc_full = umma::smem_u32(&bars[2 * kMaxStages]), acc_empty = umma::smem_u32(&bars[2 * kMaxStages + 1]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    abort_word = 0;
    for (int s = 0; s < p.stages; ++s) { umma::mbar_init(full(s), 1); umma::mbar_init(empty(s), 1); }
    umma::mbar_init(acc_full, 1); umma::mbar_init(acc_empty, 128);
    umma::fence_mbar_init();
  }
  if (warp == 1) { umma::tmem_alloc(umma::smem_u32(&tmem_slot), 512); umma::tmem_relinquish(); }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  const int n_units = p.n_co_tiles * p.n_ci_blocks * p.n_groups * p.ksplit;
  auto decode = [&](int u, int& ks, int& g, int& cb, int& ct) {
    // the tap groups / channel blocks / co tiles of ONE K range are neighbours in the unit order: CTAs that run at the same
    // time then read the same dZ and X rows and L2 serves the repeats (the K-range-major order re-read both from HBM once
    // per group: 1.7 GB of DRAM reads for 0.84 GB of operands, profiles/r01_ncu_summary.txt)
    g = u % p.n_groups; u /= p.n_groups; cb = u % p.n_ci_blocks; u /= p.n_ci_blocks; ct = u % p.n_co_tiles; ks = u / p.n_co_tiles;
  };
  const uint32_t abort_addr = umma::smem_u32(&abort_word);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (warp-uniform loop, elected issue)
    int stage = 0; uint32_t phase = 0;
    long long w_empty = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      int ks, g, cb, ct; decode(u, ks, g, cb, ct);
      const int c0 = ks * p.chunks_per_split, c1 = min(p.n_kchunks, c0 + p.chunks_per_split);
      for (int c = c0; c < c1; ++c) {
        SSDN_TIMED(w_empty, umma::mbar_wait(empty(stage), phase ^ 1, abort_addr, p.error_flag, 11));
        const uint32_t av = sbase + stage * stage_bytes;
        const uint32_t bv = av + 2 * p.a_plane_bytes;
        const int row = c * p.KC, brow = row + p.groups[g].row_off;
        if (umma::elect_one()) {
          // two TMA ops per stage: [plane][nba co blocks][KC rows][128 B] and [plane][nbx ci blocks][b_rows][128 B]
          umma::mbar_expect_tx(full(stage), 2 * (p.a_plane_bytes + p.b_plane_bytes));
          umma::tma_load_4d(av, &map_dz, full(stage), 0, row, ct * 2, 0);
          umma::tma_load_4d(bv, &map_x, full(stage), 0, brow, p.ci_start[cb] / 64, 0);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
    if (p.stats && lane == 0) p.stats[blockIdx.x * 16 + 0] = w_empty;
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: per (stage, tap) an unrolled block of KC/16 x 3 MMAs
    int stage = 0; uint32_t phase = 0; int it = 0;
    const uint64_t adesc = umma::make_desc_base(p.KC * 128, 1024, umma::LAYOUT_SW128);
    const uint64_t bdesc = umma::make_desc_base(p.b_rows * 128, 1024, umma::LAYOUT_SW128);
    const int nk = p.KC / 16;                       // k-steps per stage (16 pixel rows = 2048 bytes = 128 descriptor units each)
    long long w_acc = 0, w_full = 0;
    const long long t_start = clock64();
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
      int ks, g, cb, ct; decode(u, ks, g, cb, ct);
      const int c0 = ks * p.chunks_per_split, c1 = min(p.n_kchunks, c0 + p.chunks_per_split);
      const int N = p.ci_n[cb];
      const uint32_t idesc = umma::make_idesc_f16(p.m64 ? 64 : 128, N, 1, 1);
      SSDN_TIMED(w_acc, umma::mbar_wait(acc_empty, (it & 1) ^ 1, abort_addr, p.error_flag, 12));
      umma::tc_fence_after();
      uint32_t acc = 0;
      // one elected block of (taps x KC/16 k-steps x 3 products) MMAs per stage; descriptors advance by uniform adds in
      // 16-byte units (low words), so the tensor core's queue does not drain between instructions
      const uint32_t desc_hi = (uint32_t)(adesc >> 32);
      const uint32_t a_lbo = (uint32_t)(adesc & 0xffff0000u), b_lbo = (uint32_t)(bdesc & 0xffff0000u);
      const uint32_t a_pl = p.a_plane_bytes >> 4, b_pl = p.b_plane_bytes >> 4;
      const int ntaps = p.groups[g].ntaps;
      const bool unit_step = (ntaps == 3 * p.rows_per_group) && p.groups[g].tap_rel[1] == p.groups[g].tap_rel[0] + 1 && p.groups[g].tap_rel[2] == p.groups[g].tap_rel[0] + 2;
      const int rstep = p.row_stride * 8;            // descriptor units (16 bytes) between the X windows of two stencil rows
      for (int c = c0; c < c1; ++c) {
        SSDN_TIMED(w_full, umma::mbar_wait(full(stage), phase, abort_addr, p.error_flag, 12));
        umma::tc_fence_after();
        const uint32_t av = sbase + stage * stage_bytes;
        const uint32_t bv0 = av + 2 * p.a_plane_bytes;
        const uint32_t a0 = ((av >> 4) & 0x3fffu) | a_lbo;
        if (unit_step || ntaps == 1) {
          const uint32_t b0 = (((bv0 + p.groups[g].tap_rel[0] * 128) >> 4) & 0x3fffu) | b_lbo;
          if (umma::elect_one()) {
            // taps innermost: consecutive MMAs accumulate into DIFFERENT accumulators (see conv_igemm.cuh).  Two fully
            // unrolled variants - the elected thread must have next to nothing to do between two MMAs (the queue is short).
            // The A tile (dZ) of a k-step is the same for every tap: it is fetched from shared memory ONCE per plane use
            // (collector FILL on the first tap, USE on the others, LASTUSE on the last - the dz_lo tile serves one product, the
            // dz_hi tile the hi*lo and the hi*hi product back to back) instead of once per MMA.
            if (p.rows_per_group == 3) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {            // k-steps of 16 pixels (2048 bytes = 128 units); a tap shifts X by one row (8 units)
                if (k < nk) {
#pragma unroll
                  for (int prod = 0; prod < 3; ++prod) {
#pragma unroll
                    for (int r = 0; r < 3; ++r) {      // all three stencil rows of the 3x3 kernel: nine accumulators
#pragma unroll
                      for (int j = 0; j < 3; ++j) {
                        const uint32_t d = tmem + (3 * r + j) * N;
                        const uint32_t ah = a0 + k * 128, al = ah + a_pl, bh = b0 + r * rstep + j * 8 + k * 128, bl = bh + b_pl;
                        const uint32_t a_ = prod == 0 ? al : ah, b_ = prod == 1 ? bl : bh, acc_ = (k == 0 && prod == 0) ? acc : 1u;
                        const bool first = (prod == 0 || prod == 1) && r == 0 && j == 0, last = (prod == 0 || prod == 2) && r == 2 && j == 2;
                        if (first) umma::mma_f16_lo_cu<umma::CU_FILL, false>(d, a_, b_, desc_hi, idesc, acc_);
                        else if (last) umma::mma_f16_lo_cu<umma::CU_LASTUSE, false>(d, a_, b_, desc_hi, idesc, acc_);
                        else umma::mma_f16_lo_cu<umma::CU_USE, false>(d, a_, b_, desc_hi, idesc, acc_);
                      }
                    }
                  }
                }
              }
            } else if (ntaps == 3) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (k < nk) {
#pragma unroll
                  for (int prod = 0; prod < 3; ++prod) {
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                      const uint32_t d = tmem + t * N;
                      const uint32_t ah = a0 + k * 128, al = ah + a_pl, bh = b0 + t * 8 + k * 128, bl = bh + b_pl;
                      const uint32_t a_ = prod == 0 ? al : ah, b_ = prod == 1 ? bl : bh, acc_ = (k == 0 && prod == 0) ? acc : 1u;
                      const bool first = (prod == 0 || prod == 1) && t == 0, last = (prod == 0 || prod == 2) && t == 2;
                      if (first) umma::mma_f16_lo_cu<umma::CU_FILL, false>(d, a_, b_, desc_hi, idesc, acc_);
                      else if (last) umma::mma_f16_lo_cu<umma::CU_LASTUSE, false>(d, a_, b_, desc_hi, idesc, acc_);
                      else umma::mma_f16_lo_cu<umma::CU_USE, false>(d, a_, b_, desc_hi, idesc, acc_);
                    }
                  }
                }
              }
            } else {                                   // one tap (1x1 layers): the dz_hi tile still serves two products in a row
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (k < nk) {
                  const uint32_t ah = a0 + k * 128, al = ah + a_pl, bh = b0 + k * 128, bl = bh + b_pl;
                  umma::mma_f16_lo(tmem, al, bh, desc_hi, idesc, k == 0 ? acc : 1u);
                  umma::mma_f16_lo_cu<umma::CU_FILL, false>(tmem, ah, bl, desc_hi, idesc, 1u);
                  umma::mma_f16_lo_cu<umma::CU_LASTUSE, false>(tmem, ah, bh, desc_hi, idesc, 1u);
                }
              }
            }
            umma::mma_commit(empty(stage));
          }
          __syncwarp();
        } else {
          const uint32_t al = av + p.a_plane_bytes;
          for (int t = 0; t < ntaps; ++t) {
            const uint32_t d = tmem + t * N;
            const uint32_t bv = bv0 + p.groups[g].tap_rel[t] * 128, bl = bv + p.b_plane_bytes;
            if (umma::elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (k < nk) {
                  const uint32_t o = k * 16 * 128;
                  umma::mma_f16_ss(d, umma::desc_at(adesc, al + o), umma::desc_at(bdesc, bv + o), idesc, k == 0 ? acc : 1u);
                  umma::mma_f16_ss(d, umma::desc_at(adesc, av + o), umma::desc_at(bdesc, bl + o), idesc, 1);
                  umma::mma_f16_ss(d, umma::desc_at(adesc, av + o), umma::desc_at(bdesc, bv + o), idesc, 1);
                }
              }
            }
            __syncwarp();
          }
          if (umma::elect_one()) umma::mma_commit(empty(stage));
          __syncwarp();
        }
        acc = 1;
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (umma::elect_one()) umma::mma_commit(acc_full);
      __syncwarp();
    }
    if (p.stats && lane == 0) { p.stats[blockIdx.x * 16 + 1] = w_acc; p.stats[blockIdx.x * 16 + 2] = w_full; p.stats[blockIdx.x * 16 + 3] = clock64() - t_start; }
  } else {
    const int ew = warp & 3;   // TMEM sub-partition of this warp (warps 2,3,4,5 -> 2,3,0,1)
    int it = 0;
    long long w_accf = 0;
    const long long t_start = clock64();
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
      int ks, g, cb, ct; decode(u, ks, g, cb, ct);
      const int N = p.ci_n[cb];
      SSDN_TIMED(w_accf, umma::mbar_wait(acc_full, it & 1, abort_addr, p.error_flag, 13));
      umma::tc_fence_after();
      const int co = p.m64 ? (lane < 16 ? ew * 16 + lane : p.cout) : ct * 128 + ew * 32 + lane;
      // every accumulator of this unit took (K chunks of the split) x KC/16 k-steps x 3 products accumulate steps
      const int kc0 = ks * p.chunks_per_split, kc1 = min(p.n_kchunks, kc0 + p.chunks_per_split);
      const float comp = (1.0f + p.acc_beta * (float)((kc1 - kc0) * (p.KC / 16) * 3)) *
                         exp2_int(-((p.k_dz ? __ldg(p.k_dz) : 0) + (p.k_x ? __ldg(p.k_x) : 0)));
      for (int t = 0; t < p.groups[g].ntaps; ++t) {
        const int tap = p.groups[g].tap_id[t];
        float* dst = p.partial + (((long long)ks * p.ntaps_total + tap) * p.cout + co) * p.cin_pitch + p.ci_start[cb];
        for (int n0 = 0; n0 < N; n0 += 16) {
          uint32_t r[16];
          umma::tmem_ld16(tmem + (uint32_t(ew * 32) << 16) + t * N + n0, r);
          umma::tmem_ld_wait();
          if (co < p.cout) {
            if (p.ci_start[cb] + n0 + 16 <= p.cin_pitch) {      // a lane owns a contiguous, 16-byte aligned row (pad columns are ignored)
#pragma unroll
              for (int i = 0; i < 16; i += 4)
                *reinterpret_cast<float4*>(dst + n0 + i) = make_float4(comp * __uint_as_float(r[i]), comp * __uint_as_float(r[i + 1]),
                                                                      comp * __uint_as_float(r[i + 2]), comp * __uint_as_float(r[i + 3]));
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (p.ci_start[cb] + n0 + i < p.cin) dst[n0 + i] = comp * __uint_as_float(r[i]);
            }
          }
        }
      }
      umma::tc_fence_before();
      umma::mbar_arrive(acc_empty);
    }
    if (p.stats && threadIdx.x == 64) { p.stats[blockIdx.x * 16 + 4] = w_accf; p.stats[blockIdx.x * 16 + 5] = clock64() - t_start; }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem, 512);
}

// dW[co][ci][tap] (PyTorch layout) = (accumulate ? dW : 0) + sum_ks partial[ks][tap][co][ci]
// block = (32 outputs, 16 K-split lanes): thread (x, y) sums splits y, y+16, ... of output blockIdx.x*32 + x with four
// independent accumulators (coalesced 128-byte rows, many loads in flight), then the 16 lanes are combined in a fixed
// order through shared memory - deterministic, and ~10x faster than one thread walking all splits serially.
__global__ void __launch_bounds__(512) wgrad_reduce_kernel(const float* __restrict__ partial, int ksplit, int ntaps, int cout, int cin,
                                                           int pitch, float* __restrict__ dw, int accumulate) {
  __shared__ float sm[16][33];
  pdl_wait();
  const long long n = (long long)cout * pitch * ntaps;
  const long long idx = blockIdx.x * 32LL + threadIdx.x;   // over [tap][co][ci < pitch] (rows padded to `pitch` floats)
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (idx < n) {
    int k = threadIdx.y;
    for (; k + 48 < ksplit; k += 64) {
      a0 += __ldg(partial + (long long)k * n + idx); a1 += __ldg(partial + (long long)(k + 16) * n + idx);
      a2 += __ldg(partial + (long long)(k + 32) * n + idx); a3 += __ldg(partial + (long long)(k + 48) * n + idx);
    }
    for (; k < ksplit; k += 16) a0 += __ldg(partial + (long long)k * n + idx);
  }
  sm[threadIdx.y][threadIdx.x] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (threadIdx.y == 0 && idx < n) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < 16; ++w) acc += sm[w][threadIdx.x];
    const int ci = (int)(idx % pitch); long long t = idx / pitch;
    const int co = (int)(t % cout); const int tap = (int)(t / cout);
    if (ci < cin) {
      const long long o = ((long long)co * cin + ci) * ntaps + tap;
      dw[o] = accumulate ? dw[o] + acc : acc;
    }
  }
}
// All weight gradients of a network in ONE launch at the end of the backward pass (every layer keeps its own partial
// buffer): a block finds its job from the prefix sums of the jobs' block counts, then works like wgrad_reduce_kernel.
struct WgradReduceJob { const float* partial; float* dw; int ksplit, ntaps, cout, cin, pitch, first_block; };
constexpr int kMaxReduceJobs = 24;
struct WgradReduceJobs { WgradReduceJob j[kMaxReduceJobs]; int n_jobs; };
__global__ void __launch_bounds__(512) wgrad_reduce_batched_kernel(const __grid_constant__ WgradReduceJobs jobs) {
  // block = (32 lanes x 4 outputs each, 16 K-split lanes): 16-byte loads of the float4-aligned partial rows
  __shared__ float4 sm[16][33];
  pdl_wait();
  int ji = 0;
  while (ji + 1 < jobs.n_jobs && (int)blockIdx.x >= jobs.j[ji + 1].first_block) ++ji;
  const WgradReduceJob& q = jobs.j[ji];
  const long long n = (long long)q.cout * q.pitch * q.ntaps;              // a multiple of 4 (pitch is)
  const long long idx = (((long long)blockIdx.x - q.first_block) * 32LL + threadIdx.x) * 4;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
  if (idx < n) {
    int k = threadIdx.y;
    for (; k + 16 < q.ksplit; k += 32) {
      const float4 u = __ldg(reinterpret_cast<const float4*>(q.partial + (long long)k * n + idx));
      const float4 v = __ldg(reinterpret_cast<const float4*>(q.partial + (long long)(k + 16) * n + idx));
      a0.x += u.x; a0.y += u.y; a0.z += u.z; a0.w += u.w; a1.x += v.x; a1.y += v.y; a1.z += v.z; a1.w += v.w;
    }
    for (; k < q.ksplit; k += 16) {
      const float4 u = __ldg(reinterpret_cast<const float4*>(q.partial + (long long)k * n + idx));
      a0.x += u.x; a0.y += u.y; a0.z += u.z; a0.w += u.w;
    }
  }
  sm[threadIdx.y][threadIdx.x] = make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
  __syncthreads();
  if (threadIdx.y == 0 && idx < n) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int w = 0; w < 16; ++w) { const float4 t = sm[w][threadIdx.x]; acc[0] += t.x; acc[1] += t.y; acc[2] += t.z; acc[3] += t.w; }
    const int ci0 = (int)(idx % q.pitch); long long t = idx / q.pitch;       // the 4 outputs share (tap, co): pitch % 4 == 0
    const int co = (int)(t % q.cout); const int tap = (int)(t / q.cout);
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (ci0 + e < q.cin) q.dw[((long long)co * q.cin + ci0 + e) * q.ntaps + tap] = acc[e];
  }
}
static inline int wgrad_cin_pitch(int cin) { return (cin + 3) / 4 * 4; }
static inline void wgrad_reduce_launch(const float* partial, int ksplit, int ntaps, int cout, int cin, int pitch, float* dw, int accumulate,
                                       cudaStream_t st) {
  const long long n = (long long)cout * pitch * ntaps;
  SSDN_PROF(K_WGRAD_REDUCE, 0, (double)n * 4 * (ksplit + 1), st,
            (wgrad_reduce_kernel<<<(unsigned)((n + 31) / 32), dim3(32, 16), 0, st>>>(partial, ksplit, ntaps, cout, cin, pitch, dw, accumulate)));
}

}  // namespace wgradk

// ------------------------------------------------------------------------------------------ host
#include <algorithm>

static inline int wgrad_ci_cap(int ntaps) { return ntaps == 9 ? 144 : 192; }   // widest ci block (MMA N) per unit
static inline int wgrad_n_ci_blocks(int cin, int ntaps) { const int c16 = (cin + 15) / 16 * 16; return (c16 + wgrad_ci_cap(ntaps) - 1) / wgrad_ci_cap(ntaps); }

static inline size_t wgrad_partial_floats(int ksplit, int ntaps, int cout, int cin) {
  return (size_t)ksplit * ntaps * cout * ((cin + 3) / 4 * 4);
}
// pixels per pipeline stage: 64 (4 k-steps per barrier round trip) measured 10 % faster over the step than 32 (profiles/r02_layer_times.log)
static inline int wgrad_kc() { static const int kc = getenv("SSDN_WGRAD_KC") ? atoi(getenv("SSDN_WGRAD_KC")) : 64; return kc == 32 ? 32 : 64; }

// Narrow ci blocks (9 taps x N <= 512 TMEM columns, i.e. cin <= 48) keep ALL nine accumulators of a 3x3 stencil in one unit: dZ is
// then streamed once instead of once per stencil row, and the K split has three times as many units to hand out.  Needs the X
// window of three image rows (2 x pitch + 2 more pixels than the chunk) in one TMA box of <= 256 rows.
static inline bool wgrad_all_taps_in_one_unit(int cin, int ntaps, int pitch) {
  static const bool off = getenv("SSDN_WGRAD_NINE") && atoi(getenv("SSDN_WGRAD_NINE")) == 0;
  const int n = (cin + 15) / 16 * 16;
  return !off && ntaps == 9 && 9 * n <= 512 && (wgrad_kc() + 2 * pitch + 2 + 7) / 8 * 8 <= 256;
}
// Decides the K split for a layer (so that the grid fills the chip) without needing pointers.  pitch: row pitch of the geometry.
static inline int wgrad_pick_ksplit(long long k_total, int cout, int cin, int ntaps, int num_sms, int pitch) {
  const int KC = wgrad_kc();
  const int n_kchunks = (int)((k_total + KC - 1) / KC);
  const int n_co_tiles = (cout + 127) / 128, n_ci_blocks = wgrad_n_ci_blocks(cin, ntaps);
  const int n_groups = ntaps == 9 ? (wgrad_all_taps_in_one_unit(cin, ntaps, pitch) ? 1 : 3) : ntaps;
  const int others = n_co_tiles * n_ci_blocks * n_groups;
  static const int units_per_sm = getenv("SSDN_WGRAD_UNITS_PER_SM") ? atoi(getenv("SSDN_WGRAD_UNITS_PER_SM")) : 1;   // measured: 1 > 2 > 3 (fewer partials to write and reduce)
  int ks = std::max(1, (units_per_sm * num_sms) / others);   // whole waves only: one unit more would cost a whole extra wave
  // at least 4 chunks per unit on the large levels (fewer partials to write and reduce); on the small pyramid levels the serial MMA
  // chain of a unit IS the kernel's duration (4 chunks x 108 MMAs = 8 us of a 15 us kernel): one or two chunks per unit there
  const int min_chunks = n_kchunks <= 256 ? 1 : 4;
  ks = std::min(ks, std::max(1, n_kchunks / min_chunks));
  ks = std::max(1, std::min(ks, n_kchunks));
  const int cps = (n_kchunks + ks - 1) / ks;      // no K split may be empty: its accumulator would be undefined
  return (n_kchunks + cps - 1) / cps;
}

// taps: flat offsets of X relative to dZ for each weight tap (the FORWARD offsets of the conv).
// dz / x: fp16 plane pairs; k_dz / k_x: device pointers to their scale exponents (null = unscaled).
static inline int wgrad_plan_init(WgradPlan* plan, long long k_total, const __half* dz_hi, const __half* dz_lo, int dz_cpitch,
                                  int dz_coff, int cout, const __half* x_hi, const __half* x_lo, int x_cpitch, int x_coff,
                                  int cin, const ConvTaps& taps, int ksplit, float* partial, int* error_flag, int num_sms,
                                  const int* k_dz, const int* k_x) {
  WgradParams& p = plan->p;
  p = WgradParams{};
  p.k_total = k_total; p.KC = wgrad_kc(); p.n_kchunks = (int)((k_total + p.KC - 1) / p.KC);
  p.ksplit = ksplit; p.chunks_per_split = (p.n_kchunks + ksplit - 1) / ksplit;
  p.cout = cout; p.cin = cin; p.cin_pitch = (cin + 3) / 4 * 4; p.n_co_tiles = (cout + 127) / 128;
  p.k_dz = k_dz; p.k_x = k_x;
  p.acc_beta = (getenv("SSDN_ACC_COMP") && atoi(getenv("SSDN_ACC_COMP")) == 0) ? 0.0f : SSDN_ACC_BETA;
  p.m64 = (cout <= 64 && !(getenv("SSDN_WGRAD_M64") && atoi(getenv("SSDN_WGRAD_M64")) == 0)) ? 1 : 0;
  if (dz_cpitch % 8 || dz_coff % 8 || x_cpitch % 8 || x_coff % 8) return -16;     // TMA: 16-byte strides and bases
  // ci blocks: as wide as TMEM allows (3 taps x N <= 512 columns, N <= 256), so that dZ is streamed as few times as possible
  const int cap = wgrad_ci_cap(taps.n);
  p.n_ci_blocks = 0;
  const int cin16 = (cin + 15) / 16 * 16;
  const int nblocks = (cin16 + cap - 1) / cap;
  const int per = ((cin16 + nblocks - 1) / nblocks + 63) / 64 * 64;      // block starts must be multiples of 64 channels
  for (int c = 0; c < cin; c += per) {
    const int rem = std::min(per, cin - c);
    p.ci_start[p.n_ci_blocks] = c; p.ci_n[p.n_ci_blocks] = (rem + 15) / 16 * 16;
    ++p.n_ci_blocks;
  }
  p.nbx = (p.ci_n[0] + 63) / 64;
  p.nba = std::min(2, (cout + 63) / 64);
  p.ntaps_total = taps.n;
  int span = 0;
  p.rows_per_group = 1; p.row_stride = 0;
  bool nine = taps.n == 9 && wgrad_all_taps_in_one_unit(cin, taps.n, taps.off[3] - taps.off[0]) && taps.off[6] - taps.off[3] == taps.off[3] - taps.off[0] &&
              taps.off[3] - taps.off[0] > 0;
  for (int r = 0; r < 3 && nine; ++r) nine = taps.off[3 * r + 1] == taps.off[3 * r] + 1 && taps.off[3 * r + 2] == taps.off[3 * r] + 2;
  if (nine) {
    p.n_groups = 1; p.rows_per_group = 3; p.row_stride = taps.off[3] - taps.off[0];
    p.groups[0].row_off = taps.off[0]; p.groups[0].ntaps = 9;
    for (int t = 0; t < 9; ++t) { p.groups[0].tap_id[t] = t; p.groups[0].tap_rel[t] = taps.off[t] - taps.off[0]; span = std::max(span, taps.off[t] - taps.off[0]); }
  } else if (taps.n == 9) {
    p.n_groups = 3;
    for (int g = 0; g < 3; ++g) {
      int lo = std::min({taps.off[3 * g], taps.off[3 * g + 1], taps.off[3 * g + 2]});
      p.groups[g].row_off = lo; p.groups[g].ntaps = 3;
      for (int k = 0; k < 3; ++k) { p.groups[g].tap_id[k] = 3 * g + k; p.groups[g].tap_rel[k] = taps.off[3 * g + k] - lo; span = std::max(span, taps.off[3 * g + k] - lo); }
    }
  } else {
    p.n_groups = taps.n;
    for (int g = 0; g < taps.n; ++g) { p.groups[g].row_off = taps.off[g]; p.groups[g].ntaps = 1; p.groups[g].tap_id[0] = g; p.groups[g].tap_rel[0] = 0; }
  }
  p.b_rows = (p.KC + span + 7) / 8 * 8;
  p.a_plane_bytes = p.nba * p.KC * 128;                    // an M = 128 MMA addresses 2 co blocks; a missing one aliases whatever follows (ignored lanes)
  p.b_plane_bytes = (uint32_t)(p.nbx * p.b_rows * 128);    // b_rows is a multiple of 8 => 1024-byte multiple
  const uint32_t stage_bytes = 2 * (p.a_plane_bytes + p.b_plane_bytes);
  const size_t slack = 2 * p.KC * 128 + 1024;              // so that aliased co blocks stay inside the allocation
  static const int smem_kb = getenv("SSDN_WGRAD_SMEM_KB") ? atoi(getenv("SSDN_WGRAD_SMEM_KB")) : 224;
  p.stages = std::max(2, std::min(wgradk::kMaxStages, (int)((smem_kb * 1024 - slack) / stage_bytes)));
  plan->smem = (size_t)p.stages * stage_bytes + slack;
  if (plan->smem > 226 * 1024) return -10;
  p.partial = partial; p.error_flag = error_flag;
  plan->grid = std::min(p.n_co_tiles * p.n_ci_blocks * p.n_groups * p.ksplit, num_sms);
  // 4-D maps: (64 channels of a block, flat pixel, channel block [stride 128 B], plane).  The block dimension has a
  // smaller stride than the pixel dimension; cuTensorMapEncodeTiled accepts that (profiles/r01_tma_map_test.log) and the
  // box then lands in shared memory as [plane][block][pixel][64 ch], exactly the MN-major operand layout.  A block that
  // runs past the tensor's channels reads the next pixel's first channels: those rows / columns of the MMA are never used.
  const long long dz_plane = (long long)((const char*)dz_lo - (const char*)dz_hi), x_plane = (long long)((const char*)x_lo - (const char*)x_hi);
  if (dz_plane <= 0 || x_plane <= 0 || dz_plane % 16 || x_plane % 16) return -11;
  uint64_t d1[4] = {64, (uint64_t)k_total, (uint64_t)((cout + 63) / 64), 2};
  uint64_t s1[3] = {(uint64_t)dz_cpitch * 2, 128, (uint64_t)dz_plane};
  uint32_t b1[4] = {64, (uint32_t)p.KC, (uint32_t)p.nba, 2};
  uint64_t d2[4] = {64, (uint64_t)k_total, (uint64_t)((cin + 63) / 64), 2};
  uint64_t s2[3] = {(uint64_t)x_cpitch * 2, 128, (uint64_t)x_plane};
  uint32_t b2[4] = {64, (uint32_t)p.b_rows, (uint32_t)p.nbx, 2};
  int r;
  if ((r = umma::encode_f16(&plan->dz, (void*)(dz_hi + dz_coff), 4, d1, s1, b1, CU_TENSOR_MAP_SWIZZLE_128B))) return r;
  if ((r = umma::encode_f16(&plan->x, (void*)(x_hi + x_coff), 4, d2, s2, b2, CU_TENSOR_MAP_SWIZZLE_128B))) return r;
  return 0;
}

static inline cudaError_t wgrad_launch(const WgradPlan& plan, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgradk::wgrad_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  static const bool want_stats = getenv("SSDN_CONV_STATS") != nullptr;
  if (want_stats && profiler().on) {
    static unsigned long long* dev = nullptr;
    if (!dev) cudaMalloc(&dev, 1024 * 16 * sizeof(unsigned long long));
    cudaMemsetAsync(dev, 0, 1024 * 16 * sizeof(unsigned long long), stream);
    WgradParams p = plan.p; p.stats = dev;
    profiler().begin(K_WGRAD, plan.flops, plan.bytes, stream);
    launch_pdl(wgradk::wgrad_igemm_kernel, dim3(plan.grid), dim3(wgradk::kThreads), plan.smem, stream, plan.dz, plan.x, p);
    profiler().end(stream);
    std::vector<unsigned long long> h((size_t)plan.grid * 16);
    cudaMemcpyAsync(h.data(), dev, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream);
    cudaError_t e = cudaStreamSynchronize(stream);
    double s[8] = {0};
    for (int b = 0; b < plan.grid; ++b) for (int k = 0; k < 8; ++k) s[k] += (double)h[(size_t)b * 16 + k] / plan.grid;
    fprintf(stderr, "[wgrad stats] grid %d cout %d cin %d ci_blocks %d groups %d ksplit %d chunks/split %d stages %d nba %d nbx %d b_rows %d | mma loop %.0f clk: wait acc_empty %.1f%% "
            "full %.1f%% | producer wait empty %.1f%% | epi loop %.0f clk: wait acc_full %.1f%%\n", plan.grid, p.cout, p.cin, p.n_ci_blocks, p.n_groups, p.ksplit,
            p.chunks_per_split, p.stages, p.nba, p.nbx, p.b_rows, s[3], 100 * s[1] / s[3], 100 * s[2] / s[3], 100 * s[0] / s[3], s[5], 100 * s[4] / s[5]);
    return e != cudaSuccess ? e : cudaGetLastError();
  }
  if (profiler().on) profiler().begin(K_WGRAD, plan.flops, plan.bytes, stream);
  cudaError_t e = launch_pdl(wgradk::wgrad_igemm_kernel, dim3(plan.grid), dim3(wgradk::kThreads), plan.smem, stream, plan.dz, plan.x, plan.p);
  if (profiler().on) profiler().end(stream);
  return e != cudaSuccess ? e : cudaGetLastError();
}
