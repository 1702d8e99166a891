// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), forward and data-gradient.
//
//   D[flat pixel j, co] = sum_taps sum_ci  A[j + off_tap, ci] * Wt[tap][co][ci]
//
// A is a padded-flat NHWC activation (common.cuh); because zero padding is stored in the halo,
// every tap is a constant row offset into A.  Work decomposition:
//   unit   = T consecutive tiles of 128 flat pixels  x  one N-tile (<= 192 output channels)
//   chunk  = 32 input channels (one SWIZZLE_64B K-major smem tile row = 64 bytes of fp16, 2 k-steps of K = 16); 1x1 layers
//            use 64-channel chunks (SWIZZLE_128B, 4 k-steps): without tap reuse the TMA row rate is what binds there.
//            A channel count that is not a multiple of 32 ends in a partial chunk: TMA zero-fills the missing channels
//            and a chunk with <= 16 real channels issues only the first of its two k-steps
//   group  = a set of taps that share one staged window of A rows (halo reuse: the window is
//            loaded ONCE by TMA and each tap is only a different UMMA start address)
//   B tile = the [N][32] weight slab of one (chunk, tap), streamed through its own smem ring (3 slabs per stage).
// Warp roles (16 warps): 0 = A producer (TMA), 1 = B producer (TMA), 2 = MMA issuer (one elected thread),
// 3 = TMEM allocator, 4..15 = epilogue - three warps per TMEM lane quadrant; the (tile, 32-channel slice) work items of a CTA's
// units are dealt to them round-robin: TMEM -> registers -> one FFMA (accumulator scale 2^-(k_a + k_b), truncation compensation
// and the destination's scale; bias) -> LeakyReLU (+ sign-mask word out) or LeakyReLU' from the sign-mask word -> transposed
// through shared memory -> optional column sums of the staged tile (bias gradient) -> fp16 hi/lo split (+ running max|v| for
// the destination's next scale) -> 16-byte stores with the upsample / un-rotate / NCHW scatter in the address.
// Accumulators are double buffered in TMEM so the epilogue of unit i overlaps the MMAs of unit i+1.
// PAIR = true: the kernel runs as clusters of two CTAs that issue cta_group::2 MMAs of M = 256 (umma.cuh): each CTA loads
// its own pixel rows of A and half of the N rows of B, the leader issues, both epilogues drain their own TMEM.
// Precision: two-term fp16 split with per-tensor power-of-two scales (see common.cuh) => fp32-grade results, 3 MMAs of
// kind::f16 per (tile, k-step of 16 channels), issued back to back: lo*hi, then hi*lo and hi*hi, which share the A_hi tile
// through the tensor core's operand collector (umma::mma_f16_lo_cu).
#pragma once
#include "common.cuh"
#include "umma.cuh"

struct TapGroup {
  int row_off;        // window start relative to the unit's first flat pixel
  int ntaps;
  int tap_rel[9];     // tap row offset inside the window (>= 0)
  int tap_id[9];      // index of the tap in the weight slab
};

struct ConvParams {
  Geom src;           // geometry of A == geometry in which output pixels are enumerated
  int T, N;           // tiles per unit, MMA N
  int n_units_m, n_tiles_n;
  int pair;           // 1: CTA pairs (cta_group::2, M = 256); n_pairs = ceil(n_units_m / 2) * n_tiles_n pair-units
  int n_pairs;
  int n_chunks, ksteps_last;
  int n_groups, ntaps_total;
  TapGroup groups[9];
  int nbox, box_rows; // every group window is loaded as nbox TMA boxes of box_rows rows (one op for both planes if nbox == 1)
  int a_stages, b_stages;
  int lookahead;      // row3 issue path: poll the next operand stages while the first tap of a stencil row is queued (see the MMA issuer)
  int bg;             // weight slabs ((chunk, tap) pairs, in consumption order) per B stage: ONE TMA op loads bg x 2 planes
  int wide;           // 1x1 convolutions with cin % 64 == 0: chunks of 64 channels (128-byte rows, SWIZZLE_128B, 4 k-steps), one
                      //    weight slab per B stage; A is not reused by other taps there, so its TMA rate (rows/clk) is what binds
  int row3;           // 1: every B stage holds the 3 taps of one stencil row of one group, tap_rel advancing by tap_step
  int tap_step;       //    (+1 forward, -1 data-gradient): the MMA warp then issues 3 x T x 6 MMAs per barrier wait
  uint32_t a_plane_bytes, b_stage_bytes;   // smem bytes of one A plane of one stage / of one B stage
  uint32_t epi_off;   // byte offset of the epilogue staging area (8 warps x 32 pixels x 36 floats) in dynamic smem
  float acc_comp;     // 1 + SSDN_ACC_BETA x (MMA instructions accumulated into one output): truncation-bias compensation
  const int* k_a; const int* k_b;   // scale exponents of the A tensor and of the weight slab (device; null = 0)
  uint64_t magic_s, magic_p;   // ceil(2^64 / src.S), ceil(2^64 / src.P): the epilogue's pixel -> (image, row, column) divisions
  uint32_t wait_hint;          // suspend-time hint (ns) of the producers' / epilogue's mbarrier waits (0 = plain try_wait spin)
  int epi_per_quad;   // epilogue warps per TMEM lane quadrant that work (1 .. kEpiPerQuad); the others exit at once.  3: every layer
                      //    but the NCHW output (a lane walks its pixel's channels there); 2 / 1: ablations (fewer staging rows buy a weight
                      //    stage: measured no gain)
  ConvDst dst;
  int* error_flag;
  int debug;          // experiments only: 2 = skip MMA issue (generic issue path)
  unsigned long long* stats;   // developer instrumentation (SSDN_CONV_STATS=1): per-CTA clocks spent in each role's waits
};

struct ConvPlan {
  ConvParams p;
  double flops = 0;   // algorithmic 2*MAC of this launch (valid pixels, real channels); filled by the owner
  double bytes = 0;   // algorithmic HBM bytes of this launch (operand planes in, result out); filled by the owner
  CUtensorMap a, b;   // a: [plane][flat pixel][channel] (3-D), b: weight slabs [slab x plane][N][32] (3-D), fp16
  int grid; size_t smem;
};

namespace convk {

#ifndef SSDN_EPI_WARPS_PER_QUADRANT
#define SSDN_EPI_WARPS_PER_QUADRANT 3
#endif
constexpr int kEpiPerQuad = SSDN_EPI_WARPS_PER_QUADRANT;   // epilogue warps per TMEM lane quadrant: warp j of a quadrant takes the 32-channel
constexpr int kEpiWarps = 4 * kEpiPerQuad;                 // slices s with s % kEpiPerQuad == j (N = 96: one slice per warp and tile).  The
                                                           // epilogue is latency-bound per warp, so more warps = more slices in flight
constexpr int kThreads = 128 + 32 * kEpiWarps;
constexpr int kMaxStages = 8;

struct Ring {
  int stage = 0; uint32_t phase = 0; int n;
  __device__ explicit Ring(int n_) : n(n_) {}
  __device__ void advance() { if (++stage == n) { stage = 0; phase ^= 1; } }
};

constexpr int kStagePitch = 36;   // floats per staged pixel row (32 channels + 4 pad: conflict-free float4 access)
constexpr int kMaxSlices = 6;     // 32-channel slices of one N tile (N <= 192)

// What the epilogue needs to know about one lane's pixel of one 128-pixel tile, for one work item (tile, 32-channel slice).
// Computed one item ahead so that the LeakyReLU sign-mask word (the only global LOAD of the epilogue) is in flight while the
// previous item is written.
struct LanePixel {
  uint32_t pk;       // destination flat pixel (first of 1 or 4) in bits [0,26) | number of destinations (0 = halo / out of range:
                     // nothing is written, 1, or 4) << 26 | write zeros (row shifted in by Shift2d) << 29 | rotation branch << 30
  int b, y, x;       // image / row / column of the source pixel (MAP_NCHW)
  uint32_t mw;       // sign-mask word of the slice (EP_ACT_GRAD)
};
constexpr uint32_t kPkPixelMask = (1u << 26) - 1;

// n / d for n < 2^32 with m = ceil(2^64 / d) (host: div_magic_for): two wide multiplies instead of the ~70-instruction
// 64-bit division subroutine the epilogue used to call twice per tile
__device__ __forceinline__ uint32_t div_magic(uint32_t n, uint64_t m) { return (uint32_t)__umul64hi((uint64_t)n, m); }

__device__ __forceinline__ void sts128(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ void lane_pixel(const ConvParams& p, int um, int nt, int tile, int s, int T, int ew, int lane, LanePixel& o) {
  const ConvDst& d = p.dst;
  const Geom& sg = p.src;
  const uint32_t j = (uint32_t)um * (128u * T) + tile * 128 + ew * 32 + lane;
  const uint32_t b = div_magic(j, p.magic_s);
  const uint32_t rem = j - b * (uint32_t)sg.S;
  const uint32_t rr = div_magic(rem, p.magic_p);
  const int x = (int)(rem - rr * (uint32_t)sg.P), y = (int)rr - sg.row0;
  const bool valid = ((int)b < sg.B) && (y >= 0) && (x < sg.W);
  o.b = (int)b; o.y = y; o.x = x; o.mw = 0xffffffffu;
  uint32_t d0 = 0, nd = 0, zero = 0, br = 0;
  if (valid) {
    const Geom& dg = d.g;
    if (d.map == MAP_IDENT) { d0 = b * dg.S + (y + dg.row0) * dg.P + x; nd = 1; }
    else if (d.map == MAP_UP2) { d0 = b * dg.S + (2 * y + dg.row0) * dg.P + 2 * x; nd = 4; }
    else if (d.map == MAP_UNROT) {
      const int H = sg.H, W = sg.W;
      br = ((int)b >= d.nimg) + ((int)b >= 2 * d.nimg) + ((int)b >= 3 * d.nimg);
      const int n = (int)b - (int)br * d.nimg;
      const int pp = (y + 1 == H) ? 0 : y + 1, q = x;
      zero = (y + 1 == H);
      int i, jj;
      if (br == 0) { i = pp; jj = q; } else if (br == 1) { i = q; jj = H - 1 - pp; }
      else if (br == 2) { i = H - 1 - pp; jj = W - 1 - q; } else { i = W - 1 - q; jj = pp; }
      d0 = n * dg.S + (i + dg.row0) * dg.P + jj; nd = 1;
    } else if (d.map == MAP_UNROT_INV) {
      const int brn = nt, H = sg.H, W = sg.W;
      int pp, q;
      if (brn == 0) { pp = y; q = x; } else if (brn == 1) { pp = H - 1 - x; q = y; }
      else if (brn == 2) { pp = H - 1 - y; q = W - 1 - x; } else { pp = x; q = W - 1 - y; }
      if (pp > 0) { d0 = (brn * d.nimg + b) * dg.S + (pp - 1 + dg.row0) * dg.P + q; nd = 1; }
    } else { nd = 1; }   // MAP_NCHW
  }
  o.pk = d0 | (nd << 26) | (zero << 29) | (br << 30);
  if ((d.flags & EP_ACT_GRAD) && nd) {
    // word of slice s: channel (first channel of the slice) / 32, in the mask row of the destination (or source) pixel
    const long long row = (d.flags & EP_ACT_AT_SRC) ? (long long)j : (long long)d0;
    const int c_first = ((d.flags & EP_ACT_AT_SRC) || d.map != MAP_UNROT_INV) ? nt * p.N : 0;
    if (c_first + 32 * s < d.mask_in_words * 32) o.mw = __ldg(d.mask_in + row * d.mask_in_words + (c_first >> 5) + s);
  }
}

template <int T, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[4 * kMaxStages + 4];
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t abort_word;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = umma::smem_u32(smem);
  const uint32_t a_stage_bytes = 2 * p.a_plane_bytes, b_stage_bytes = p.b_stage_bytes;
  const uint32_t a_base = sbase, b_base = sbase + p.a_stages * a_stage_bytes;
  const uint32_t bar0 = umma::smem_u32(&bars[0]);
  auto full_a = [&](int s) { return bar0 + 8 * s; };
  auto empty_a = [&](int s) { return bar0 + 8 * (kMaxStages + s); };
  auto full_b = [&](int s) { return bar0 + 8 * (2 * kMaxStages + s); };
  auto empty_b = [&](int s) { return bar0 + 8 * (3 * kMaxStages + s); };
  auto tmem_full = [&](int b) { return bar0 + 8 * (4 * kMaxStages + b); };
  auto tmem_empty = [&](int b) { return bar0 + 8 * (4 * kMaxStages + 2 + b); };
  const uint32_t abort_addr = umma::smem_u32(&abort_word);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Work decomposition.  Single CTA: unit u = (M unit u / n_tiles_n, N tile u % n_tiles_n), CTAs stride over units.
  // CTA pair: pair-unit q covers M units 2 * (q / n_tiles_n) + {0, 1} (rank 0 / 1) of N tile q % n_tiles_n, clusters
  // stride over pair-units; both CTAs of a pair run the same loops in lockstep through the shared barriers.
  const uint32_t rank = PAIR ? umma::cluster_ctarank() : 0;
  const bool leader = rank == 0;
  const int n_units = PAIR ? p.n_pairs : p.n_units_m * p.n_tiles_n;
  const int u_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, u_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto unit_m = [&](int u) { return PAIR ? 2 * (u / p.n_tiles_n) + (int)rank : u / p.n_tiles_n; };

  if (threadIdx.x == 0) {
    abort_word = 0;
    const int nprod = PAIR ? 2 : 1;     // a pair's leader barrier collects one arrival (+ bytes) from each CTA's producer
    for (int s = 0; s < p.a_stages; ++s) { umma::mbar_init(full_a(s), nprod); umma::mbar_init(empty_a(s), 1); }
    for (int s = 0; s < p.b_stages; ++s) { umma::mbar_init(full_b(s), nprod); umma::mbar_init(empty_b(s), 1); }
    for (int b = 0; b < 2; ++b) { umma::mbar_init(tmem_full(b), 1); umma::mbar_init(tmem_empty(b), nprod * 128 * p.epi_per_quad); }
    umma::fence_mbar_init();
  }
  if (warp == 3) {
    if (PAIR) { umma::tmem_alloc_pair(umma::smem_u32(&tmem_slot), 512); umma::tmem_relinquish_pair(); }
    else { umma::tmem_alloc(umma::smem_u32(&tmem_slot), 512); umma::tmem_relinquish(); }
  }
  if (warp == 0 && lane == 0) umma::tma_prefetch_desc(&map_a);
  if (warp == 1 && lane == 0) umma::tma_prefetch_desc(&map_b);
  umma::tc_fence_before();
  __syncthreads();
  if (PAIR) umma::cluster_sync();       // the peer's barriers must be initialised before anything arrives on them
  umma::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_launch_dependents();      // the next kernel on the stream may start its prologue on idle SMs ...
  pdl_wait();                   // ... and this one touches global memory only after its predecessor has completed

  if (warp == 0) {
    // ------------------------------------------------------------ A producer (warp-uniform loop, elected issue)
    Ring ra(p.a_stages);
    const int cw_ch = p.wide ? 64 : 32;                       // channels per chunk
    const uint32_t box_bytes = p.box_rows * cw_ch * 2;
    long long w_empty = 0;
    for (int u = u_first; u < n_units; u += u_stride) {
      const int um = unit_m(u);
      const int j0 = um * 128 * T;
      for (int ch = 0; ch < p.n_chunks; ++ch)
        for (int g = 0; g < p.n_groups; ++g) {
          SSDN_TIMED(w_empty, umma::mbar_wait(empty_a(ra.stage), ra.phase ^ 1, abort_addr, p.error_flag, 1, p.wait_hint));
          const uint32_t dst = a_base + ra.stage * a_stage_bytes;
          const int row = j0 + p.groups[g].row_off;
          if (umma::elect_one()) {
            if (PAIR) {             // own rows into own shared memory, bytes signalled on the leader's barrier
              const uint32_t lbar = umma::mapa(full_a(ra.stage), 0);
              umma::mbar_expect_tx_cluster(lbar, 2 * p.nbox * box_bytes);
              if (p.nbox == 1) {
                umma::tma_load_3d_pair(dst, &map_a, lbar, ch * cw_ch, row, 0);
              } else {
                for (int bx = 0; bx < p.nbox; ++bx) {
                  umma::tma_load_3d_pair(dst + bx * box_bytes, &map_a, lbar, ch * cw_ch, row + bx * p.box_rows, 0);
                  umma::tma_load_3d_pair(dst + p.a_plane_bytes + bx * box_bytes, &map_a, lbar, ch * cw_ch, row + bx * p.box_rows, 1);
                }
              }
            } else {
              umma::mbar_expect_tx(full_a(ra.stage), 2 * p.nbox * box_bytes);
              if (p.nbox == 1) {      // one op brings both planes: box (32 ch, rows, 2 planes)
                umma::tma_load_3d(dst, &map_a, full_a(ra.stage), ch * cw_ch, row, 0);
              } else {
                for (int bx = 0; bx < p.nbox; ++bx) {
                  umma::tma_load_3d(dst + bx * box_bytes, &map_a, full_a(ra.stage), ch * cw_ch, row + bx * p.box_rows, 0);
                  umma::tma_load_3d(dst + p.a_plane_bytes + bx * box_bytes, &map_a, full_a(ra.stage), ch * cw_ch, row + bx * p.box_rows, 1);
                }
              }
            }
          }
          __syncwarp();
          ra.advance();
        }
    }
    if (p.stats && lane == 0) p.stats[blockIdx.x * 16 + 0] = w_empty;
  } else if (warp == 1) {
    // ------------------------------------------------------------ B producer: one TMA per stage = bg slabs x 2 planes
    Ring rb(p.b_stages);
    const int n_slabs = p.n_chunks * p.ntaps_total, n_bstages = n_slabs / p.bg;
    long long w_empty = 0;
    for (int u = u_first; u < n_units; u += u_stride) {
      const int nt = u % p.n_tiles_n;
      for (int i = 0; i < n_bstages; ++i) {
        SSDN_TIMED(w_empty, umma::mbar_wait(empty_b(rb.stage), rb.phase ^ 1, abort_addr, p.error_flag, 2, p.wait_hint));
        if (umma::elect_one()) {
          if (PAIR) {               // this CTA's half of the N rows of every slab (b_stage_bytes is the per-CTA size)
            const uint32_t lbar = umma::mapa(full_b(rb.stage), 0);
            umma::mbar_expect_tx_cluster(lbar, b_stage_bytes);
            umma::tma_load_3d_pair(b_base + rb.stage * b_stage_bytes, &map_b, lbar, 0, (int)rank * (p.N / 2), 2 * (nt * n_slabs + i * p.bg));
          } else {
            umma::mbar_expect_tx(full_b(rb.stage), b_stage_bytes);
            umma::tma_load_3d(b_base + rb.stage * b_stage_bytes, &map_b, full_b(rb.stage), 0, 0, 2 * (nt * n_slabs + i * p.bg));
          }
        }
        __syncwarp();
        rb.advance();
      }
    }
    if (p.stats && lane == 0) p.stats[blockIdx.x * 16 + 1] = w_empty;
  } else if (warp == 2 && (!PAIR || leader)) {
    // ------------------------------------------------------------ MMA issuer: uniform loops, one elected lane issues a fully
    // unrolled block of T x 2 x 3 MMAs per (chunk, tap) whose descriptors are base + compile-time offsets
    Ring ra(p.a_stages), rb(p.b_stages);
    constexpr uint64_t desc = umma::make_desc_base(16, 512, umma::LAYOUT_SW64);
    const uint32_t idesc = umma::make_idesc_f16(PAIR ? 256 : 128, p.N, 0, 0);
    // single CTA: tcgen05.mma.cta_group::1 and a local commit; pair leader: cta_group::2 and a commit multicast to both CTAs
    auto mma = [&](uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t acc) {
      if (PAIR) umma::mma_f16_lo_pair(d, a_lo, b_lo, hi, idesc, acc); else umma::mma_f16_lo(d, a_lo, b_lo, hi, idesc, acc);
    };
    auto mma_fill = [&](uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi) { umma::mma_f16_lo_cu<umma::CU_FILL, PAIR>(d, a_lo, b_lo, hi, idesc, 1u); };
    auto mma_last = [&](uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi) { umma::mma_f16_lo_cu<umma::CU_LASTUSE, PAIR>(d, a_lo, b_lo, hi, idesc, 1u); };
    auto commit = [&](uint32_t bar) { if (PAIR) umma::mma_commit_pair(bar); else umma::mma_commit(bar); };
    int it = 0;
    long long w_tmem = 0, w_a = 0, w_b = 0;
    const long long t_start = clock64();
    if (p.wide) {
      // 1x1 convolutions: per 64-channel chunk one A stage and one B stage, T tiles x (4 k-steps x 3 products) MMAs
      constexpr uint64_t wdesc = umma::make_desc_base(16, 1024, umma::LAYOUT_SW128);
      const uint32_t desc_hi = (uint32_t)(wdesc >> 32);
      const uint32_t lbo_bits = (uint32_t)(wdesc & 0xffff0000u);
      const uint32_t a_pl = p.a_plane_bytes >> 4, b_pl = (uint32_t)((PAIR ? p.N / 2 : p.N) * 128) >> 4;
      for (int u = u_first; u < n_units; u += u_stride, ++it) {
        const int buf = it & 1;
        SSDN_TIMED(w_tmem, umma::mbar_wait(tmem_empty(buf), ((it >> 1) & 1) ^ 1, abort_addr, p.error_flag, 3));
        umma::tc_fence_after();
        const uint32_t d0 = tmem + buf * T * p.N;
        uint32_t first = 0;
        for (int ch = 0; ch < p.n_chunks; ++ch) {
          SSDN_TIMED(w_a, umma::mbar_wait(full_a(ra.stage), ra.phase, abort_addr, p.error_flag, 3));
          SSDN_TIMED(w_b, umma::mbar_wait(full_b(rb.stage), rb.phase, abort_addr, p.error_flag, 3));
          umma::tc_fence_after();
          const uint32_t a0 = (((a_base + ra.stage * a_stage_bytes) >> 4) & 0x3fffu) | lbo_bits;
          const uint32_t b0 = (((b_base + rb.stage * b_stage_bytes) >> 4) & 0x3fffu) | lbo_bits;
          if (umma::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
              for (int tile = 0; tile < T; ++tile) {
                // the three products of one (tile, k-step) back to back: the hi*lo and hi*hi products share the A_hi tile,
                // which the second one takes from the tensor core's collector instead of shared memory (umma::mma_f16_lo_cu)
                const uint32_t d = d0 + tile * p.N;
                const uint32_t av = a0 + tile * 1024 + 2 * k, al = av + a_pl, bv = b0 + 2 * k, bl = bv + b_pl;
                mma(d, al, bv, desc_hi, k == 0 ? first : 1u);
                mma_fill(d, av, bl, desc_hi);
                mma_last(d, av, bv, desc_hi);
              }
            }
            commit(empty_b(rb.stage));
            commit(empty_a(ra.stage));
          }
          __syncwarp();
          first = 1;
          rb.advance(); ra.advance();
        }
        if (umma::elect_one()) commit(tmem_full(buf));
        __syncwarp();
      }
    } else if (p.row3) {
      // 3x3 stencils: one barrier wait and one block of 3 taps x T tiles x (2 k-steps x 3 products) MMAs per B stage.
      // Descriptors are base + multiples of uniform strides in 16-byte units, so the elected thread has almost nothing
      // to do between two tcgen05.mma: the tensor core's short queue never drains (profiles/r01_role_waits.log).
      const uint32_t desc_hi = (uint32_t)(desc >> 32);
      const uint32_t lbo_bits = (uint32_t)(desc & 0xffff0000u);
      const uint32_t a_pl = p.a_plane_bytes >> 4, b_pl = (uint32_t)((PAIR ? p.N / 2 : p.N) * 64) >> 4, slab = 2 * b_pl;
      const int step = p.tap_step * 4;                       // one pixel row of the A window = 64 bytes
      // The issuing thread's waits are NOT free: the tensor core's queue is a few instructions deep, so the ~200 clk an mbarrier
      // wait + fence + elect costs - even on a barrier that completed long ago - idles the pipe (the loop ran 17 % above the
      // tensor floor, the sum of its full_a / full_b wait shares).  Each stage is therefore LOOKED AT one stage ahead: after
      // the first tap of a stencil row has been queued (12 MMAs, ~550 clk of tensor work) the warp polls the NEXT weight
      // stage - and the next activation window when this is the last row of the current one - before it issues the other taps.
      int n_my_units = 0;
      for (int u = u_first; u < n_units; u += u_stride) ++n_my_units;
      int rows_per_chunk = 0;
      for (int g = 0; g < p.n_groups; ++g) rows_per_chunk += p.groups[g].ntaps / 3;
      long long b_left = (long long)n_my_units * p.n_chunks * rows_per_chunk;      // weight stages still to be consumed
      long long a_left = (long long)n_my_units * p.n_chunks * p.n_groups;          // activation windows still to be consumed
      bool a_ready = false, b_ready = false;                                         // the current stage has already been waited for
      for (int u = u_first; u < n_units; u += u_stride, ++it) {
        const int buf = it & 1;
        SSDN_TIMED(w_tmem, umma::mbar_wait(tmem_empty(buf), ((it >> 1) & 1) ^ 1, abort_addr, p.error_flag, 3));
        umma::tc_fence_after();
        const uint32_t d0 = tmem + buf * T * p.N;
        uint32_t first = 0;
        for (int ch = 0; ch < p.n_chunks; ++ch) {
          const bool two = (ch != p.n_chunks - 1) || (p.ksteps_last == 2);
          for (int g = 0; g < p.n_groups; ++g) {
            if (!a_ready) SSDN_TIMED(w_a, umma::mbar_wait(full_a(ra.stage), ra.phase, abort_addr, p.error_flag, 3));
            a_ready = false; --a_left;
            const uint32_t a_stage = a_base + ra.stage * a_stage_bytes;
            const int a_cur = ra.stage;
            ra.advance();                                     // (ra now names the NEXT window)
            const int nrows = p.groups[g].ntaps / 3;
            for (int r = 0; r < nrows; ++r) {
              if (!b_ready) SSDN_TIMED(w_b, umma::mbar_wait(full_b(rb.stage), rb.phase, abort_addr, p.error_flag, 3));
              b_ready = false; --b_left;
              umma::tc_fence_after();
              const uint32_t a0 = (((a_stage + p.groups[g].tap_rel[3 * r] * 64) >> 4) & 0x3fffu) | lbo_bits;
              const uint32_t b0 = (((b_base + rb.stage * b_stage_bytes) >> 4) & 0x3fffu) | lbo_bits;
              const int b_cur = rb.stage;
              rb.advance();                                   // (rb now names the NEXT weight stage)
              // per (tile, k-step) the products lo*hi, hi*lo, hi*hi back to back (collector, see above)
              auto issue_tap = [&](int j) {
                const uint32_t aj = a0 + j * step, bj = b0 + j * slab;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                  if (ks == 0 || two) {
                    const uint32_t ko = ks ? 2u : 0u;
#pragma unroll
                    for (int tile = 0; tile < T; ++tile) {
                      const uint32_t d = d0 + tile * p.N;
                      const uint32_t av = aj + tile * 512 + ko, al = av + a_pl, bv = bj + ko, bl = bv + b_pl;
                      mma(d, al, bv, desc_hi, (j == 0 && ks == 0) ? first : 1u);
                      mma_fill(d, av, bl, desc_hi);      // A_hi: fetched once for the two products that use it
                      mma_last(d, av, bv, desc_hi);
                    }
                  }
                }
              };
              if (!p.lookahead) {                             // short stages (N <= 48, one tile): the look-ahead is pure overhead there
                if (umma::elect_one()) {
                  issue_tap(0); issue_tap(1); issue_tap(2);
                  commit(empty_b(b_cur));
                  if (r == nrows - 1) commit(empty_a(a_cur));
                }
                __syncwarp();
                first = 1;
                continue;
              }
              if (umma::elect_one()) issue_tap(0);
              __syncwarp();
              // ONE non-blocking look at the next weight stage (and at the next activation window when this is the last row of
              // the current one) while the first tap is in the queue: if it has landed, the blocking wait at the top of the next
              // stage - whose fixed cost would idle the pipe - is skipped; if not, nothing is lost (blocking here instead made the
              // N = 48 layers, whose short weight stages are latency-bound, 9 % slower)
              if (b_left > 0) b_ready = __all_sync(0xffffffffu, umma::mbar_try_wait(full_b(rb.stage), rb.phase));
              if (r == nrows - 1 && a_left > 0) a_ready = __all_sync(0xffffffffu, umma::mbar_try_wait(full_a(ra.stage), ra.phase));
              if (umma::elect_one()) {
                issue_tap(1);
                issue_tap(2);
                commit(empty_b(b_cur));
                if (r == nrows - 1) commit(empty_a(a_cur));
              }
              __syncwarp();
              first = 1;
            }
          }
        }
        if (umma::elect_one()) commit(tmem_full(buf));
        __syncwarp();
      }
    } else
    for (int u = u_first; u < n_units; u += u_stride, ++it) {
      const int buf = it & 1;
      SSDN_TIMED(w_tmem, umma::mbar_wait(tmem_empty(buf), ((it >> 1) & 1) ^ 1, abort_addr, p.error_flag, 3));
      umma::tc_fence_after();
      const uint32_t d0 = tmem + buf * T * p.N;
      uint32_t first = 0;       // becomes 1 after the first tap: accumulate flag of the very first MMA of each tile
      int sb = 0;               // slab index inside the current B stage
      const uint32_t slab_bytes = 2 * p.N * 64;
      for (int ch = 0; ch < p.n_chunks; ++ch) {
        const bool two = (ch != p.n_chunks - 1) || (p.ksteps_last == 2);
        for (int g = 0; g < p.n_groups; ++g) {
          SSDN_TIMED(w_a, umma::mbar_wait(full_a(ra.stage), ra.phase, abort_addr, p.error_flag, 3));
          const uint32_t av0 = a_base + ra.stage * a_stage_bytes;
          for (int t = 0; t < p.groups[g].ntaps; ++t) {
            if (sb == 0) SSDN_TIMED(w_b, umma::mbar_wait(full_b(rb.stage), rb.phase, abort_addr, p.error_flag, 3));
            umma::tc_fence_after();
            const uint32_t av = av0 + p.groups[g].tap_rel[t] * 64, al = av + p.a_plane_bytes;
            const uint32_t bv = b_base + rb.stage * b_stage_bytes + sb * slab_bytes, bl = bv + p.N * 64;
            const bool last_slab = (sb + 1 == p.bg);
            if (umma::elect_one()) {
              if (!(p.debug & 2))
#pragma unroll
              for (int tile = 0; tile < T; ++tile) {
                const uint32_t d = d0 + tile * p.N;
                const uint32_t ao = tile * (128 * 64);
                umma::mma_f16_ss(d, umma::desc_at(desc, al + ao), umma::desc_at(desc, bv), idesc, first);
                umma::mma_f16_ss(d, umma::desc_at(desc, av + ao), umma::desc_at(desc, bl), idesc, 1);
                umma::mma_f16_ss(d, umma::desc_at(desc, av + ao), umma::desc_at(desc, bv), idesc, 1);
                if (two) {
                  umma::mma_f16_ss(d, umma::desc_at(desc, al + ao + 32), umma::desc_at(desc, bv + 32), idesc, 1);
                  umma::mma_f16_ss(d, umma::desc_at(desc, av + ao + 32), umma::desc_at(desc, bl + 32), idesc, 1);
                  umma::mma_f16_ss(d, umma::desc_at(desc, av + ao + 32), umma::desc_at(desc, bv + 32), idesc, 1);
                }
              }
              if (last_slab) commit(empty_b(rb.stage));
            }
            __syncwarp();
            first = 1;
            if (last_slab) { sb = 0; rb.advance(); } else ++sb;
          }
          if (umma::elect_one()) commit(empty_a(ra.stage));
          __syncwarp();
          ra.advance();
        }
      }
      if (umma::elect_one()) commit(tmem_full(buf));
      __syncwarp();
    }
    if (p.stats && lane == 0) {
      unsigned long long* s = p.stats + blockIdx.x * 16;
      s[2] = w_tmem; s[3] = w_a; s[4] = w_b; s[5] = clock64() - t_start;
    }
  } else if (warp >= 4 && warp < 4 + 4 * p.epi_per_quad) {
    // ------------------------------------------------------------ epilogue
    // TMEM -> registers (one pixel per lane, 32 channels per slice) -> scales / bias / LeakyReLU (+ sign-mask word out) or
    // LeakyReLU' from the sign-mask word -> optional column sums -> staged through shared memory -> fp16 hi/lo split ->
    // global memory with the upsample / (un-)rotate scatter in the address.
    // WORK ITEMS: the (tile, slice) pairs of this CTA's units, numbered in consumption order w = unit * T * nsl + tile * nsl + s;
    // warp j of a TMEM lane quadrant takes the items with w % kEpiPerQuad == j, so all three warps of a quadrant are busy
    // whatever N is (N = 48: 4 items per unit over 3 warps; the old "slice s belongs to warp s % 3" left one warp idle and two
    // with double work).  The epilogue is what binds the N = 48 / upsampling / 1x1 layers: the instruction count per item is
    // what matters here (profiles/r02_epilogue_source_counters.txt).
    const ConvDst& d = p.dst;
    const Geom& sg = p.src;
    const int ew = (warp - 4) & 3;        // TMEM lane quadrant (a warp may only read lanes 32 * (warp % 4) ...)
    const int half = (warp - 4) >> 2;     // which of the quadrant's warps
    const int n_epi_warps = 4 * p.epi_per_quad;
    const int stride = p.epi_per_quad;
    const int nsl = (p.N + 31) >> 5, ipu = T * nsl;      // slices per tile, items per unit
    __shared__ __align__(16) float s_bias[400];
    // accumulators carry 2^(k_a + k_b); operand destinations are written with their own scale 2^k_dst, which is folded into
    // the accumulator scale and the bias (LeakyReLU commutes with a positive factor): one FFMA per element
    const float dst_scale = ((d.flags & EP_WRITE_LO) && d.scale.k) ? exp2_int(__ldg(d.scale.k)) : 1.0f;
    const float inv_dst_scale = 1.0f / dst_scale;          // exact: a power of two
    const float out_scale = p.acc_comp * exp2_int(-((p.k_a ? __ldg(p.k_a) : 0) + (p.k_b ? __ldg(p.k_b) : 0))) * dst_scale;
    if (d.flags & EP_BIAS)
      for (int i = threadIdx.x - 128; i < 400; i += 32 * n_epi_warps) s_bias[i] = i < d.cvalid ? __ldg(d.bias + i) * dst_scale : 0.f;
    asm volatile("bar.sync 1, %0;" ::"r"(32 * n_epi_warps) : "memory");
    const uint32_t bias_s = umma::smem_u32(s_bias);
    const uint32_t stage_s = sbase + p.epi_off + (uint32_t)(warp - 4) * (32 * kStagePitch * 4);     // this warp's staging area
    const uint32_t row_s = stage_s + lane * (kStagePitch * 4);
    float csum[kMaxSlices];
#pragma unroll
    for (int s = 0; s < kMaxSlices; ++s) csum[s] = 0.f;
    float amax_l = 0.f;        // running max|v| (in destination-scaled units) of what this lane wrote
    long long w_full = 0;
    const long long t_start = clock64();
    const bool up2 = d.map == MAP_UP2;
    const uint32_t tmem_empty_lead0 = PAIR ? umma::mapa(tmem_empty(0), 0) : tmem_empty(0);   // the MMA issuer's barrier
    // cursor over this warp's items: (c_it, c_u, c_i) = unit counter, unit, item inside the unit
    int c_it = 0, c_u = u_first, c_i = half;
    auto normalise = [&]() { while (c_i >= ipu && c_u < n_units) { c_i -= ipu; ++c_it; c_u += u_stride; } };
    normalise();
    LanePixel nxt;
    if (c_u < n_units) { const int t0 = c_i >= nsl ? 1 : 0; lane_pixel(p, unit_m(c_u), c_u % p.n_tiles_n, t0, c_i - t0 * nsl, T, ew, lane, nxt); }
    int it = 0;
    for (int u = u_first; u < n_units; u += u_stride, ++it) {
      const int buf = it & 1;
      const int nt = u % p.n_tiles_n;
      SSDN_TIMED(w_full, umma::mbar_wait(tmem_full(buf), (it >> 1) & 1, abort_addr, p.error_flag, 4, p.wait_hint));
      umma::tc_fence_after();
#pragma unroll 1
      while (c_it == it && c_u < n_units) {
        const int tile = c_i >= nsl ? 1 : 0, s = c_i - tile * nsl;
        const LanePixel cur = nxt;
        {   // next item of this warp: its pixel mapping + mask word (the load stays in flight while this item is written)
          c_i += stride;
          normalise();
          if (c_u < n_units) { const int t1 = c_i >= nsl ? 1 : 0; lane_pixel(p, unit_m(c_u), c_u % p.n_tiles_n, t1, c_i - t1 * nsl, T, ew, lane, nxt); }
        }
        const int c0 = 32 * s;
        // A slice is always handled as 32 columns: when N is not a multiple of 32 the last slice reads 16 columns that
        // belong to nobody (all 512 TMEM columns are allocated) and channel validity is enforced where values leave
        // the warp.  No per-element predication => less than half the instructions (the epilogue is issue-bound).
        const int cg0 = (d.map == MAP_UNROT_INV ? 0 : nt * p.N) + c0;  // first channel of the slice among this conv's outputs
        const uint32_t trow = tmem + (uint32_t(ew * 32) << 16) + (buf * T + tile) * p.N;
        uint32_t r[32];
        umma::tmem_ld16(trow + c0, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
        umma::tmem_ld16(trow + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
        umma::tmem_ld_wait();
        if (cg0 < d.cvalid) {
          float f[32];
          if (d.flags & EP_BIAS) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 bq = lds128(bias_s + (cg0 + i) * 4);     // zero beyond cvalid
              f[i] = fmaf(__uint_as_float(r[i]), out_scale, bq.x); f[i + 1] = fmaf(__uint_as_float(r[i + 1]), out_scale, bq.y);
              f[i + 2] = fmaf(__uint_as_float(r[i + 2]), out_scale, bq.z); f[i + 3] = fmaf(__uint_as_float(r[i + 3]), out_scale, bq.w);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(r[i]) * out_scale;
          }
          const uint32_t cur_nd = (cur.pk >> 26) & 7u;
          if (d.flags & EP_LRELU) {
            uint32_t word = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) { const bool pos = f[i] > 0.f; word |= (pos ? 1u : 0u) << i; f[i] = pos ? f[i] : SSDN_LRELU_SLOPE * f[i]; }
            if (d.mask_out && cur_nd == 1)
              d.mask_out[(long long)(cur.pk & kPkPixelMask) * d.mask_out_words + (((cur.pk >> 30) * d.cvalid + cg0) >> 5)] = (cur.pk & (1u << 29)) ? 0u : word;
          }
          if (d.flags & EP_ACT_GRAD) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = ((cur.mw >> i) & 1u) ? f[i] : SSDN_LRELU_SLOPE * f[i];
          }
          if (d.colsum && cur_nd == 0) {        // halo / out-of-range pixels hold garbage accumulators: they must not reach the column sums
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = 0.f;
          }
          // Every slice goes through shared memory (36-float rows: conflict-free float4 access).  NHWC destinations:
          // consecutive lanes then write one pixel's contiguous bytes (fp16 planes: 4 lanes x 16 bytes per plane, fp32:
          // 8 lanes x 16 bytes) - whole sectors per store instruction instead of 32 fragments, which is what the load/store
          // unit can sustain (upsampling writes each value 4 times).  NCHW destination (network output): a lane keeps its
          // own pixel and walks the channels.
#pragma unroll
          for (int i = 0; i < 32; i += 4) sts128(row_s + i * 4, make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]));
          __syncwarp();
          if (d.colsum) {
            // column sums of the staged tile (bias gradient): lane c adds channel c of the 32 pixel rows - bank (4 px + c) % 32,
            // conflict-free - in a fixed order; half the instructions of a shuffle transpose-reduce over the registers
            float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll
            for (int px = 0; px < 32; px += 4) {
              t0 += lds32(stage_s + ((px + 0) * kStagePitch + lane) * 4); t1 += lds32(stage_s + ((px + 1) * kStagePitch + lane) * 4);
              t2 += lds32(stage_s + ((px + 2) * kStagePitch + lane) * 4); t3 += lds32(stage_s + ((px + 3) * kStagePitch + lane) * 4);
            }
            const float total = (t0 + t1) + (t2 + t3);
#pragma unroll
            for (int k = 0; k < kMaxSlices; ++k) csum[k] += (k == s) ? total : 0.f;
          }
          if (d.map == MAP_NCHW) {
            if (cur_nd) {
              const long long hw = (long long)sg.H * sg.W;
              float* dst = d.v + ((long long)cur.b * d.cvalid + cg0) * hw + (long long)cur.y * sg.W + cur.x;
              const int nc = min(min(32, p.N - c0), d.cvalid - cg0);     // a slice may be cut by the N tile or by cvalid
              for (int c = 0; c < nc; ++c) dst[c * hw] = lds32(row_s + c * 4);
            }
          } else if (d.flags & EP_WRITE_LO) {
            // fp16 operand planes: 4 lanes per pixel (8 channels = 16 bytes per plane each), 8 pixels per pass
            const int sub = lane >> 2, q8 = lane & 3;
            const bool chan_ok = (c0 + 8 * q8 < p.N) && (cg0 + 8 * q8 < d.cvalid);
            const int cb0 = d.coff + cg0 + 8 * q8;
#pragma unroll
            for (int q = 0; q < 32; q += 8) {
              const int px = q + sub;
              const uint32_t ppk = __shfl_sync(0xffffffffu, cur.pk, px);
              if (chan_ok && (ppk & (7u << 26))) {
                const float4 o0 = lds128(stage_s + (px * kStagePitch + 8 * q8) * 4);
                const float4 o1 = lds128(stage_s + (px * kStagePitch + 8 * q8 + 4) * 4);
                float f8[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
                if (ppk & (1u << 29)) {                      // row shifted in by Shift2d
#pragma unroll
                  for (int i = 0; i < 8; ++i) f8[i] = 0.f;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) amax_l = fmaxf(amax_l, fabsf(f8[i]));
                uint4 h, l;
                f16_split8(f8, h, l);
                const long long oi = (long long)(ppk & kPkPixelMask) * d.cpitch + (cb0 + (int)(ppk >> 30) * d.cvalid);
                *reinterpret_cast<uint4*>(d.hi + oi) = h;
                *reinterpret_cast<uint4*>(d.lo + oi) = l;
                if (up2) {                                   // nearest-neighbour upsample: the 2 x 2 block
                  const long long o1i = oi + d.cpitch, o2i = oi + (long long)d.g.P * d.cpitch, o3i = o2i + d.cpitch;
                  *reinterpret_cast<uint4*>(d.hi + o1i) = h; *reinterpret_cast<uint4*>(d.lo + o1i) = l;
                  *reinterpret_cast<uint4*>(d.hi + o2i) = h; *reinterpret_cast<uint4*>(d.lo + o2i) = l;
                  *reinterpret_cast<uint4*>(d.hi + o3i) = h; *reinterpret_cast<uint4*>(d.lo + o3i) = l;
                }
              }
            }
          } else {
            const int sub = lane >> 3, q4 = lane & 7;                  // 8 lanes per pixel, 4 pixels per pass
            const bool chan_ok = (c0 + 4 * q4 < p.N) && (cg0 + 4 * q4 < d.cvalid);
            const int cb0 = d.coff + cg0 + 4 * q4;
#pragma unroll 4
            for (int q = 0; q < 32; q += 4) {
              const int px = q + sub;
              const uint32_t ppk = __shfl_sync(0xffffffffu, cur.pk, px);
              if (chan_ok && (ppk & (7u << 26))) {
                float4 o = lds128(stage_s + (px * kStagePitch + 4 * q4) * 4);
                if (ppk & (1u << 29)) o = make_float4(0.f, 0.f, 0.f, 0.f);     // row shifted in by Shift2d
                const long long oi = (long long)(ppk & kPkPixelMask) * d.cpitch + (cb0 + (int)(ppk >> 30) * d.cvalid);
                *reinterpret_cast<float4*>(d.v + oi) = o;
                if (up2) {
                  const long long o2i = oi + (long long)d.g.P * d.cpitch;
                  *reinterpret_cast<float4*>(d.v + oi + d.cpitch) = o;
                  *reinterpret_cast<float4*>(d.v + o2i) = o;
                  *reinterpret_cast<float4*>(d.v + o2i + d.cpitch) = o;
                }
              }
            }
          }
          __syncwarp();
        }
      }
      umma::tc_fence_before();
      if (PAIR) umma::mbar_arrive_cluster(tmem_empty_lead0 + 8 * buf); else umma::mbar_arrive(tmem_empty(buf));
    }
    if (d.flags & EP_WRITE_LO) amax_commit(d.scale.amax, amax_l * inv_dst_scale);
    if (d.colsum) {
      // this CTA only ever sees one N tile when gridDim.x is a multiple of n_tiles_n (the host guarantees it); a warp may have
      // handled any slice of it.  Sums were taken in destination-scaled units.
      const int nt = u_first % p.n_tiles_n;
      const int c_first = (d.map == MAP_UNROT_INV) ? 0 : nt * p.N;
      float* row = d.colsum + (long long)(blockIdx.x * n_epi_warps + (warp - 4)) * d.colsum_pitch;
      for (int c = lane; c < d.colsum_pitch; c += 32) row[c] = 0.f;
      __syncwarp();
      // lane c holds the column sum of channel c of each slice
#pragma unroll
      for (int s = 0; s < kMaxSlices; ++s)
        if (32 * s + lane < p.N && c_first + 32 * s + lane < d.colsum_pitch && (u_first < n_units)) row[c_first + 32 * s + lane] = csum[s] * inv_dst_scale;
    }
    if (p.stats && threadIdx.x == 128) { p.stats[blockIdx.x * 16 + 6] = w_full; p.stats[blockIdx.x * 16 + 7] = clock64() - t_start; }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (PAIR) umma::cluster_sync();       // the peer may still be signalling this CTA's barriers / reading its shared memory
  if (warp == 3) { if (PAIR) umma::tmem_dealloc_pair(tmem, 512); else umma::tmem_dealloc(tmem, 512); }
}

}  // namespace convk

// ------------------------------------------------------------------------------------------ host
#include <algorithm>
#include <cstdlib>
#include <vector>

struct ConvTaps { int n; int off[9]; };
#ifndef SSDN_WAIT_HINT_DEFAULT
#define SSDN_WAIT_HINT_DEFAULT 0
#endif
// ceil(2^64 / d) for d >= 2: umulhi64(n, m) == n / d for every n < 2^32 (convk::div_magic)
static inline uint64_t div_magic_for(uint32_t d) { return ~0ULL / d + 1; }   // flat-pixel offsets of the taps, in weight-slab order

// 1x1 convolutions whose input channels come in whole 64-channel groups use the wide chunk (see ConvParams::wide).
static inline bool conv_is_wide(int cin, int ntaps) { return ntaps == 1 && cin % 64 == 0; }
// Number of channel chunks (32 wide, or 64 in wide mode) and k-steps (of 16 channels) of the last chunk for `cin` input channels.
static inline void conv_chunks(int cin, int* n_chunks, int* ksteps_last, bool wide = false) {
  if (wide) { *n_chunks = cin / 64; *ksteps_last = 4; return; }
  *n_chunks = (cin + 31) / 32;
  const int rem = cin - (*n_chunks - 1) * 32;
  *ksteps_last = rem > 16 ? 2 : 1;
}

// Fills plan->p (everything except tensor maps' base pointers) and the tensor maps.
//   a_hi/a_lo : source planes (fp16) with `a_cpitch` channels per pixel, the conv reads channels [a_coff, a_coff+cin)
//   w_slab    : prepared weight slab [n_tiles_n][n_chunks][ntaps][plane][N][32] (see pw::weight_prep_kernel)
//   k_a / k_b : device pointers to the scale exponents of the A tensor and of the weights
static inline int conv_plan_init(ConvPlan* plan, const Geom& src, const __half* a_hi, const __half* a_lo, int a_cpitch,
                                 int a_coff, int cin, const __half* w_slab, int cout_padded, int N,
                                 const ConvTaps& taps, const ConvDst& dst, int* error_flag, int num_sms,
                                 const int* k_a, const int* k_b, size_t smem_limit = 222 * 1024) {
  ConvParams& p = plan->p;
  p = ConvParams{};
  p.src = src; p.N = N; p.dst = dst; p.error_flag = error_flag; p.k_a = k_a; p.k_b = k_b;
  p.debug = getenv("SSDN_CONV_DEBUG") ? atoi(getenv("SSDN_CONV_DEBUG")) : 0;
  {
    static const int hint = getenv("SSDN_WAIT_HINT") ? atoi(getenv("SSDN_WAIT_HINT")) : SSDN_WAIT_HINT_DEFAULT;
    p.wait_hint = (uint32_t)hint;
  }
  // the epilogue's index arithmetic: 32-bit flat pixels, destination pixels packed into 26 bits (LanePixel::pk)
  if (src.total() + 512 >= (1LL << 31) || src.S < 2 || src.P < 2) return -17;
  if (dst.map != MAP_NCHW && dst.g.total() >= (1LL << 26)) return -17;
  p.magic_s = div_magic_for((uint32_t)src.S); p.magic_p = div_magic_for((uint32_t)src.P);
  p.n_tiles_n = cout_padded / N;
  p.wide = conv_is_wide(cin, taps.n) ? 1 : 0;
  conv_chunks(cin, &p.n_chunks, &p.ksteps_last, p.wide);
  const int cw_ch = p.wide ? 64 : 32;
  p.ntaps_total = taps.n;
  p.T = (2 * 2 * N <= 512) ? 2 : 1;
  const long long total = src.total();
  if (p.T == 2) {
    // two tiles per unit halve the weight traffic, but on the small pyramid levels they leave SMs idle: compare waves x work
    const long long tiles = (total + 127) / 128;
    const long long cost2 = (((tiles + 1) / 2 * p.n_tiles_n + num_sms - 1) / num_sms) * 2, cost1 = (tiles * p.n_tiles_n + num_sms - 1) / num_sms;
    if (cost1 * 5 <= cost2 * 4) p.T = 1;
  }
  p.n_units_m = (int)((total + 128LL * p.T - 1) / (128LL * p.T));
  // B stage = bg consecutive weight slabs (both planes) loaded by ONE TMA op: a TMA instruction costs ~450 clk of the
  // SM's TMA unit whatever its size (profiles/r01_tma_rate.log), so operands must arrive in few, large boxes.
  const int n_slabs = p.n_chunks * taps.n;
  p.bg = p.wide ? 1 : ((taps.n % 3 == 0) ? 3 : (n_slabs % 3 == 0 ? 3 : (n_slabs % 2 == 0 ? 2 : 1)));
  // CTA pairs (cta_group::2): 3x3 stencils whose rows advance by +-1 pixel (the row3 issue path), N/2 a whole number of
  // 8-row swizzle groups, and at least one pair-unit per cluster worth of work
  {
    bool rows_ok = taps.n == 9 && p.bg == 3;
    for (int r = 0; r < 3 && rows_ok; ++r) {
      const int s1 = taps.off[3 * r + 1] - taps.off[3 * r], s2 = taps.off[3 * r + 2] - taps.off[3 * r + 1];
      rows_ok = s1 == s2 && (s1 == 1 || s1 == -1);
    }
    const char* e = getenv("SSDN_CONV_PAIR");
    // measured (profiles/r01_pair_split_ablation.log): pairs never lose except on the one-chunk first convolution
    // (1x1 layers are shared-memory-bound INCLUDING the TMA fill traffic - no tap reuses a staged byte - and a pair halves the
    // B bytes per CTA: measured -12 % on the un-rotating head data-gradient (4 N tiles of 96), +8 % on the 96-wide head conv)
    const bool wide_ok = p.wide && (cout_padded >= 192 || (e && atoi(e) == 2)) && p.n_chunks >= 4;   // SSDN_CONV_PAIR=2: every deep wide layer (ablation)
    p.pair = (rows_ok || wide_ok) && (N % 16 == 0) && (num_sms % 2 == 0) && p.n_chunks >= 2 && !(e && atoi(e) == 0);
    if (p.wide && e && atoi(e) == 3) p.pair = 0;   // ablation: pairs on the 3x3 layers only
  }
  p.b_stage_bytes = (uint32_t)(p.bg * 2 * (p.pair ? N / 2 : N) * cw_ch * 2);   // per CTA
  {
    const int ksteps = p.wide ? 4 * p.n_chunks : 2 * (p.n_chunks - 1) + p.ksteps_last;
    const bool on = !(getenv("SSDN_ACC_COMP") && atoi(getenv("SSDN_ACC_COMP")) == 0);
    p.acc_comp = on ? 1.0f + SSDN_ACC_BETA * (float)(ksteps * taps.n * 3) : 1.0f;
  }
  // epilogue: both warps of a TMEM lane quadrant - measured never slower (profiles/r01_pair_split_ablation.log), except
  // with the upsampling epilogue (4x the stores: the two warps then only fight over the load/store unit)
  {
    p.epi_per_quad = dst.map != MAP_NCHW ? convk::kEpiPerQuad : 1;      // (upsampling epilogues too, now that they write whole sectors: -6 %)
    if (const char* e = getenv("SSDN_EPI_SPLIT")) p.epi_per_quad = atoi(e) != 0 ? convk::kEpiPerQuad : 1;
    if (const char* e = getenv("SSDN_EPI_PER_QUAD")) { const int v = atoi(e); if (v >= 1 && v <= convk::kEpiPerQuad && p.epi_per_quad > 1) p.epi_per_quad = v; }
  }
  // choose the tap grouping: all taps in one window if it fits in shared memory, else one window per
  // distinct row offset (dy), else one window per tap.
  int rows_unit = 128 * p.T;
  auto try_group = [&](int mode) -> bool {
    // mode 0: single window, 1: group by rows of the 3x3 stencil (taps sorted in slab order, 3 per row), 2: per tap
    std::vector<std::vector<int>> gs;
    if (mode == 0) { gs.emplace_back(); for (int t = 0; t < taps.n; ++t) gs[0].push_back(t); }
    else if (mode == 1 && taps.n == 9) { for (int r = 0; r < 3; ++r) gs.push_back({3 * r, 3 * r + 1, 3 * r + 2}); }
    else { for (int t = 0; t < taps.n; ++t) gs.push_back({t}); }
    int max_rows = 0;
    for (auto& g : gs) {
      int lo = INT32_MAX, hi = INT32_MIN;
      for (int t : g) { lo = std::min(lo, taps.off[t]); hi = std::max(hi, taps.off[t]); }
      max_rows = std::max(max_rows, rows_unit + hi - lo);
    }
    int nbox = (max_rows + 255) / 256;
    int box_rows = ((max_rows + nbox - 1) / nbox + 15) / 16 * 16;      // multiple of 16 rows => planes/boxes stay 1024-byte aligned
    if (box_rows > 256) { ++nbox; box_rows = ((max_rows + nbox - 1) / nbox + 15) / 16 * 16; }
    uint32_t plane = (uint32_t)(nbox * box_rows * cw_ch * 2);
    static const int max_a_stages = getenv("SSDN_CONV_A_STAGES") ? atoi(getenv("SSDN_CONV_A_STAGES")) : 3;
    for (int stages = std::max(2, std::min(3, max_a_stages)); stages >= 2; --stages)
    for (int bst = 4; bst >= 2; --bst) {
      const size_t epi = (size_t)4 * p.epi_per_quad * 32 * convk::kStagePitch * sizeof(float);   // epilogue staging rows of the working warps
      size_t need = (size_t)stages * 2 * plane + (size_t)bst * p.b_stage_bytes + epi + 2048;
      static const size_t limit_env = getenv("SSDN_CONV_SMEM_KB") ? (size_t)atoi(getenv("SSDN_CONV_SMEM_KB")) * 1024 : 0;   // ablation: fewer stages
      if (need > (limit_env ? std::min(limit_env, smem_limit) : smem_limit)) continue;
      p.epi_off = (uint32_t)((size_t)stages * 2 * plane + (size_t)bst * p.b_stage_bytes);
      p.n_groups = (int)gs.size(); p.nbox = nbox; p.box_rows = box_rows; p.a_plane_bytes = plane; p.a_stages = stages; p.b_stages = bst;
      for (size_t gi = 0; gi < gs.size(); ++gi) {
        int lo = INT32_MAX;
        for (int t : gs[gi]) lo = std::min(lo, taps.off[t]);
        TapGroup& tg = p.groups[gi];
        tg.row_off = lo; tg.ntaps = (int)gs[gi].size();
        for (size_t k = 0; k < gs[gi].size(); ++k) { tg.tap_id[k] = gs[gi][k]; tg.tap_rel[k] = taps.off[gs[gi][k]] - lo; }
      }
      plan->smem = need;
      return true;
    }
    return false;
  };
  auto try_all = [&]() { return try_group(0) || try_group(1) || try_group(2); };
  if (!try_all()) {
    bool ok = false;
    if (p.T == 2) {                                      // two tiles per unit do not fit in shared memory (wide 1x1 layers with N = 128)
      p.T = 1; rows_unit = 128;
      p.n_units_m = (int)((total + 127) / 128);
      ok = try_all();
    }
    if (!ok && taps.n == 1 && p.bg > 1 && !p.pair) {     // 1x1 layers with wide N: one weight slab per B stage instead of three
      p.bg = 1;
      p.b_stage_bytes = (uint32_t)(p.bg * 2 * N * cw_ch * 2);
      ok = try_all();
    }
    if (!ok) return -10;
  }
  // fast issue path: every B stage = the three taps of one stencil row, with a constant step between their A offsets
  p.row3 = 0; p.tap_step = 0;
  if (taps.n == 9 && p.bg == 3) {
    bool ok = true; int step = 0;
    for (int g = 0; g < p.n_groups && ok; ++g) {
      if (p.groups[g].ntaps % 3) { ok = false; break; }
      for (int r = 0; r < p.groups[g].ntaps / 3; ++r) {
        const int s1 = p.groups[g].tap_rel[3 * r + 1] - p.groups[g].tap_rel[3 * r], s2 = p.groups[g].tap_rel[3 * r + 2] - p.groups[g].tap_rel[3 * r + 1];
        if (s1 != s2 || (s1 != 1 && s1 != -1) || (step && s1 != step)) ok = false;
        step = s1;
      }
    }
    if (ok) { p.row3 = 1; p.tap_step = step; }
  }
  // measured (interleaved A/B against the build without it): -5 % / -2.5 % on the two largest layers, -2 % on the data-gradients,
  // +4 % on the 48-channel layers and +1-2 us on every one-tile layer if it were on everywhere
  p.lookahead = (p.row3 && N >= 96 && !(getenv("SSDN_LOOKAHEAD") && atoi(getenv("SSDN_LOOKAHEAD")) == 0)) ? 1 : 0;
  if (getenv("SSDN_LOOKAHEAD") && atoi(getenv("SSDN_LOOKAHEAD")) == 2) p.lookahead = p.row3;
  if (p.pair && !p.row3 && !p.wide) return -14;      // pairs are implemented for the row3 and the wide issue paths
  if (p.pair) {
    p.n_pairs = (p.n_units_m + 1) / 2 * p.n_tiles_n;
    int clusters = std::min(p.n_pairs, num_sms / 2);
    if (dst.colsum) { clusters -= clusters % p.n_tiles_n; if (clusters < p.n_tiles_n) return -12; }
    plan->grid = 2 * clusters;
  } else {
    plan->grid = std::min(p.n_units_m * p.n_tiles_n, num_sms);
    if (dst.colsum) {                                  // column sums: every CTA must stay on one N tile (see the epilogue)
      plan->grid -= plan->grid % p.n_tiles_n;
      if (plan->grid < p.n_tiles_n) return -12;
    }
  }
  if (dst.map != MAP_NCHW && (dst.cvalid % 4 || dst.cpitch % 4 || dst.coff % 4)) return -13;   // float4 stores
  if ((dst.flags & EP_WRITE_LO) && (dst.cvalid % 8 || dst.cpitch % 8 || dst.coff % 8 || !dst.hi || !dst.lo)) return -13;   // 8 halves per store
  if (a_cpitch % 8 || a_coff % 8) return -16;                                                  // TMA: 16-byte strides and base
  if ((dst.mask_out || dst.mask_in) && p.n_tiles_n > 1 && N % 32) return -15;                  // mask words are per 32 channels
  // tensor maps.  A: 3-D (channel, flat pixel, plane); the lo plane must follow the hi plane at a constant byte distance.
  const long long plane_stride = (long long)((const char*)a_lo - (const char*)a_hi);
  if (plane_stride <= 0 || plane_stride % 16) return -11;
  uint64_t adims[3] = {(uint64_t)cin, (uint64_t)total, 2};
  uint64_t astr[2] = {(uint64_t)a_cpitch * 2, (uint64_t)plane_stride};
  uint32_t abox[3] = {(uint32_t)cw_ch, (uint32_t)p.box_rows, (uint32_t)(p.nbox == 1 ? 2 : 1)};
  const CUtensorMapSwizzle swz = p.wide ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  int r;
  if ((r = umma::encode_f16(&plan->a, (void*)(a_hi + a_coff), 3, adims, astr, abox, swz))) return r;
  // B: slabs [n_tile][chunk][tap][plane][N][32] -> 3-D (32, N, slab x plane), box = bg slabs x 2 planes
  uint64_t bdims[3] = {(uint64_t)cw_ch, (uint64_t)N, (uint64_t)p.n_tiles_n * n_slabs * 2};
  uint64_t bstr[2] = {(uint64_t)cw_ch * 2, (uint64_t)N * cw_ch * 2};
  uint32_t bbox[3] = {(uint32_t)cw_ch, (uint32_t)(p.pair ? N / 2 : N), (uint32_t)(2 * p.bg)};
  if ((r = umma::encode_f16(&plan->b, (void*)w_slab, 3, bdims, bstr, bbox, swz))) return r;
  return 0;
}

// halves of the combined (hi + lo) weight slab
static inline size_t conv_weight_slab_halves(int cin, int cout_padded, int ntaps) {
  const bool wide = conv_is_wide(cin, ntaps);
  int nc, kl; conv_chunks(cin, &nc, &kl, wide);
  return (size_t)cout_padded * nc * ntaps * (wide ? 64 : 32) * 2;
}

// Optional per-launch timing (bench.py roofline): when enabled every kernel launch of the engine is bracketed by CUDA events
// on the stream it is launched on, with its algorithmic FLOPs and HBM bytes.
enum ProfKind : int { K_CONV_FWD = 0, K_CONV_DGRAD, K_WGRAD, K_WGRAD_REDUCE, K_POOL_FWD, K_POOL_BWD, K_UP_BWD, K_PACK, K_WEIGHT_PREP,
                      K_BIAS, K_SCALE, K_POSTERIOR_FWD, K_POSTERIOR_BWD, K_ADAM, K_OTHER, K_COUNT };
struct LaunchProfiler {
  bool on = false;
  struct Rec { int kind; double flops, bytes; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  void begin(int kind, double flops, double bytes, cudaStream_t st) {
    Rec r{kind, flops, bytes, nullptr, nullptr}; cudaEventCreate(&r.a); cudaEventCreate(&r.b); cudaEventRecord(r.a, st); recs.push_back(r);
  }
  void end(cudaStream_t st) { cudaEventRecord(recs.back().b, st); }
};
inline LaunchProfiler& profiler() { static LaunchProfiler p; return p; }
#define SSDN_PROF(kind, flops, bytes, st, launch) do { const bool on_ = profiler().on; if (on_) profiler().begin(kind, flops, bytes, st); launch; \
                                                       if (on_) profiler().end(st); } while (0)

static inline cudaError_t conv_launch_raw(const ConvPlan& plan, const ConvParams& p, cudaStream_t stream);
// Developer instrumentation: one synchronous launch with the per-role wait clocks collected and printed to stderr.
static inline cudaError_t conv_launch_with_stats(const ConvPlan& plan, cudaStream_t stream, int kind) {
  static unsigned long long* dev = nullptr;
  if (!dev) cudaMalloc(&dev, 1024 * 16 * sizeof(unsigned long long));
  cudaMemsetAsync(dev, 0, 1024 * 16 * sizeof(unsigned long long), stream);
  ConvParams p = plan.p; p.stats = dev;
  profiler().begin(kind, plan.flops, plan.bytes, stream);
  conv_launch_raw(plan, p, stream);
  profiler().end(stream);
  std::vector<unsigned long long> h((size_t)plan.grid * 16);
  cudaMemcpyAsync(h.data(), dev, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream);
  cudaError_t e = cudaStreamSynchronize(stream);
  double s[8] = {0};
  for (int b = 0; b < plan.grid; ++b) for (int k = 0; k < 8; ++k) s[k] += (double)h[(size_t)b * 16 + k] / plan.grid;
  fprintf(stderr, "[conv stats] kind %d pair %d grid %d T %d N %d ntiles_n %d chunks %d groups %d taps %d a_st %d b_st %d bg %d | mma loop %.0f clk: wait tmem_empty %.1f%% "
          "full_a %.1f%% full_b %.1f%% | prodA wait empty %.1f%% prodB wait empty %.1f%% | epi loop %.0f clk: wait tmem_full %.1f%%\n",
          kind, p.pair, plan.grid, p.T, p.N, p.n_tiles_n, p.n_chunks, p.n_groups, p.ntaps_total, p.a_stages, p.b_stages, p.bg, s[5], 100 * s[2] / s[5], 100 * s[3] / s[5],
          100 * s[4] / s[5], 100 * s[0] / s[5], 100 * s[1] / s[5], s[7], 100 * s[6] / s[7]);
  return e != cudaSuccess ? e : cudaGetLastError();
}

// One launch of the right instantiation; pairs go out as clusters of two CTAs.
static inline cudaError_t conv_launch_raw(const ConvPlan& plan, const ConvParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(convk::conv_igemm_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(convk::conv_igemm_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(convk::conv_igemm_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(convk::conv_igemm_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (!p.pair) {
    if (p.T == 2) return launch_pdl(convk::conv_igemm_kernel<2, false>, dim3(plan.grid), dim3(convk::kThreads), plan.smem, stream, plan.a, plan.b, p);
    return launch_pdl(convk::conv_igemm_kernel<1, false>, dim3(plan.grid), dim3(convk::kThreads), plan.smem, stream, plan.a, plan.b, p);
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(plan.grid); cfg.blockDim = dim3(convk::kThreads); cfg.dynamicSmemBytes = plan.smem; cfg.stream = stream;
  cudaLaunchAttribute attr[3];
  cfg.attrs = attr; cfg.numAttrs = launch_attrs(attr, true);
  if (p.T == 2) return cudaLaunchKernelEx(&cfg, convk::conv_igemm_kernel<2, true>, plan.a, plan.b, p);
  return cudaLaunchKernelEx(&cfg, convk::conv_igemm_kernel<1, true>, plan.a, plan.b, p);
}

static inline cudaError_t conv_launch(const ConvPlan& plan, cudaStream_t stream, int kind = 0) {
  static const bool want_stats = getenv("SSDN_CONV_STATS") != nullptr;
  if (want_stats && profiler().on) return conv_launch_with_stats(plan, stream, kind);
  if (profiler().on) profiler().begin(kind, plan.flops, plan.bytes, stream);
  cudaError_t e = conv_launch_raw(plan, plan.p, stream);
  if (profiler().on) profiler().end(stream);
  return e;
}
