// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), forward and data-gradient.
//
//   D[flat pixel j, co] = sum_taps sum_ci  A[j + off_tap, ci] * Wt[tap][co][ci]
//
// A is a padded-flat NHWC activation (common.cuh); because zero padding is stored in the halo,
// every tap is a constant row offset into A.  Work decomposition:
//   unit   = T consecutive tiles of 128 flat pixels  x  one N-tile (<= 128 output channels)
//   chunk  = 16 input channels (one SWIZZLE_64B K-major smem tile row = 64 bytes, 2 tf32 k-steps)
//   group  = a set of taps that share one staged window of A rows (halo reuse: the window is
//            loaded ONCE by TMA and each tap is only a different UMMA start address)
//   B tile = the [N][16] weight slab of one (chunk, tap), streamed through its own smem ring.
// Warp roles: 0 = A producer (TMA), 1 = B producer (TMA), 2 = MMA issuer (one thread),
// 3 = TMEM allocator, 4..7 = epilogue (TMEM -> registers -> bias/LeakyReLU/grad-mask -> HBM,
// with the upsample / un-rotate / NCHW scatter fused into the store address).
// Accumulators are double buffered in TMEM so the epilogue of unit i overlaps the MMAs of unit i+1.
// Precision: 3xTF32 (see common.cuh) => fp32-grade results, 3 MMAs per (tile, k-step).
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
  int n_chunks, ksteps_last;
  int n_groups, ntaps_total;
  TapGroup groups[9];
  int nbox, box_rows; // every group window is loaded as nbox TMA boxes of box_rows rows (one op for both planes if nbox == 1)
  int a_stages, b_stages;
  int bg;             // weight slabs ((chunk, tap) pairs, in consumption order) per B stage: ONE TMA op loads bg x 2 planes
  uint32_t a_plane_bytes, b_stage_bytes;   // smem bytes of one A plane of one stage / of one B stage
  uint32_t epi_off;   // byte offset of the epilogue staging area (4 warps x 32 pixels x 36 floats) in dynamic smem
  ConvDst dst;
  int* error_flag;
  int debug;          // experiments only: 2 = skip MMA issue
};

struct ConvPlan {
  ConvParams p;
  double flops = 0;   // algorithmic 2*MAC of this launch (valid pixels, real channels); filled by the owner
  CUtensorMap a, b;   // a: [plane][flat pixel][channel] (3-D), b: weight slabs [slab x plane][N][16] (3-D)
  int grid; size_t smem;
};

namespace convk {

constexpr int kThreads = 256;
constexpr int kMaxStages = 8;

struct Ring {
  int stage = 0; uint32_t phase = 0; int n;
  __device__ explicit Ring(int n_) : n(n_) {}
  __device__ void advance() { if (++stage == n) { stage = 0; phase ^= 1; } }
};

constexpr int kStagePitch = 36;   // floats per staged pixel row (32 channels + 4 pad: conflict-free float4 access)

// Copy-out of one staged slice (CW = 32 or 16 channels x 32 pixels of this warp).  CW/4 consecutive lanes write one
// pixel's contiguous bytes.  All shared-memory reads and the (read-only) activation loads of the slice are issued
// before the first store, so their latencies overlap instead of serialising pass after pass.
template <int CW>
__device__ __forceinline__ void copy_out_slice(const ConvDst& d, const float* stage, const int (*s_dst)[4], const int* s_meta,
                                               const int* s_src, int lane, int cg0, int act_c0) {
  constexpr int L = CW / 4, PPI = 32 / L, PASSES = 32 / PPI;
  const int sub = lane / L, f = lane - sub * L;
  const int cg = cg0 + 4 * f;
  if (cg >= d.cvalid) return;
  float4 v[PASSES]; float4 a[PASSES]; int meta[PASSES]; int dp0[PASSES];
#pragma unroll
  for (int q = 0; q < PASSES; ++q) {
    const int px = q * PPI + sub;
    meta[q] = s_meta[px];
    dp0[q] = s_dst[px][0];
    v[q] = *reinterpret_cast<const float4*>(stage + px * kStagePitch + 4 * f);
    if ((d.flags & EP_ACT_GRAD) && (meta[q] & 7)) {
      const long long ai = (d.flags & EP_ACT_AT_SRC) ? (long long)s_src[px] * d.act_cpitch + d.act_coff + act_c0 + 4 * f
                                                     : (long long)dp0[q] * d.act_cpitch + d.act_coff + cg;
      a[q] = __ldg(reinterpret_cast<const float4*>(d.act + ai));
    }
  }
#pragma unroll
  for (int q = 0; q < PASSES; ++q) {
    const int nd = meta[q] & 7;
    if (nd == 0) continue;
    float4 o = v[q];
    if (meta[q] & 8) o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (d.flags & EP_ACT_GRAD) {
      o.x = a[q].x > 0.f ? o.x : SSDN_LRELU_SLOPE * o.x; o.y = a[q].y > 0.f ? o.y : SSDN_LRELU_SLOPE * o.y;
      o.z = a[q].z > 0.f ? o.z : SSDN_LRELU_SLOPE * o.z; o.w = a[q].w > 0.f ? o.w : SSDN_LRELU_SLOPE * o.w;
    }
    float4 h = o, l = o;
    if (d.flags & EP_WRITE_LO) { tf32_split(o.x, h.x, l.x); tf32_split(o.y, h.y, l.y); tf32_split(o.z, h.z, l.z); tf32_split(o.w, h.w, l.w); }
    const int coff = d.coff + (meta[q] >> 8) + cg;
    const int px = q * PPI + sub;
    for (int k = 0; k < nd; ++k) {
      const long long oi = (long long)(k == 0 ? dp0[q] : s_dst[px][k]) * d.cpitch + coff;
      *reinterpret_cast<float4*>(d.v + oi) = h;
      if (d.flags & EP_WRITE_LO) *reinterpret_cast<float4*>(d.lo + oi) = l;
    }
  }
}

template <int T>
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
  const int n_units = p.n_units_m * p.n_tiles_n;

  if (threadIdx.x == 0) {
    abort_word = 0;
    for (int s = 0; s < p.a_stages; ++s) { umma::mbar_init(full_a(s), 1); umma::mbar_init(empty_a(s), 1); }
    for (int s = 0; s < p.b_stages; ++s) { umma::mbar_init(full_b(s), 1); umma::mbar_init(empty_b(s), 1); }
    for (int b = 0; b < 2; ++b) { umma::mbar_init(tmem_full(b), 1); umma::mbar_init(tmem_empty(b), 128); }
    umma::fence_mbar_init();
  }
  if (warp == 3) {
    umma::tmem_alloc(umma::smem_u32(&tmem_slot), 512);
    umma::tmem_relinquish();
  }
  if (warp == 0 && lane == 0) umma::tma_prefetch_desc(&map_a);
  if (warp == 1 && lane == 0) umma::tma_prefetch_desc(&map_b);
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ A producer (warp-uniform loop, elected issue)
    Ring ra(p.a_stages);
    const uint32_t box_bytes = p.box_rows * 64;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const int um = u / p.n_tiles_n;
      const int j0 = um * 128 * T;
      for (int ch = 0; ch < p.n_chunks; ++ch)
        for (int g = 0; g < p.n_groups; ++g) {
          umma::mbar_wait(empty_a(ra.stage), ra.phase ^ 1, abort_addr, p.error_flag, 1);
          const uint32_t dst = a_base + ra.stage * a_stage_bytes;
          const int row = j0 + p.groups[g].row_off;
          if (umma::elect_one()) {
            umma::mbar_expect_tx(full_a(ra.stage), 2 * p.nbox * box_bytes);
            if (p.nbox == 1) {      // one op brings both planes: box (16 ch, rows, 2 planes)
              umma::tma_load_3d(dst, &map_a, full_a(ra.stage), ch * 16, row, 0);
            } else {
              for (int bx = 0; bx < p.nbox; ++bx) {
                umma::tma_load_3d(dst + bx * box_bytes, &map_a, full_a(ra.stage), ch * 16, row + bx * p.box_rows, 0);
                umma::tma_load_3d(dst + p.a_plane_bytes + bx * box_bytes, &map_a, full_a(ra.stage), ch * 16, row + bx * p.box_rows, 1);
              }
            }
          }
          __syncwarp();
          ra.advance();
        }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ B producer: one TMA per stage = bg slabs x 2 planes
    Ring rb(p.b_stages);
    const int n_slabs = p.n_chunks * p.ntaps_total, n_bstages = n_slabs / p.bg;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const int nt = u % p.n_tiles_n;
      for (int i = 0; i < n_bstages; ++i) {
        umma::mbar_wait(empty_b(rb.stage), rb.phase ^ 1, abort_addr, p.error_flag, 2);
        if (umma::elect_one()) {
          umma::mbar_expect_tx(full_b(rb.stage), b_stage_bytes);
          umma::tma_load_3d(b_base + rb.stage * b_stage_bytes, &map_b, full_b(rb.stage), 0, 0, 2 * (nt * n_slabs + i * p.bg));
        }
        __syncwarp();
        rb.advance();
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ MMA issuer: uniform loops, one elected lane issues a fully
    // unrolled block of T x 2 x 3 MMAs per (chunk, tap) whose descriptors are base + compile-time offsets
    Ring ra(p.a_stages), rb(p.b_stages);
    constexpr uint64_t desc = umma::make_desc_base(16, 512, umma::LAYOUT_SW64);
    const uint32_t idesc = umma::make_idesc_tf32(128, p.N, 0, 0);
    int it = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
      const int buf = it & 1;
      umma::mbar_wait(tmem_empty(buf), ((it >> 1) & 1) ^ 1, abort_addr, p.error_flag, 3);
      umma::tc_fence_after();
      const uint32_t d0 = tmem + buf * T * p.N;
      uint32_t first = 0;       // becomes 1 after the first tap: accumulate flag of the very first MMA of each tile
      int sb = 0;               // slab index inside the current B stage
      const uint32_t slab_bytes = 2 * p.N * 64;
      for (int ch = 0; ch < p.n_chunks; ++ch) {
        const bool two = (ch != p.n_chunks - 1) || (p.ksteps_last == 2);
        for (int g = 0; g < p.n_groups; ++g) {
          umma::mbar_wait(full_a(ra.stage), ra.phase, abort_addr, p.error_flag, 3);
          const uint32_t av0 = a_base + ra.stage * a_stage_bytes;
          for (int t = 0; t < p.groups[g].ntaps; ++t) {
            if (sb == 0) umma::mbar_wait(full_b(rb.stage), rb.phase, abort_addr, p.error_flag, 3);
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
                umma::mma_tf32_ss(d, umma::desc_at(desc, al + ao), umma::desc_at(desc, bv), idesc, first);
                umma::mma_tf32_ss(d, umma::desc_at(desc, av + ao), umma::desc_at(desc, bl), idesc, 1);
                umma::mma_tf32_ss(d, umma::desc_at(desc, av + ao), umma::desc_at(desc, bv), idesc, 1);
                if (two) {
                  umma::mma_tf32_ss(d, umma::desc_at(desc, al + ao + 32), umma::desc_at(desc, bv + 32), idesc, 1);
                  umma::mma_tf32_ss(d, umma::desc_at(desc, av + ao + 32), umma::desc_at(desc, bl + 32), idesc, 1);
                  umma::mma_tf32_ss(d, umma::desc_at(desc, av + ao + 32), umma::desc_at(desc, bv + 32), idesc, 1);
                }
              }
              if (last_slab) umma::mma_commit(empty_b(rb.stage));
            }
            __syncwarp();
            first = 1;
            if (last_slab) { sb = 0; rb.advance(); } else ++sb;
          }
          if (umma::elect_one()) umma::mma_commit(empty_a(ra.stage));
          __syncwarp();
          ra.advance();
        }
      }
      if (umma::elect_one()) umma::mma_commit(tmem_full(buf));
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    // TMEM -> registers (one pixel per lane) -> bias / LeakyReLU -> 32-channel slice staged in shared memory ->
    // copy-out in which 8 consecutive lanes write one pixel's 128 contiguous bytes (full lines), with the
    // LeakyReLU' mask, the hi/lo split and the upsample / (un-)rotate scatter applied on the way out.
    const ConvDst& d = p.dst;
    const Geom& sg = p.src;
    const int ew = warp - 4;
    float* stage = reinterpret_cast<float*>(smem + p.epi_off) + ew * 32 * kStagePitch;
    __shared__ int s_dst[4][32][4];
    __shared__ int s_meta[4][32];       // bits 0..2: number of destinations, bit 3: write zeros, bits 8..: channel shift
    __shared__ int s_src[4][32];        // source flat pixel (activation lookup with EP_ACT_AT_SRC)
    __shared__ float s_bias[400];
    if (d.flags & EP_BIAS)
      for (int i = threadIdx.x - 128; i < d.cvalid && i < 400; i += 128) s_bias[i] = __ldg(d.bias + i);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    int it = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
      const int buf = it & 1;
      const int um = u / p.n_tiles_n, nt = u % p.n_tiles_n;
      umma::mbar_wait(tmem_full(buf), (it >> 1) & 1, abort_addr, p.error_flag, 4);
      umma::tc_fence_after();
      for (int tile = 0; tile < T; ++tile) {
        const long long j = (long long)um * 128 * T + tile * 128 + ew * 32 + lane;
        const int b = (int)(j / sg.S);
        const int rem = (int)(j - (long long)b * sg.S);
        const int rr = rem / sg.P;
        const int x = rem - rr * sg.P, y = rr - sg.row0;
        const bool valid = (b < sg.B) && (y >= 0) && (x < sg.W);
        if (d.map != MAP_NCHW) {
          int nd = 0, zero = 0, cshift = 0, d0 = 0, d1 = 0, d2 = 0, d3 = 0;
          if (valid) {
            const Geom& dg = d.g;
            if (d.map == MAP_IDENT) {
              d0 = b * dg.S + (y + dg.row0) * dg.P + x; nd = 1;
            } else if (d.map == MAP_UP2) {
              d0 = b * dg.S + (2 * y + dg.row0) * dg.P + 2 * x; d1 = d0 + 1; d2 = d0 + dg.P; d3 = d2 + 1; nd = 4;
            } else if (d.map == MAP_UNROT) {
              const int br = b / d.nimg, n = b - br * d.nimg, H = sg.H, W = sg.W;
              const int pp = (y + 1 == H) ? 0 : y + 1, q = x;
              zero = (y + 1 == H);
              int i, jj;
              if (br == 0) { i = pp; jj = q; } else if (br == 1) { i = q; jj = H - 1 - pp; }
              else if (br == 2) { i = H - 1 - pp; jj = W - 1 - q; } else { i = W - 1 - q; jj = pp; }
              d0 = n * dg.S + (i + dg.row0) * dg.P + jj; nd = 1; cshift = br * d.cvalid;
            } else {   // MAP_UNROT_INV
              const int br = nt, H = sg.H, W = sg.W;
              int pp, q;
              if (br == 0) { pp = y; q = x; } else if (br == 1) { pp = H - 1 - x; q = y; }
              else if (br == 2) { pp = H - 1 - y; q = W - 1 - x; } else { pp = x; q = W - 1 - y; }
              if (pp > 0) { d0 = (br * d.nimg + b) * dg.S + (pp - 1 + dg.row0) * dg.P + q; nd = 1; }
            }
          }
          s_dst[ew][lane][0] = d0; s_dst[ew][lane][1] = d1; s_dst[ew][lane][2] = d2; s_dst[ew][lane][3] = d3;
          s_meta[ew][lane] = nd | (zero << 3) | (cshift << 8);
          s_src[ew][lane] = (int)j;
        }
        const uint32_t trow = tmem + (uint32_t(ew * 32) << 16) + (buf * T + tile) * p.N;
        for (int c0 = 0; c0 < p.N; c0 += 32) {
          const int cw = min(32, p.N - c0);                           // 32 or 16 channels in this slice
          const int cg0 = (d.map == MAP_UNROT_INV ? 0 : nt * p.N) + c0; // first channel of the slice among this conv's outputs
          uint32_t r[32];
          umma::tmem_ld16(trow + c0, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
          if (cw == 32) umma::tmem_ld16(trow + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
          umma::tmem_ld_wait();
          if (cg0 >= d.cvalid) continue;
          if (d.flags & (EP_BIAS | EP_LRELU)) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (i < cw) {
                float a = __uint_as_float(r[i]);
                if ((d.flags & EP_BIAS) && cg0 + i < d.cvalid) a += s_bias[cg0 + i];
                if (d.flags & EP_LRELU) a = lrelu(a);
                r[i] = __float_as_uint(a);
              }
            }
          }
          if (d.map == MAP_NCHW) {
            if (valid) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < cw && cg0 + i < d.cvalid)
                  d.v[(((long long)b * d.cvalid + cg0 + i) * sg.H + y) * sg.W + x] = __uint_as_float(r[i]);
            }
            continue;
          }
          // stage this lane's pixel row
          float* row = stage + lane * kStagePitch;
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            if (i < cw) *reinterpret_cast<uint4*>(row + i) = make_uint4(r[i], r[i + 1], r[i + 2], r[i + 3]);
          __syncwarp();
          if (cw == 32) copy_out_slice<32>(d, stage, s_dst[ew], s_meta[ew], s_src[ew], lane, cg0, nt * p.N + c0);
          else copy_out_slice<16>(d, stage, s_dst[ew], s_meta[ew], s_src[ew], lane, cg0, nt * p.N + c0);
          __syncwarp();
        }
      }
      umma::tc_fence_before();
      umma::mbar_arrive(tmem_empty(buf));
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 3) umma::tmem_dealloc(tmem, 512);
}

}  // namespace convk

// ------------------------------------------------------------------------------------------ host
#include <algorithm>
#include <cstdlib>
#include <vector>

struct ConvTaps { int n; int off[9]; };   // flat-pixel offsets of the taps, in weight-slab order

// Number of 16-channel chunks and k-steps of the last chunk for `cin` input channels.
static inline void conv_chunks(int cin, int* n_chunks, int* ksteps_last) {
  *n_chunks = (cin + 15) / 16;
  const int rem = cin - (*n_chunks - 1) * 16;
  *ksteps_last = rem > 8 ? 2 : 1;
}

// Fills plan->p (everything except tensor maps' base pointers) and the tensor maps.
//   a_v/a_lo : source planes with `a_cpitch` channels per pixel, the conv reads channels [a_coff, a_coff+cin)
//   w_slab   : prepared weight slab [n_tiles_n][n_chunks][ntaps][plane][N][16] (see pw::weight_prep_kernel)
static inline int conv_plan_init(ConvPlan* plan, const Geom& src, const float* a_v, const float* a_lo, int a_cpitch,
                                 int a_coff, int cin, const float* w_slab, int cout_padded, int N,
                                 const ConvTaps& taps, const ConvDst& dst, int* error_flag, int num_sms,
                                 size_t smem_limit = 200 * 1024) {
  ConvParams& p = plan->p;
  p = ConvParams{};
  p.src = src; p.N = N; p.dst = dst; p.error_flag = error_flag;
  p.debug = getenv("SSDN_CONV_DEBUG") ? atoi(getenv("SSDN_CONV_DEBUG")) : 0;
  p.n_tiles_n = cout_padded / N;
  conv_chunks(cin, &p.n_chunks, &p.ksteps_last);
  p.ntaps_total = taps.n;
  p.T = (2 * 2 * N <= 512) ? 2 : 1;
  const long long total = src.total();
  p.n_units_m = (int)((total + 128LL * p.T - 1) / (128LL * p.T));
  // B stage = bg consecutive weight slabs (both planes) loaded by ONE TMA op: a TMA instruction costs ~450 clk of the
  // SM's TMA unit whatever its size (profiles/r01_tma_rate.log), so operands must arrive in few, large boxes.
  const int n_slabs = p.n_chunks * taps.n;
  p.bg = (taps.n % 3 == 0) ? 3 : (n_slabs % 3 == 0 ? 3 : (n_slabs % 2 == 0 ? 2 : 1));
  p.b_stage_bytes = (uint32_t)(p.bg * 2 * N * 64);
  // choose the tap grouping: all taps in one window if it fits in shared memory, else one window per
  // distinct row offset (dy), else one window per tap.
  const int rows_unit = 128 * p.T;
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
    uint32_t plane = (uint32_t)(nbox * box_rows * 64);
    int stages = (mode == 0) ? 2 : 3;
    for (int bst = 4; bst >= 2; --bst) {
      const size_t epi = 4 * 32 * convk::kStagePitch * sizeof(float);
      size_t need = (size_t)stages * 2 * plane + (size_t)bst * p.b_stage_bytes + epi + 2048;
      if (need > smem_limit) continue;
      p.n_groups = (int)gs.size(); p.nbox = nbox; p.box_rows = box_rows; p.a_plane_bytes = plane; p.a_stages = stages; p.b_stages = bst;
      p.epi_off = (uint32_t)((size_t)stages * 2 * plane + (size_t)bst * p.b_stage_bytes);
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
  if (!try_group(0) && !try_group(1) && !try_group(2)) return -10;
  plan->grid = std::min(p.n_units_m * p.n_tiles_n, num_sms);
  // tensor maps.  A: 3-D (channel, flat pixel, plane); the lo plane must follow the hi plane at a constant byte distance.
  const long long plane_stride = (long long)((const char*)a_lo - (const char*)a_v);
  if (plane_stride <= 0 || plane_stride % 16) return -11;
  uint64_t adims[3] = {(uint64_t)cin, (uint64_t)total, 2};
  uint64_t astr[2] = {(uint64_t)a_cpitch * 4, (uint64_t)plane_stride};
  uint32_t abox[3] = {16, (uint32_t)p.box_rows, (uint32_t)(p.nbox == 1 ? 2 : 1)};
  int r;
  if ((r = umma::encode_f32(&plan->a, (void*)(a_v + a_coff), 3, adims, astr, abox, CU_TENSOR_MAP_SWIZZLE_64B))) return r;
  // B: slabs [n_tile][chunk][tap][plane][N][16] -> 3-D (16, N, slab x plane), box = bg slabs x 2 planes
  uint64_t bdims[3] = {16, (uint64_t)N, (uint64_t)p.n_tiles_n * n_slabs * 2};
  uint64_t bstr[2] = {64, (uint64_t)N * 64};
  uint32_t bbox[3] = {16, (uint32_t)N, (uint32_t)(2 * p.bg)};
  if ((r = umma::encode_f32(&plan->b, (void*)w_slab, 3, bdims, bstr, bbox, CU_TENSOR_MAP_SWIZZLE_64B))) return r;
  return 0;
}

// floats of the combined (hi + lo) weight slab
static inline size_t conv_weight_slab_floats(int cin, int cout_padded, int ntaps) {
  int nc, kl; conv_chunks(cin, &nc, &kl);
  return (size_t)cout_padded * nc * ntaps * 16 * 2;
}

// Optional per-launch timing (bench.py roofline): when enabled every GEMM launch is bracketed by CUDA events.
struct LaunchProfiler {
  bool on = false;
  struct Rec { int kind; double flops; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  void begin(int kind, double flops, cudaStream_t st) { Rec r{kind, flops, nullptr, nullptr}; cudaEventCreate(&r.a); cudaEventCreate(&r.b); cudaEventRecord(r.a, st); recs.push_back(r); }
  void end(cudaStream_t st) { cudaEventRecord(recs.back().b, st); }
};
inline LaunchProfiler& profiler() { static LaunchProfiler p; return p; }

static inline cudaError_t conv_launch(const ConvPlan& plan, cudaStream_t stream, int kind = 0) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(convk::conv_igemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(convk::conv_igemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (profiler().on) profiler().begin(kind, plan.flops, stream);
  if (plan.p.T == 2) convk::conv_igemm_kernel<2><<<plan.grid, convk::kThreads, plan.smem, stream>>>(plan.a, plan.b, plan.p);
  else convk::conv_igemm_kernel<1><<<plan.grid, convk::kThreads, plan.smem, stream>>>(plan.a, plan.b, plan.p);
  if (profiler().on) profiler().end(stream);
  return cudaGetLastError();
}
