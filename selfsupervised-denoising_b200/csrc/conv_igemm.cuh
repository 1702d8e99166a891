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
  int nbox, box_rows; // every group window is loaded as nbox TMA boxes of box_rows rows
  int a_stages, b_stages;
  uint32_t a_plane_bytes, b_plane_bytes;   // smem bytes of one plane of one stage (1024-aligned)
  ConvDst dst;
  int* error_flag;
};

struct ConvPlan {
  ConvParams p;
  double flops = 0;   // algorithmic 2*MAC of this launch (valid pixels, real channels); filled by the owner
  CUtensorMap a_v, a_lo, b_v, b_lo;
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

// Writes 16 consecutive output channels [cbase, cbase+16) of one destination pixel.
//   act_index: flat float index of the forward activation for channel cbase (EP_ACT_GRAD), else unused.
__device__ __forceinline__ void store_pixel(const ConvDst& d, long long dflat, int coff, int cbase, long long act_index,
                                            const float (&val)[16], bool write_zero) {
  float out[16];
  const bool full = (cbase + 16 <= d.cvalid);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float x = val[i];
    if (d.flags & EP_ACT_GRAD) {
      float a = (full || cbase + i < d.cvalid) ? __ldg(d.act + act_index + i) : 1.f;
      x = a > 0.f ? x : SSDN_LRELU_SLOPE * x;
    }
    out[i] = write_zero ? 0.f : x;
  }
  float* pv = d.v + dflat * d.cpitch + coff + cbase;
  float* pl = d.lo + dflat * d.cpitch + coff + cbase;
  float lo[16];
  if (d.flags & EP_WRITE_LO) {
#pragma unroll
    for (int i = 0; i < 16; ++i) { float h; tf32_split(out[i], h, lo[i]); out[i] = h; }
  }
  if (full) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      *reinterpret_cast<float4*>(pv + i) = make_float4(out[i], out[i + 1], out[i + 2], out[i + 3]);
      if (d.flags & EP_WRITE_LO) *reinterpret_cast<float4*>(pl + i) = make_float4(lo[i], lo[i + 1], lo[i + 2], lo[i + 3]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (cbase + i < d.cvalid) {
        pv[i] = out[i];
        if (d.flags & EP_WRITE_LO) pl[i] = lo[i];
      }
  }
}

__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_a_v, const __grid_constant__ CUtensorMap map_a_lo,
                  const __grid_constant__ CUtensorMap map_b_v, const __grid_constant__ CUtensorMap map_b_lo,
                  const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[4 * kMaxStages + 4];
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = umma::smem_u32(smem);
  const uint32_t a_stage_bytes = 2 * p.a_plane_bytes, b_stage_bytes = 2 * p.b_plane_bytes;
  const uint32_t a_base = sbase, b_base = sbase + p.a_stages * a_stage_bytes;
  auto full_a = [&](int s) { return umma::smem_u32(&bars[s]); };
  auto empty_a = [&](int s) { return umma::smem_u32(&bars[kMaxStages + s]); };
  auto full_b = [&](int s) { return umma::smem_u32(&bars[2 * kMaxStages + s]); };
  auto empty_b = [&](int s) { return umma::smem_u32(&bars[3 * kMaxStages + s]); };
  auto tmem_full = [&](int b) { return umma::smem_u32(&bars[4 * kMaxStages + b]); };
  auto tmem_empty = [&](int b) { return umma::smem_u32(&bars[4 * kMaxStages + 2 + b]); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_units = p.n_units_m * p.n_tiles_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.a_stages; ++s) { umma::mbar_init(full_a(s), 1); umma::mbar_init(empty_a(s), 1); }
    for (int s = 0; s < p.b_stages; ++s) { umma::mbar_init(full_b(s), 1); umma::mbar_init(empty_b(s), 1); }
    for (int b = 0; b < 2; ++b) { umma::mbar_init(tmem_full(b), 1); umma::mbar_init(tmem_empty(b), 128); }
    umma::fence_mbar_init();
  }
  if (warp == 3) {
    umma::tmem_alloc(umma::smem_u32(&tmem_slot), 512);
    umma::tmem_relinquish();
  }
  if (warp == 0 && lane == 0) { umma::tma_prefetch_desc(&map_a_v); umma::tma_prefetch_desc(&map_a_lo); }
  if (warp == 1 && lane == 0) { umma::tma_prefetch_desc(&map_b_v); umma::tma_prefetch_desc(&map_b_lo); }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  bool ok = true;

  if (warp == 0) {
    // ------------------------------------------------------------ A producer
    if (lane == 0) {
      Ring ra(p.a_stages);
      const uint32_t box_bytes = p.box_rows * 64;
      for (int u = blockIdx.x; u < n_units && ok; u += gridDim.x) {
        const int um = u / p.n_tiles_n;
        const long long j0 = (long long)um * 128 * p.T;
        for (int ch = 0; ch < p.n_chunks && ok; ++ch)
          for (int g = 0; g < p.n_groups; ++g) {
            if (!(ok = umma::mbar_wait(empty_a(ra.stage), ra.phase ^ 1))) break;
            umma::mbar_expect_tx(full_a(ra.stage), 2 * p.nbox * box_bytes);
            const uint32_t dst = a_base + ra.stage * a_stage_bytes;
            const int row = (int)(j0 + p.groups[g].row_off);
            for (int bx = 0; bx < p.nbox; ++bx) {
              umma::tma_load_2d(dst + bx * box_bytes, &map_a_v, full_a(ra.stage), ch * 16, row + bx * p.box_rows);
              umma::tma_load_2d(dst + p.a_plane_bytes + bx * box_bytes, &map_a_lo, full_a(ra.stage), ch * 16,
                                row + bx * p.box_rows);
            }
            ra.advance();
          }
      }
      if (!ok) atomicExch(p.error_flag, 1);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ B producer
    if (lane == 0) {
      Ring rb(p.b_stages);
      const uint32_t tile_bytes = p.N * 64;
      for (int u = blockIdx.x; u < n_units && ok; u += gridDim.x) {
        const int nt = u % p.n_tiles_n;
        for (int ch = 0; ch < p.n_chunks && ok; ++ch)
          for (int g = 0; g < p.n_groups && ok; ++g)
            for (int t = 0; t < p.groups[g].ntaps; ++t) {
              if (!(ok = umma::mbar_wait(empty_b(rb.stage), rb.phase ^ 1))) break;
              umma::mbar_expect_tx(full_b(rb.stage), 2 * tile_bytes);
              const uint32_t dst = b_base + rb.stage * b_stage_bytes;
              const int row = ((nt * p.n_chunks + ch) * p.ntaps_total + p.groups[g].tap_id[t]) * p.N;
              umma::tma_load_2d(dst, &map_b_v, full_b(rb.stage), 0, row);
              umma::tma_load_2d(dst + p.b_plane_bytes, &map_b_lo, full_b(rb.stage), 0, row);
              rb.advance();
            }
      }
      if (!ok) atomicExch(p.error_flag, 2);
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      Ring ra(p.a_stages), rb(p.b_stages);
      const uint64_t desc = umma::make_desc_base(16, 512, umma::LAYOUT_SW64);
      const uint32_t idesc = umma::make_idesc_tf32(128, p.N, 0, 0);
      int it = 0;
      for (int u = blockIdx.x; u < n_units && ok; u += gridDim.x, ++it) {
        const int buf = it & 1;
        if (!(ok = umma::mbar_wait(tmem_empty(buf), ((it >> 1) & 1) ^ 1))) break;
        umma::tc_fence_after();
        bool first = true;
        for (int ch = 0; ch < p.n_chunks && ok; ++ch) {
          const int ks = (ch == p.n_chunks - 1) ? p.ksteps_last : 2;
          for (int g = 0; g < p.n_groups && ok; ++g) {
            if (!(ok = umma::mbar_wait(full_a(ra.stage), ra.phase))) break;
            const uint32_t av = a_base + ra.stage * a_stage_bytes, al = av + p.a_plane_bytes;
            for (int t = 0; t < p.groups[g].ntaps; ++t) {
              if (!(ok = umma::mbar_wait(full_b(rb.stage), rb.phase))) break;
              umma::tc_fence_after();
              const uint32_t bv = b_base + rb.stage * b_stage_bytes, bl = bv + p.b_plane_bytes;
              const uint32_t rel = p.groups[g].tap_rel[t] * 64;
              for (int tile = 0; tile < p.T; ++tile) {
                const uint32_t d = tmem + (buf * p.T + tile) * p.N;
                for (int k = 0; k < ks; ++k) {
                  const uint32_t ao = rel + tile * (128 * 64) + k * 32, bo = k * 32;
                  umma::mma_tf32_ss(d, umma::desc_at(desc, al + ao), umma::desc_at(desc, bv + bo), idesc,
                                    !(first && k == 0));
                  umma::mma_tf32_ss(d, umma::desc_at(desc, av + ao), umma::desc_at(desc, bl + bo), idesc, 1);
                  umma::mma_tf32_ss(d, umma::desc_at(desc, av + ao), umma::desc_at(desc, bv + bo), idesc, 1);
                }
              }
              first = false;
              umma::mma_commit(empty_b(rb.stage));
              rb.advance();
            }
            umma::mma_commit(empty_a(ra.stage));
            ra.advance();
          }
        }
        umma::mma_commit(tmem_full(buf));
      }
      if (!ok) atomicExch(p.error_flag, 3);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const ConvDst& d = p.dst;
    const Geom& sg = p.src;
    const int ew = warp - 4;
    int it = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
      const int buf = it & 1;
      const int um = u / p.n_tiles_n, nt = u % p.n_tiles_n;
      if (!umma::mbar_wait(tmem_full(buf), (it >> 1) & 1)) { ok = false; }
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) { if (lane == 0) atomicExch(p.error_flag, 4); break; }
      umma::tc_fence_after();
      for (int tile = 0; tile < p.T; ++tile) {
        const long long j = (long long)um * 128 * p.T + tile * 128 + ew * 32 + lane;
        const int b = (int)(j / sg.S);
        const int rem = (int)(j - (long long)b * sg.S);
        const int rr = rem / sg.P;
        const int x = rem - rr * sg.P, y = rr - sg.row0;
        const bool valid = (b < sg.B) && (y >= 0) && (x < sg.W);
        // destination pixel(s)
        long long dflat[4]; int ndst = 0; bool zero = false; int cshift = 0;
        if (valid) {
          const Geom& dg = d.g;
          if (d.map == MAP_IDENT) {
            dflat[0] = (long long)b * dg.S + (y + dg.row0) * dg.P + x; ndst = 1;
          } else if (d.map == MAP_UP2) {
            const long long o = (long long)b * dg.S + (2 * y + dg.row0) * dg.P + 2 * x;
            dflat[0] = o; dflat[1] = o + 1; dflat[2] = o + dg.P; dflat[3] = o + dg.P + 1; ndst = 4;
          } else if (d.map == MAP_UNROT) {
            const int br = b / d.nimg, n = b - br * d.nimg, H = sg.H, W = sg.W;
            const int pp = (y + 1 == H) ? 0 : y + 1, q = x;
            zero = (y + 1 == H);
            int i, jj;
            if (br == 0) { i = pp; jj = q; } else if (br == 1) { i = q; jj = H - 1 - pp; }
            else if (br == 2) { i = H - 1 - pp; jj = W - 1 - q; } else { i = W - 1 - q; jj = pp; }
            dflat[0] = (long long)n * dg.S + (i + dg.row0) * dg.P + jj; ndst = 1; cshift = br * d.cvalid;
          } else if (d.map == MAP_UNROT_INV) {
            const int br = nt, H = sg.H, W = sg.W;
            int pp, q;
            if (br == 0) { pp = y; q = x; } else if (br == 1) { pp = H - 1 - x; q = y; }
            else if (br == 2) { pp = H - 1 - y; q = W - 1 - x; } else { pp = x; q = W - 1 - y; }
            if (pp > 0) { dflat[0] = (long long)(br * d.nimg + b) * dg.S + (pp - 1 + dg.row0) * dg.P + q; ndst = 1; }
          } else {
            ndst = 1;
          }
        }
        const uint32_t trow = tmem + (uint32_t(ew * 32) << 16) + (buf * p.T + tile) * p.N;
        for (int c0 = 0; c0 < p.N; c0 += 16) {
          uint32_t r[16];
          umma::tmem_ld16(trow + c0, r);
          umma::tmem_ld_wait();
          if (ndst == 0) continue;
          const int cg = (d.map == MAP_UNROT_INV ? 0 : nt * p.N) + c0;   // channel within this conv's outputs
          if (cg >= d.cvalid) continue;
          float val[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float a = __uint_as_float(r[i]);
            if ((d.flags & EP_BIAS) && cg + i < d.cvalid) a += __ldg(d.bias + cg + i);
            if (d.flags & EP_LRELU) a = lrelu(a);
            val[i] = a;
          }
          if (d.map == MAP_NCHW) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (cg + i < d.cvalid)
                d.v[(((long long)b * d.cvalid + cg + i) * sg.H + y) * sg.W + x] = val[i];
          } else {
            for (int k = 0; k < ndst; ++k) {
              // forward activation for the LeakyReLU' mask: at the destination pixel, or (EP_ACT_AT_SRC) at the
              // source pixel with the GEMM's true output channel (head input holds the un-rotated branch outputs)
              const long long ai = (d.flags & EP_ACT_AT_SRC) ? j * d.act_cpitch + d.act_coff + nt * p.N + c0
                                                             : dflat[k] * d.act_cpitch + d.act_coff + cg;
              store_pixel(d, dflat[k], d.coff + cshift, cg, ai, val, zero);
            }
          }
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
//   w_v/w_lo : prepared weight slab [n_tiles_n][n_chunks][ntaps][N][16]
static inline int conv_plan_init(ConvPlan* plan, const Geom& src, const float* a_v, const float* a_lo, int a_cpitch,
                                 int a_coff, int cin, const float* w_v, const float* w_lo, int cout_padded, int N,
                                 const ConvTaps& taps, const ConvDst& dst, int* error_flag, int num_sms,
                                 size_t smem_limit = 200 * 1024) {
  ConvParams& p = plan->p;
  p = ConvParams{};
  p.src = src; p.N = N; p.dst = dst; p.error_flag = error_flag;
  p.n_tiles_n = cout_padded / N;
  conv_chunks(cin, &p.n_chunks, &p.ksteps_last);
  p.ntaps_total = taps.n;
  p.T = (2 * 2 * N <= 512) ? 2 : 1;
  const long long total = src.total();
  p.n_units_m = (int)((total + 128LL * p.T - 1) / (128LL * p.T));
  p.b_stages = 6;
  p.b_plane_bytes = (uint32_t)((N * 64 + 1023) / 1024 * 1024);
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
    int box_rows = ((max_rows + nbox - 1) / nbox + 7) / 8 * 8;
    uint32_t plane = (uint32_t)((nbox * box_rows * 64 + 1023) / 1024 * 1024);
    int stages = (mode == 0) ? 2 : 3;
    size_t need = (size_t)stages * 2 * plane + (size_t)p.b_stages * 2 * p.b_plane_bytes + 1024;
    if (need > smem_limit) return false;
    p.n_groups = (int)gs.size(); p.nbox = nbox; p.box_rows = box_rows; p.a_plane_bytes = plane; p.a_stages = stages;
    for (size_t gi = 0; gi < gs.size(); ++gi) {
      int lo = INT32_MAX;
      for (int t : gs[gi]) lo = std::min(lo, taps.off[t]);
      TapGroup& tg = p.groups[gi];
      tg.row_off = lo; tg.ntaps = (int)gs[gi].size();
      for (size_t k = 0; k < gs[gi].size(); ++k) { tg.tap_id[k] = gs[gi][k]; tg.tap_rel[k] = taps.off[gs[gi][k]] - lo; }
    }
    plan->smem = need;
    return true;
  };
  if (!try_group(0) && !try_group(1) && !try_group(2)) return -10;
  plan->grid = std::min(p.n_units_m * p.n_tiles_n, num_sms);
  // tensor maps
  uint64_t adims[2] = {(uint64_t)cin, (uint64_t)total};
  uint64_t astr[1] = {(uint64_t)a_cpitch * 4};
  uint32_t abox[2] = {16, (uint32_t)p.box_rows};
  int r;
  if ((r = umma::encode_f32(&plan->a_v, (void*)(a_v + a_coff), 2, adims, astr, abox, CU_TENSOR_MAP_SWIZZLE_64B))) return r;
  if ((r = umma::encode_f32(&plan->a_lo, (void*)(a_lo + a_coff), 2, adims, astr, abox, CU_TENSOR_MAP_SWIZZLE_64B))) return r;
  uint64_t bdims[2] = {16, (uint64_t)p.n_tiles_n * p.n_chunks * taps.n * N};
  uint64_t bstr[1] = {64};
  uint32_t bbox[2] = {16, (uint32_t)N};
  if ((r = umma::encode_f32(&plan->b_v, (void*)w_v, 2, bdims, bstr, bbox, CU_TENSOR_MAP_SWIZZLE_64B))) return r;
  if ((r = umma::encode_f32(&plan->b_lo, (void*)w_lo, 2, bdims, bstr, bbox, CU_TENSOR_MAP_SWIZZLE_64B))) return r;
  return 0;
}

static inline size_t conv_weight_slab_floats(int cin, int cout_padded, int ntaps) {
  int nc, kl; conv_chunks(cin, &nc, &kl);
  return (size_t)cout_padded * nc * ntaps * 16;
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
    cudaError_t e = cudaFuncSetAttribute(convk::conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (profiler().on) profiler().begin(kind, plan.flops, stream);
  convk::conv_igemm_kernel<<<plan.grid, convk::kThreads, plan.smem, stream>>>(plan.a_v, plan.a_lo, plan.b_v, plan.b_lo, plan.p);
  if (profiler().on) profiler().end(stream);
  return cudaGetLastError();
}
