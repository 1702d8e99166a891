// HBM-bound kernels of the U-Net: layout packing (with the 4-rotation stack), shifted max-pool
// forward/backward, upsample backward, weight slab preparation, bias gradients, operand-scale bookkeeping.
// All of them stream each byte once; operand planes are fp16 (hi, lo) pairs read / written as 16-byte vectors of 8 channels
// (common.cuh: values carry the tensor's power-of-two scale), raw gradients are fp32 read as float4.
// Index work is exact integer arithmetic (bit-exact versus the oracle).
#pragma once
#include "common.cuh"

namespace pw {

constexpr int kBlock = 256;
static inline int grid_for(long long n, int block = kBlock) { return (int)((n + block - 1) / block); }

// ---------------------------------------------------------------------------- NCHW -> padded flat
// rot4 != 0: the output holds 4*B images, image (r*B + b) = rotate(x[b], 90*r)   (utils/data.py:42-67)
// One thread per element (tests, single-operator entry points).
__global__ void pack_nchw_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo,
                                 int B, int C, int H, int W, Geom g, int cpitch, int coff, int rot4, ScaleRef sc) {
  pdl_wait();
  const long long n = (long long)(rot4 ? 4 : 1) * B * C * H * W;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const float s = sc.k ? exp2_int(__ldg(sc.k)) : 1.0f;
  float m = 0.f;
  if (idx < n) {
    const int j = (int)(idx % W); long long t = idx / W;
    const int i = (int)(t % H); t /= H;
    const int c = (int)(t % C); const int bo = (int)(t / C);
    const int r = rot4 ? bo / B : 0, b = rot4 ? bo % B : bo;
    int si = i, sj = j;
    if (r == 1) { si = j; sj = W - 1 - i; } else if (r == 2) { si = H - 1 - i; sj = W - 1 - j; }
    else if (r == 3) { si = H - 1 - j; sj = i; }
    const float val = __ldg(x + (((long long)b * C + c) * H + si) * W + sj);
    const long long o = ((long long)bo * g.S + (i + g.row0) * g.P + j) * cpitch + coff + c;
    f16_split1(val * s, hi[o], lo[o]);
    m = fabsf(val);
  }
  amax_commit(sc.amax, m);
}
// the same mapping into a single fp32 plane (raw gradient buffers; test hook)
__global__ void pack_nchw_f32_kernel(const float* __restrict__ x, float* __restrict__ v, int B, int C, int H, int W, Geom g, int cpitch, int coff) {
  pdl_wait();
  const long long n = (long long)B * C * H * W;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int j = (int)(idx % W); long long t = idx / W;
  const int i = (int)(t % H); t /= H;
  const int c = (int)(t % C); const int b = (int)(t / C);
  v[((long long)b * g.S + (i + g.row0) * g.P + j) * cpitch + coff + c] = __ldg(x + idx);
}

// Same mapping, one thread per output PIXEL writing all C (<= 16) channels as 16-byte vectors of 8 channels per plane (the
// destination's channel offset and pitch are multiples of 8; the pad channels of the last vector are written as zeros): the
// index arithmetic is done once per pixel (the network input and the loss gradient have 3..12 channels).
__global__ void pack_nchw_pixel_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo,
                                       int B, int C, int H, int W, Geom g, int cpitch, int coff, int rot4, ScaleRef sc) {
  pdl_wait();
  const long long n = (long long)(rot4 ? 4 : 1) * B * H * W;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const float s = sc.k ? exp2_int(__ldg(sc.k)) : 1.0f;
  float m = 0.f;
  if (idx < n) {
    const int j = (int)(idx % W); long long t = idx / W;
    const int i = (int)(t % H); const int bo = (int)(t / H);
    const int r = rot4 ? bo / B : 0, b = rot4 ? bo % B : bo;
    int si = i, sj = j;
    if (r == 1) { si = j; sj = W - 1 - i; } else if (r == 2) { si = H - 1 - i; sj = W - 1 - j; }
    else if (r == 3) { si = H - 1 - j; sj = i; }
    const float* src = x + ((long long)b * C * H + si) * W + sj;
    const long long o = ((long long)bo * g.S + (i + g.row0) * g.P + j) * cpitch + coff;
    for (int c0 = 0; c0 < C; c0 += 8) {
      float f[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float val = (c0 + q < C) ? __ldg(src + (long long)(c0 + q) * H * W) : 0.f;
        m = fmaxf(m, fabsf(val));
        f[q] = val * s;
      }
      uint4 h, l;
      f16_split8(f, h, l);
      *reinterpret_cast<uint4*>(hi + o + c0) = h; *reinterpret_cast<uint4*>(lo + o + c0) = l;
    }
  }
  amax_commit(sc.amax, m);
}

// The FIRST convolution runs as a 1x1 GEMM over an im2col'd input (net.cuh): xcol[flat pixel][k], k = c * 9 + kh * 3 + kw (the
// flat order of a PyTorch [co][ci][kh][kw] weight row), holds x(c, i + kh - sh, j + kw - 1) of the (rotated) image and zero
// outside it (sh = 2: half-plane ShiftConv2d, 1: plain conv).  With 3 input channels a stencil tap is a K = 3 sliver of a
// K = 16 MMA: nine taps cost nine MMAs per k-step and product where one N = 32 (weight gradient) / one chunk (forward) now do.
// One thread = 8 consecutive k of one pixel (one 16-byte store per plane); the gathers hit L1 / L2 (the input is a few MB).
__global__ void pack_im2col3x3_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo,
                                      int B, int C, int H, int W, Geom g, int cpitch, int sh, int rot4, ScaleRef sc) {
  pdl_wait();
  const int groups = cpitch >> 3;
  const long long n = (long long)(rot4 ? 4 : 1) * B * H * W * groups;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const float s = sc.k ? exp2_int(__ldg(sc.k)) : 1.0f;
  float m = 0.f;
  if (idx < n) {
    const int q = (int)(idx % groups); long long t = idx / groups;
    const int j = (int)(t % W); t /= W;
    const int i = (int)(t % H); const int bo = (int)(t / H);
    const int r = rot4 ? bo / B : 0, b = rot4 ? bo % B : bo;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = 8 * q + e;
      float val = 0.f;
      if (k < 9 * C) {
        const int c = k / 9, tap = k - 9 * c, kh = tap / 3, kw = tap - 3 * kh;
        const int ii = i + kh - sh, jj = j + kw - 1;
        if (ii >= 0 && ii < H && jj >= 0 && jj < W) {
          int si = ii, sj = jj;
          if (r == 1) { si = jj; sj = W - 1 - ii; } else if (r == 2) { si = H - 1 - ii; sj = W - 1 - jj; }
          else if (r == 3) { si = H - 1 - jj; sj = ii; }
          val = __ldg(x + (((long long)b * C + c) * H + si) * W + sj);
        }
      }
      m = fmaxf(m, fabsf(val));
      f[e] = val * s;
    }
    uint4 h, l;
    f16_split8(f, h, l);
    const long long o = ((long long)bo * g.S + (i + g.row0) * g.P + j) * cpitch + 8 * q;
    *reinterpret_cast<uint4*>(hi + o) = h; *reinterpret_cast<uint4*>(lo + o) = l;
  }
  amax_commit(sc.amax, m);
}

// Network input in ONE pass: the (rotated) NCHW image goes (a) into its channel slot of the last concat buffer and (b) into the
// im2col operand of the first convolution (layout of pack_im2col3x3_kernel above).  A block owns a 16 x 32 tile of one rotated
// image: it stages the tile + stencil halo of every channel in shared memory with reads that are coalesced in the SOURCE
// image whatever the rotation (odd rotations walk the tile column-major), then writes whole 16-byte channel groups, consecutive
// threads = consecutive (pixel, group).  The two thread-per-pixel kernels this replaces took 95 us of the step for 77 MB
// (the rotated branches read one sector per element).
constexpr int kPackTileH = 16, kPackTileW = 32, kPackRegH = kPackTileH + 2, kPackRegW = kPackTileW + 2, kPackRegPitch = kPackRegW + 1;
__global__ void __launch_bounds__(256) pack_input_kernel(const float* __restrict__ x, __half* __restrict__ c_hi, __half* __restrict__ c_lo, int c_pitch, int c_off,
                                                         ScaleRef c_sc, __half* __restrict__ k_hi, __half* __restrict__ k_lo, int k_pitch, ScaleRef k_sc,
                                                         int B, int C, int H, int W, Geom g, int sh, int rot4) {
  extern __shared__ float pk_tile[];                 // [C][kPackRegH][kPackRegPitch]
  pdl_wait();
  const int tiles_x = (W + kPackTileW - 1) / kPackTileW, tiles_y = (H + kPackTileH - 1) / kPackTileH;
  int t = blockIdx.x;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y; const int bo = t / tiles_y;
  const int r = rot4 ? bo / B : 0, b = rot4 ? bo % B : bo;
  const int i0 = ty * kPackTileH, j0 = tx * kPackTileW;
  constexpr int kRegElems = kPackRegH * kPackRegW;
  const int n_reg = C * kRegElems;
  for (int e0 = threadIdx.x; e0 < n_reg; e0 += 4 * blockDim.x) {
    float v[4]; int so[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {          // four independent loads in flight per thread
      const int e = e0 + u * blockDim.x;
      v[u] = 0.f; so[u] = -1;
      if (e < n_reg) {
        const int c = e / kRegElems, rem = e - c * kRegElems;
        int ri, rj;
        if (r & 1) { rj = rem / kPackRegH; ri = rem - rj * kPackRegH; } else { ri = rem / kPackRegW; rj = rem - ri * kPackRegW; }
        const int ii = i0 - sh + ri, jj = j0 - 1 + rj;
        so[u] = (c * kPackRegH + ri) * kPackRegPitch + rj;
        if (ii >= 0 && ii < H && jj >= 0 && jj < W) {
          int si = ii, sj = jj;
          if (r == 1) { si = jj; sj = W - 1 - ii; } else if (r == 2) { si = H - 1 - ii; sj = W - 1 - jj; }
          else if (r == 3) { si = H - 1 - jj; sj = ii; }
          v[u] = __ldg(x + (((long long)b * C + c) * H + si) * W + sj);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) if (so[u] >= 0) pk_tile[so[u]] = v[u];
  }
  __syncthreads();
  const float s_c = c_sc.k ? exp2_int(__ldg(c_sc.k)) : 1.0f, s_k = k_sc.k ? exp2_int(__ldg(k_sc.k)) : 1.0f;
  // whole 32-byte sectors when the slot allows it: a 16-byte write into a sector L2 does not hold costs a DRAM read to fill it
  const int groups_k = k_hi ? (k_pitch >> 3) : 0, groups_c = (c_pitch - c_off >= ((C + 15) & ~15) && (c_off & 15) == 0) ? 2 * ((C + 15) >> 4) : (C + 7) >> 3;
  float m = 0.f;
  // A thread keeps ONE channel group for the whole tile (the loop strides are multiples of the group counts), so the tile
  // offsets of its 8 values are computed once; an item is then 8 shared-memory reads, the split and two 16-byte stores.
  auto run = [&](int ng, bool col, float scale, __half* __restrict__ d_hi, __half* __restrict__ d_lo, int pitch, int off0) {
    const int active = ((int)blockDim.x / ng) * ng;
    if ((int)threadIdx.x >= active) return;
    const int q = threadIdx.x % ng;
    int off[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = 8 * q + e;
      if (col) { const int c = k / 9, tap = k - 9 * c, kh = tap / 3, kw = tap - 3 * kh; off[e] = k < 9 * C ? (c * kPackRegH + kh) * kPackRegPitch + kw : -1; }
      else off[e] = k < C ? (k * kPackRegH + sh) * kPackRegPitch + 1 : -1;
    }
    for (int px = threadIdx.x / ng; px < kPackTileH * kPackTileW; px += active / ng) {
      const int pi = px / kPackTileW, pj = px - pi * kPackTileW;
      const int i = i0 + pi, j = j0 + pj;
      if (i >= H || j >= W) continue;
      const float* base = pk_tile + pi * kPackRegPitch + pj;
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float val = off[e] >= 0 ? base[off[e]] : 0.f;
        m = fmaxf(m, fabsf(val));
        f[e] = val * scale;
      }
      uint4 h, l;
      f16_split8(f, h, l);
      const long long o = ((long long)bo * g.S + (i + g.row0) * g.P + j) * pitch + off0 + 8 * q;
      *reinterpret_cast<uint4*>(d_hi + o) = h; *reinterpret_cast<uint4*>(d_lo + o) = l;
    }
  };
  if (groups_k) run(groups_k, true, s_k, k_hi, k_lo, k_pitch, 0);
  run(groups_c, false, s_c, c_hi, c_lo, c_pitch, c_off);
  // both tensors hold the same values: one maximum serves both slots
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) {
    if (c_sc.amax) atomicMax(c_sc.amax, __float_as_uint(m));
    if (k_sc.amax) atomicMax(k_sc.amax, __float_as_uint(m));
  }
}

// padded flat -> dense NCHW (tests / debugging): plane 0: value = (hi + lo) * 2^-k, 1: lo, 2: hi (as stored, scaled)
__global__ void unpack_nchw_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, const int* __restrict__ k, int plane,
                                   float* __restrict__ y, int B, int C, int H, int W, Geom g, int cpitch, int coff) {
  pdl_wait();
  const long long n = (long long)B * C * H * W;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int j = (int)(idx % W); long long t = idx / W;
  const int i = (int)(t % H); t /= H;
  const int c = (int)(t % C); const int b = (int)(t / C);
  const long long o = ((long long)b * g.S + (i + g.row0) * g.P + j) * cpitch + coff + c;
  if (plane == 0) y[idx] = f16_join1(hi[o], lo[o]) * exp2_int(k ? -__ldg(k) : 0);
  else y[idx] = __half2float(plane == 1 ? lo[o] : hi[o]);
}
__global__ void unpack_nchw_f32_kernel(const float* __restrict__ v, float* __restrict__ y, int B, int C, int H, int W, Geom g, int cpitch, int coff) {
  pdl_wait();
  const long long n = (long long)B * C * H * W;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int j = (int)(idx % W); long long t = idx / W;
  const int i = (int)(t % H); t /= H;
  const int c = (int)(t % C); const int b = (int)(t / C);
  y[idx] = v[((long long)b * g.S + (i + g.row0) * g.P + j) * cpitch + coff + c];
}

// LeakyReLU sign masks of a two-plane tensor (test hook: after an activation buffer was overwritten from outside).
__global__ void mask_from_planes_kernel(const __half* __restrict__ hi, long long pixels, int cpitch, uint32_t* __restrict__ mask, int words) {
  pdl_wait();
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= pixels * words) return;
  const long long px = idx / words; const int w = (int)(idx - px * words);
  uint32_t m = 0;
  for (int i = 0; i < 32 && 32 * w + i < cpitch; ++i) m |= (__half2float(hi[px * cpitch + 32 * w + i]) > 0.f ? 1u : 0u) << i;
  mask[idx] = m;
}

// ---------------------------------------------------------------------------- max-pool 2x2
// 8 channels of both planes at element offset i -> hi + lo (scaled)
__device__ __forceinline__ void ld8(const __half* __restrict__ hi, const __half* __restrict__ lo, long long i, float (&f)[8]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(hi + i)), b = __ldg(reinterpret_cast<const uint4*>(lo + i));
  f16_join8(a, b, f);
}
__device__ __forceinline__ void st8(__half* __restrict__ hi, __half* __restrict__ lo, long long i, const float (&f)[8]) {
  uint4 h, l;
  f16_split8(f, h, l);
  *reinterpret_cast<uint4*>(hi + i) = h; *reinterpret_cast<uint4*>(lo + i) = l;
}

// blind != 0: Shift2d((1,0)) then MaxPool2d(2)  (models/noise_network.py:64-67): window rows (2i-1, 2i),
// row -1 is the zero halo row of the padded layout.  One thread = one output pixel x 8 channels.
__global__ void pool_fwd_kernel(const __half* __restrict__ src, const __half* __restrict__ src_lo, Geom gs, int s_cpitch, int s_coff, ScaleRef ssc,
                                __half* __restrict__ dv, __half* __restrict__ dlo, Geom gd, int d_cpitch, int d_coff, ScaleRef dsc,
                                int C, int blind) {
  pdl_wait();
  const int c8n = C / 8;
  const long long n = (long long)gd.B * gd.H * gd.W * c8n;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int ks = ssc.k ? __ldg(ssc.k) : 0, kd = dsc.k ? __ldg(dsc.k) : 0;
  const float rescale = exp2_int(kd - ks), unscale = exp2_int(-ks);
  float mx = 0.f;
  if (idx < n) {
    const int c = (int)(idx % c8n) * 8; long long t = idx / c8n;
    const int xo = (int)(t % gd.W); t /= gd.W;
    const int yo = (int)(t % gd.H); const int b = (int)(t / gd.H);
    const int y0 = 2 * yo - (blind ? 1 : 0);
    const long long s0 = ((long long)b * gs.S + (y0 + gs.row0) * gs.P + 2 * xo) * s_cpitch + s_coff + c;
    float a[8], bq[8], cq[8], dq[8], m[8];
    ld8(src, src_lo, s0, a); ld8(src, src_lo, s0 + s_cpitch, bq);
    ld8(src, src_lo, s0 + (long long)gs.P * s_cpitch, cq); ld8(src, src_lo, s0 + (long long)(gs.P + 1) * s_cpitch, dq);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float v = fmaxf(fmaxf(a[i], bq[i]), fmaxf(cq[i], dq[i]));
      mx = fmaxf(mx, fabsf(v));
      m[i] = v * rescale;
    }
    st8(dv, dlo, ((long long)b * gd.S + (yo + gd.row0) * gd.P + xo) * d_cpitch + d_coff + c, m);
  }
  amax_commit(dsc.amax, mx * unscale);
}

constexpr int kFusedColsumBlock = 240;    // multiple of 6 and 12 (8-channel groups of 48 / 96 channels)
constexpr int kFusedColsumGrid = 592;

// Backward of [LeakyReLU -> (shift) -> max-pool]: routes g = ga (+ gb) to the arg-max of each window
// (first maximum in row-major window order wins, as in ATen's max_pool2d; a winning zero-halo element
// swallows the gradient), multiplies by LeakyReLU'(act) and writes dZ = d(loss)/d(pre-activation) for
// the full-resolution tensor (zeros elsewhere).  One thread = one window x 8 channels; the arg-max and the sign only need the
// activation up to its (positive) scale.  Grid-stride with blockDim.x a multiple of C/8: a thread keeps its channel group, so
// the column sums of dZ (the bias gradient of the conv in front of the pool) accumulate in registers; partial [gridDim.x][C],
// fixed-order second stage.
__global__ void __launch_bounds__(kFusedColsumBlock)
pool_bwd_kernel(const __half* __restrict__ act, const __half* __restrict__ act_lo, Geom ga_, int a_cpitch, int a_coff,
                const float* __restrict__ g1, int g1_cpitch, int g1_coff,
                const float* __restrict__ g2, int g2_cpitch, int g2_coff, Geom gp,
                __half* __restrict__ dv, __half* __restrict__ dlo, int d_cpitch, int d_coff, ScaleRef dsc,
                int C, int blind, float* __restrict__ colsum_partial) {
  pdl_wait();
  extern __shared__ float sm_pool[];      // [blockDim.x][8]
  const int c8n = C / 8;
  const long long n = (long long)gp.B * gp.H * gp.W * c8n;
  const float s = dsc.k ? exp2_int(__ldg(dsc.k)) : 1.0f;
  float csum[8], mx = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) csum[i] = 0.f;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % c8n) * 8; long long t = idx / c8n;
    const int xo = (int)(t % gp.W); t /= gp.W;
    const int yo = (int)(t % gp.H); const int b = (int)(t / gp.H);
    const long long pflat = (long long)b * gp.S + (yo + gp.row0) * gp.P + xo;
    float g[8];
    {
      const float4 u = __ldg(reinterpret_cast<const float4*>(g1 + pflat * g1_cpitch + g1_coff + c));
      const float4 v = __ldg(reinterpret_cast<const float4*>(g1 + pflat * g1_cpitch + g1_coff + c + 4));
      g[0] = u.x; g[1] = u.y; g[2] = u.z; g[3] = u.w; g[4] = v.x; g[5] = v.y; g[6] = v.z; g[7] = v.w;
      if (g2) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(g2 + pflat * g2_cpitch + g2_coff + c));
        const float4 q = __ldg(reinterpret_cast<const float4*>(g2 + pflat * g2_cpitch + g2_coff + c + 4));
        g[0] += p.x; g[1] += p.y; g[2] += p.z; g[3] += p.w; g[4] += q.x; g[5] += q.y; g[6] += q.z; g[7] += q.w;
      }
    }
    const int y0 = 2 * yo - (blind ? 1 : 0);
    const long long f0 = (long long)b * ga_.S + (y0 + ga_.row0) * ga_.P + 2 * xo;
    const long long fl[4] = {f0, f0 + 1, f0 + ga_.P, f0 + ga_.P + 1};
    float a[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k) ld8(act, act_lo, fl[k] * a_cpitch + a_coff + c, a[k]);
    float o[4][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float best = -INFINITY; int arg = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) { if (a[k][i] > best || a[k][i] != a[k][i]) { best = a[k][i]; arg = k; } }
      const float r = best > 0.f ? g[i] : SSDN_LRELU_SLOPE * g[i];
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k][i] = (k == arg) ? r : 0.f;
    }
    const bool halo_row = blind && yo == 0;   // window rows (-1, 0): elements 0,1 are padding
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (halo_row && k < 2) continue;        // never write the halo
      float sc8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { csum[i] += o[k][i]; mx = fmaxf(mx, fabsf(o[k][i])); sc8[i] = o[k][i] * s; }
      st8(dv, dlo, fl[k] * d_cpitch + d_coff + c, sc8);
    }
  }
  amax_commit(dsc.amax, mx);
  if (colsum_partial) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sm_pool[threadIdx.x * 8 + i] = csum[i];
    __syncthreads();
    if (threadIdx.x < C) {
      const int grp = threadIdx.x >> 3, i = threadIdx.x & 7;
      float t = 0.f;
      for (int k = grp; k < blockDim.x; k += c8n) t += sm_pool[k * 8 + i];
      colsum_partial[(long long)blockIdx.x * C + threadIdx.x] = t;
    }
  }
}

// Backward of [LeakyReLU -> nearest 2x upsample]: dZ[b,y,x,c] = LeakyReLU'(act) * sum of the 2x2 block of g.
// The forward activation's sign is read from the hi plane of its upsampled copy (geometry gg, pixel (2y, 2x)).
// One thread = one pixel x 8 channels, grid-stride with blockDim.x a multiple of C/8 (see pool_bwd_kernel).
__global__ void __launch_bounds__(kFusedColsumBlock)
up_bwd_kernel(const float* __restrict__ g, Geom gg, int g_cpitch, int g_coff,
              const __half* __restrict__ act_hi, int a_cpitch, int a_coff, Geom gl,
              __half* __restrict__ dv, __half* __restrict__ dlo, int d_cpitch, int d_coff, ScaleRef dsc, int C,
              float* __restrict__ colsum_partial) {
  pdl_wait();
  extern __shared__ float sm_up[];        // [blockDim.x][8]
  const int c8n = C / 8;
  const long long n = (long long)gl.B * gl.H * gl.W * c8n;
  const float sc = dsc.k ? exp2_int(__ldg(dsc.k)) : 1.0f;
  float csum[8], mx = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) csum[i] = 0.f;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % c8n) * 8; long long t = idx / c8n;
    const int x = (int)(t % gl.W); t /= gl.W;
    const int y = (int)(t % gl.H); const int b = (int)(t / gl.H);
    const long long p00 = (long long)b * gg.S + (2 * y + gg.row0) * gg.P + 2 * x;
    const long long s0 = p00 * g_cpitch + g_coff + c;
    float s[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(g + s0 + 4 * h));
      const float4 bq = __ldg(reinterpret_cast<const float4*>(g + s0 + g_cpitch + 4 * h));
      const float4 cq = __ldg(reinterpret_cast<const float4*>(g + s0 + (long long)gg.P * g_cpitch + 4 * h));
      const float4 dq = __ldg(reinterpret_cast<const float4*>(g + s0 + (long long)(gg.P + 1) * g_cpitch + 4 * h));
      s[4 * h + 0] = (a.x + bq.x) + (cq.x + dq.x); s[4 * h + 1] = (a.y + bq.y) + (cq.y + dq.y);
      s[4 * h + 2] = (a.z + bq.z) + (cq.z + dq.z); s[4 * h + 3] = (a.w + bq.w) + (cq.w + dq.w);
    }
    const uint4 av = __ldg(reinterpret_cast<const uint4*>(act_hi + p00 * a_cpitch + a_coff + c));
    const uint32_t* aw = reinterpret_cast<const uint32_t*>(&av);
    float o[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 a2 = f16x2_to_float2(aw[i]);
      s[2 * i] = a2.x > 0.f ? s[2 * i] : SSDN_LRELU_SLOPE * s[2 * i];
      s[2 * i + 1] = a2.y > 0.f ? s[2 * i + 1] : SSDN_LRELU_SLOPE * s[2 * i + 1];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { csum[i] += s[i]; mx = fmaxf(mx, fabsf(s[i])); o[i] = s[i] * sc; }
    const long long lf = (long long)b * gl.S + (y + gl.row0) * gl.P + x;
    st8(dv, dlo, lf * d_cpitch + d_coff + c, o);
  }
  amax_commit(dsc.amax, mx);
  if (colsum_partial) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sm_up[threadIdx.x * 8 + i] = csum[i];
    __syncthreads();
    if (threadIdx.x < C) {
      const int grp = threadIdx.x >> 3, i = threadIdx.x & 7;
      float t = 0.f;
      for (int k = grp; k < blockDim.x; k += c8n) t += sm_up[k * 8 + i];
      colsum_partial[(long long)blockIdx.x * C + threadIdx.x] = t;
    }
  }
}

// column sums of a dense NCHW tensor: partial[n][c] = sum over h, w (the bias gradient of the last conv is the sum of
// d(loss)/d(output) itself); grid = (C, N)
__global__ void nchw_colsum_kernel(const float* __restrict__ x, int C, int HW, float* __restrict__ partial) {
  pdl_wait();
  __shared__ float sm[32];
  const float* row = x + ((long long)blockIdx.y * C + blockIdx.x) * HW;
  float acc = 0.f;
  if ((HW & 3) == 0 && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {      // 16-byte loads, fixed order
    const float4* r4 = reinterpret_cast<const float4*>(row);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int i = threadIdx.x; i < (HW >> 2); i += blockDim.x) { const float4 v = __ldg(r4 + i); a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w; }
    acc = (a0 + a1) + (a2 + a3);
  } else {
    for (int i = threadIdx.x; i < HW; i += blockDim.x) acc += __ldg(row + i);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
    partial[(long long)blockIdx.y * C + blockIdx.x] = t;
  }
}

// ---------------------------------------------------------------------------- operand scales
// State of one network's scale slots (device memory, carved from the workspace).
struct ScaleState {
  int* k; int* k_next; unsigned* amax;   // [n_slots]
  int* status;                           // [0] forward pass stale, [1] backward pass stale, [2] number of stale passes so far
  unsigned* counter;                     // ticket counter of leaf_scale_kernel
};
// Start of a pass over slots [first, first + count): adopt the scales derived from the previous pass, clear the maxima.
__global__ void scale_begin_kernel(ScaleState st, int first, int count) {
  pdl_wait();
  const int i = first + threadIdx.x + blockIdx.x * blockDim.x;
  if (i < first + count) { st.k[i] = st.k_next[i]; st.amax[i] = 0u; }
}
// End of a pass: a slot whose maximum left the accurate band [2^kScaleLoLimit, 2^kScaleHiLimit) of its fp16 planes makes the
// pass stale (the host re-runs it / the optimiser skips it); every slot that saw data gets the scale that puts this pass's
// maximum at 2^kScaleTarget.  which = 0 (forward) or 1 (backward); stale_out (optional, backward) receives
// (forward stale | backward stale) as a float appended to the gradient buffer, so that it takes part in the all-reduce.
// apply != 0 (backward: nobody reads these slots any more): adopt the new scales and clear the maxima right away, so that
// the next pass over these slots needs no scale_begin_kernel.
// clear_count > 0: also clear the maxima of slots [clear_first, clear_first + clear_count) (the weight slots, whose readers are done).
__global__ void scale_finish_kernel(ScaleState st, int first, int count, int which, float* __restrict__ stale_out, int apply,
                                    int clear_first, int clear_count) {
  pdl_wait();
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  for (int i = threadIdx.x; i < clear_count; i += blockDim.x) st.amax[clear_first + i] = 0u;
  __syncthreads();
  for (int i = first + threadIdx.x; i < first + count; i += blockDim.x) {
    const unsigned a = st.amax[i];
    if (a) {
      const int e = floor_log2_bits(a) + st.k[i];
      if (e >= kScaleHiLimit || e < kScaleLoLimit || a >= 0x7f800000u) atomicOr(&bad, 1);
      st.k_next[i] = scale_for_amax(a < 0x7f800000u ? a : 0x7f7fffffu);
    }
    if (apply) { st.k[i] = st.k_next[i]; st.amax[i] = 0u; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    st.status[which] = bad;
    if (bad) st.status[2] += 1;
    if (stale_out) *stale_out = (float)((which == 1 ? st.status[0] : 0) | bad);
  }
}
// Exact scale of a LEAF tensor (dense fp32, e.g. the loss gradient) before it is packed: max|x| over n elements, then the
// last block to finish sets k[slot] (and, when init_count > 0, seeds slots [init_first, init_first + init_count) with the
// same exponent: the first backward pass of a plan has no previous maxima to go by).
__global__ void leaf_scale_kernel(const float* __restrict__ x, long long n, ScaleState st, int slot, int init_first, int init_count) {
  pdl_wait();
  __shared__ float sm[32];
  float m = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;      // four independent loads in flight per thread (see weight_scale_kernel)
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += 4 * stride) {
    const float a = fabsf(__ldg(x + i));
    const float b = i + stride < n ? fabsf(__ldg(x + i + stride)) : 0.f;
    const float c = i + 2 * stride < n ? fabsf(__ldg(x + i + 2 * stride)) : 0.f;
    const float d = i + 3 * stride < n ? fabsf(__ldg(x + i + 3 * stride)) : 0.f;
    m = fmaxf(fmaxf(m, fmaxf(a, b)), fmaxf(c, d));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, sm[w]);
    if (m > 0.f) atomicMax(&st.amax[slot], __float_as_uint(m));
    __threadfence();
    const unsigned ticket = atomicAdd(st.counter, 1u);
    if (ticket == gridDim.x - 1) {
      const unsigned a = atomicMax(&st.amax[slot], 0u);
      const int k = scale_for_amax(a < 0x7f800000u ? a : 0x7f7fffffu);
      st.k[slot] = k; st.k_next[slot] = k;
      for (int j = 0; j < init_count; ++j) if (init_first + j != slot) { st.k[init_first + j] = k; st.k_next[init_first + j] = k; }
      *st.counter = 0u;
    }
  }
}
// Exact scales of all weight tensors of a network: grid (kWeightScaleBlocks, jobs) folds max|w| of every tensor into its
// slot's amax (zero on entry: scale_finish_kernel clears the weight slots at the end of every forward pass); the weight-prep
// kernel then turns the maximum into the exponent (weight_scale_from_amax) and records it for the conv kernels.
// begin_count > 0: block (0, 0) also begins the pass over slots [begin_first, begin_first + begin_count) (scale_begin_kernel):
// this is the first kernel of a forward pass.
struct WeightScaleJob { const float* w; int n; int slot; };
constexpr int kMaxScaleJobs = 24;
constexpr int kWeightScaleBlocks = 32;
struct WeightScaleJobs { WeightScaleJob j[kMaxScaleJobs]; };
__global__ void weight_scale_kernel(const __grid_constant__ WeightScaleJobs jobs, ScaleState st, int begin_first, int begin_count) {
  pdl_wait();
  __shared__ float sm[32];
  if (blockIdx.x == 0 && blockIdx.y == 0 && (int)threadIdx.x < begin_count) { const int i = begin_first + threadIdx.x; st.k[i] = st.k_next[i]; st.amax[i] = 0u; }
  const WeightScaleJob& q = jobs.j[blockIdx.y];
  float m = 0.f;
  // four independent loads in flight per thread: this kernel is the first of every step and nothing can start before it
  // (a dependent-load loop of 36 iterations made it 18 us for 5 MB)
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < q.n; i += 4 * stride) {
    const float a = fabsf(__ldg(q.w + i));
    const float b = i + stride < q.n ? fabsf(__ldg(q.w + i + stride)) : 0.f;
    const float c = i + 2 * stride < q.n ? fabsf(__ldg(q.w + i + 2 * stride)) : 0.f;
    const float d = i + 3 * stride < q.n ? fabsf(__ldg(q.w + i + 3 * stride)) : 0.f;
    m = fmaxf(fmaxf(m, fmaxf(a, b)), fmaxf(c, d));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, sm[w]);
    if (m > 0.f) atomicMax(&st.amax[q.slot], __float_as_uint(m));
  }
}
__device__ __forceinline__ int weight_scale_from_amax(const ScaleState& st, int slot) {
  const unsigned a = st.amax[slot];
  return scale_for_amax(a < 0x7f800000u ? a : 0x7f7fffffu);
}

// ---------------------------------------------------------------------------- weights
// Builds the K-major weight slab [n_tile][chunk][tap][plane][N][CW] (plane 0 = hi, 1 = lo of the scaled fp16 split) from
// PyTorch-layout weights W[cout][cin][taps].  transpose == 0 (forward):  slab[n][k] = W[n][k][tap]
//                                           transpose == 1 (data-grad): slab[n][k] = W[k][n][tap]   (n = cin, k = cout)
// One thread = 8 consecutive k of one slab row: one 16-byte store per plane, index arithmetic in 32 bits.
__device__ __forceinline__ void weight_prep_vec8(const float* __restrict__ w, __half* __restrict__ slab, int cin, int ntaps, int n_valid, int k_valid,
                                                 int n_chunks, int N, int transpose, int CW, float s, int idx) {
  const int v8 = CW >> 3;
  const int kk = (idx % v8) * 8; int t = idx / v8;
  const int n = t % N; t /= N;                       // t = slab index ((nt * n_chunks + ch) * ntaps + tap)
  const int tap = t % ntaps; const int tc = t / ntaps;
  const int ch = tc % n_chunks; const int nt = tc / n_chunks;
  const int ng = nt * N + n, k0 = ch * CW + kk;
  float f[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int k = k0 + q;
    float val = 0.f;
    if (ng < n_valid && k < k_valid) val = __ldg(w + (transpose ? (k * cin + ng) * ntaps + tap : (ng * cin + k) * ntaps + tap));
    f[q] = val * s;
  }
  uint4 h, l;
  f16_split8(f, h, l);
  const long long o = ((long long)(t * 2) * N + n) * CW + kk;
  *reinterpret_cast<uint4*>(slab + o) = h; *reinterpret_cast<uint4*>(slab + o + (long long)N * CW) = l;
}
// slot >= 0: the weights' scale exponent comes from st.amax[slot] (weight_scale_kernel ran before) and is recorded in st.k[slot]
__global__ void weight_prep_kernel(const float* __restrict__ w, __half* __restrict__ slab, int cout, int cin, int ntaps,
                                   int n_valid, int k_valid, int n_tiles, int n_chunks, int N, int transpose, int CW, ScaleState st, int slot) {
  pdl_wait();
  const int total = n_tiles * n_chunks * ntaps * N * (CW >> 3);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = slot >= 0 ? weight_scale_from_amax(st, slot) : 0;
  if (slot >= 0 && idx == 0) { st.k[slot] = k; st.k_next[slot] = k; }
  if (idx >= total) return;
  weight_prep_vec8(w, slab, cin, ntaps, n_valid, k_valid, n_chunks, N, transpose, CW, exp2_int(k), idx);
}

// All weight slabs of a network in ONE launch: blockIdx.y selects the job, blockIdx.x strides over its 8-element vectors.
struct WeightPrepJob { const float* w; __half* slab; int cout, cin, ntaps, n_valid, k_valid, n_tiles, n_chunks, N, transpose, CW; int slot; };
constexpr int kMaxPrepJobs = 48;
struct WeightPrepJobs { WeightPrepJob j[kMaxPrepJobs]; };
__global__ void weight_prep_batched_kernel(const __grid_constant__ WeightPrepJobs jobs, ScaleState st) {
  pdl_wait();
  const WeightPrepJob& q = jobs.j[blockIdx.y];
  const int total = q.n_tiles * q.n_chunks * q.ntaps * q.N * (q.CW >> 3);
  const int k = weight_scale_from_amax(st, q.slot);
  if (blockIdx.x == 0 && threadIdx.x == 0 && !q.transpose) { st.k[q.slot] = k; st.k_next[q.slot] = k; }    // the forward slab's job records it
  const float s = exp2_int(k);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x)
    weight_prep_vec8(q.w, q.slab, q.cin, q.ntaps, q.n_valid, q.k_valid, q.n_chunks, q.N, q.transpose, q.CW, s, idx);
}

// ---------------------------------------------------------------------------- bias gradient
// db[c] = sum over flat pixels of dZ[flat][c] (hi + lo planes, unscaled).  Two deterministic stages: kColsumBlocks strips of
// rows, then a fixed-order reduction of the strip partials.  (Single-operator entry point only: inside the network every
// producer of a dZ leaves its column sums behind.)
constexpr int kColsumBlocks = 296;
__global__ void colsum_stage1_kernel(const __half* __restrict__ dz, const __half* __restrict__ dz_lo, const int* __restrict__ k, long long rows, int cpitch,
                                     int coff, int C, float* __restrict__ partial) {
  pdl_wait();
  extern __shared__ float sm[];   // [warps][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * per, r1 = r0 + per < rows ? r0 + per : rows;
  const float inv = exp2_int(k ? -__ldg(k) : 0);
  for (int c = lane; c < C; c += 32) {
    float acc = 0.f;
    for (long long r = r0 + warp; r < r1; r += nwarps) acc += f16_join1(dz[r * cpitch + coff + c], dz_lo[r * cpitch + coff + c]);
    sm[warp * C + c] = acc * inv;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int w = 0; w < nwarps; ++w) acc += sm[w * C + c];
    partial[(long long)blockIdx.x * C + c] = acc;
  }
}

// grid = ceil(C / 32), block = (32, 32): thread (lane, w) sums partials w, w+32, ... of channel blockIdx.x*32 + lane
__global__ void colsum_stage2_kernel(const float* __restrict__ partial, int nblk, int C, float* __restrict__ out, int accumulate) {
  pdl_wait();
  __shared__ float sm[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float a0 = 0.f, a1 = 0.f;
  if (c < C) {
    int b = threadIdx.y;
    for (; b + 32 < nblk; b += 64) { a0 += partial[(long long)b * C + c]; a1 += partial[(long long)(b + 32) * C + c]; }
    if (b < nblk) a0 += partial[(long long)b * C + c];
  }
  sm[threadIdx.y][threadIdx.x] = a0 + a1;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 32; ++w) t += sm[w][threadIdx.x];
    out[c] = accumulate ? out[c] + t : t;
  }
}
// All bias gradients of a network in ONE launch at the end of the backward pass (every layer has its own partial buffer):
// blockIdx.y selects the job, blockIdx.x a group of 32 channels.
struct BiasJob { const float* partial; float* out; int nblk, C; };
constexpr int kMaxBiasJobs = 24;
struct BiasJobs { BiasJob j[kMaxBiasJobs]; };
__global__ void colsum_stage2_batched_kernel(const __grid_constant__ BiasJobs jobs) {
  pdl_wait();
  __shared__ float sm[32][33];
  const BiasJob& q = jobs.j[blockIdx.y];
  const int c = blockIdx.x * 32 + threadIdx.x;
  if (blockIdx.x * 32 >= q.C) return;
  float a0 = 0.f, a1 = 0.f;
  if (c < q.C) {
    int b = threadIdx.y;
    for (; b + 32 < q.nblk; b += 64) { a0 += q.partial[(long long)b * q.C + c]; a1 += q.partial[(long long)(b + 32) * q.C + c]; }
    if (b < q.nblk) a0 += q.partial[(long long)b * q.C + c];
  }
  sm[threadIdx.y][threadIdx.x] = a0 + a1;
  __syncthreads();
  if (threadIdx.y == 0 && c < q.C) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 32; ++w) t += sm[w][threadIdx.x];
    q.out[c] = t;
  }
}
static inline void colsum_stage2_launch(const float* partial, int nblk, int C, float* out, cudaStream_t st) {
  colsum_stage2_kernel<<<(C + 31) / 32, dim3(32, 32), 0, st>>>(partial, nblk, C, out, 0);
}

static inline void colsum_launch(const __half* hi, const __half* lo, const int* k, long long rows, int cpitch, int coff, int C, float* partial, float* out,
                                 cudaStream_t st) {
  colsum_stage1_kernel<<<kColsumBlocks, 256, 8 * C * sizeof(float), st>>>(hi, lo, k, rows, cpitch, coff, C, partial);
  colsum_stage2_launch(partial, kColsumBlocks, C, out, st);
}

}  // namespace pw
