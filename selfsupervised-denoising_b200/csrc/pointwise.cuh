// HBM-bound kernels of the U-Net: layout packing (with the 4-rotation stack), shifted max-pool
// forward/backward, upsample backward, weight slab preparation, bias gradients.
// All of them stream each byte once; accesses are float4 along the channel axis where the channel
// count allows it.  Index work is exact integer arithmetic (bit-exact versus the oracle).
#pragma once
#include "common.cuh"

namespace pw {

constexpr int kBlock = 256;
static inline int grid_for(long long n, int block = kBlock) { return (int)((n + block - 1) / block); }

// ---------------------------------------------------------------------------- NCHW -> padded flat
// rot4 != 0: the output holds 4*B images, image (r*B + b) = rotate(x[b], 90*r)   (utils/data.py:42-67)
__global__ void pack_nchw_kernel(const float* __restrict__ x, float* __restrict__ v, float* __restrict__ lo,
                                 int B, int C, int H, int W, Geom g, int cpitch, int coff, int rot4) {
  const long long n = (long long)(rot4 ? 4 : 1) * B * C * H * W;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int j = (int)(idx % W); long long t = idx / W;
  const int i = (int)(t % H); t /= H;
  const int c = (int)(t % C); const int bo = (int)(t / C);
  const int r = rot4 ? bo / B : 0, b = rot4 ? bo % B : bo;
  int si = i, sj = j;
  if (r == 1) { si = j; sj = W - 1 - i; } else if (r == 2) { si = H - 1 - i; sj = W - 1 - j; }
  else if (r == 3) { si = H - 1 - j; sj = i; }
  const float val = __ldg(x + (((long long)b * C + c) * H + si) * W + sj);
  const long long o = ((long long)bo * g.S + (i + g.row0) * g.P + j) * cpitch + coff + c;
  if (lo) { float h, l; tf32_split(val, h, l); v[o] = h; lo[o] = l; }
  else v[o] = val;
}

// Same mapping, one thread per output PIXEL writing all C (<= 16) channels: the stores of a thread are contiguous and
// the index arithmetic is done once per pixel (the network input and the loss gradient have 3..12 channels).
__global__ void pack_nchw_pixel_kernel(const float* __restrict__ x, float* __restrict__ v, float* __restrict__ lo,
                                       int B, int C, int H, int W, Geom g, int cpitch, int coff, int rot4) {
  const long long n = (long long)(rot4 ? 4 : 1) * B * H * W;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int j = (int)(idx % W); long long t = idx / W;
  const int i = (int)(t % H); const int bo = (int)(t / H);
  const int r = rot4 ? bo / B : 0, b = rot4 ? bo % B : bo;
  int si = i, sj = j;
  if (r == 1) { si = j; sj = W - 1 - i; } else if (r == 2) { si = H - 1 - i; sj = W - 1 - j; }
  else if (r == 3) { si = H - 1 - j; sj = i; }
  const float* src = x + ((long long)b * C * H + si) * W + sj;
  const long long o = ((long long)bo * g.S + (i + g.row0) * g.P + j) * cpitch + coff;
  for (int c = 0; c < C; ++c) {
    const float val = __ldg(src + (long long)c * H * W);
    if (lo) { float h, l; tf32_split(val, h, l); v[o + c] = h; lo[o + c] = l; }
    else v[o + c] = val;
  }
}

// padded flat -> dense NCHW (tests / gradients w.r.t. the input)
__global__ void unpack_nchw_kernel(const float* __restrict__ v, const float* __restrict__ lo, float* __restrict__ y, int B, int C,
                                   int H, int W, Geom g, int cpitch, int coff) {
  const long long n = (long long)B * C * H * W;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int j = (int)(idx % W); long long t = idx / W;
  const int i = (int)(t % H); t /= H;
  const int c = (int)(t % C); const int b = (int)(t / C);
  const long long o = ((long long)b * g.S + (i + g.row0) * g.P + j) * cpitch + coff + c;
  y[idx] = lo ? v[o] + lo[o] : v[o];
}

// LeakyReLU sign masks of a two-plane tensor (test hook: after an activation buffer was overwritten from outside).
__global__ void mask_from_planes_kernel(const float* __restrict__ v, long long pixels, int cpitch, uint32_t* __restrict__ mask, int words) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= pixels * words) return;
  const long long px = idx / words; const int w = (int)(idx - px * words);
  uint32_t m = 0;
  for (int i = 0; i < 32 && 32 * w + i < cpitch; ++i) m |= (v[px * cpitch + 32 * w + i] > 0.f ? 1u : 0u) << i;
  mask[idx] = m;
}

// ---------------------------------------------------------------------------- max-pool 2x2
// blind != 0: Shift2d((1,0)) then MaxPool2d(2)  (models/noise_network.py:64-67): window rows (2i-1, 2i),
// row -1 is the zero halo row of the padded layout.  One thread = one output pixel x 4 channels.
__device__ __forceinline__ float4 ld2(const float* __restrict__ hi, const float* __restrict__ lo, long long i) {
  const float4 a = *reinterpret_cast<const float4*>(hi + i), b = *reinterpret_cast<const float4*>(lo + i);
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ void st2(float* __restrict__ hi, float* __restrict__ lo, long long i, float4 v) {
  float4 h, l;
  tf32_split(v.x, h.x, l.x); tf32_split(v.y, h.y, l.y); tf32_split(v.z, h.z, l.z); tf32_split(v.w, h.w, l.w);
  *reinterpret_cast<float4*>(hi + i) = h; *reinterpret_cast<float4*>(lo + i) = l;
}

__global__ void pool_fwd_kernel(const float* __restrict__ src, const float* __restrict__ src_lo, Geom gs, int s_cpitch, int s_coff,
                                float* __restrict__ dv, float* __restrict__ dlo, Geom gd, int d_cpitch, int d_coff,
                                int C, int blind) {
  const int c4n = C / 4;
  const long long n = (long long)gd.B * gd.H * gd.W * c4n;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int c = (int)(idx % c4n) * 4; long long t = idx / c4n;
  const int xo = (int)(t % gd.W); t /= gd.W;
  const int yo = (int)(t % gd.H); const int b = (int)(t / gd.H);
  const int y0 = 2 * yo - (blind ? 1 : 0);
  const long long s0 = ((long long)b * gs.S + (y0 + gs.row0) * gs.P + 2 * xo) * s_cpitch + s_coff + c;
  const float4 a = ld2(src, src_lo, s0);
  const float4 bq = ld2(src, src_lo, s0 + s_cpitch);
  const float4 cq = ld2(src, src_lo, s0 + (long long)gs.P * s_cpitch);
  const float4 dq = ld2(src, src_lo, s0 + (long long)(gs.P + 1) * s_cpitch);
  float4 m;
  m.x = fmaxf(fmaxf(a.x, bq.x), fmaxf(cq.x, dq.x)); m.y = fmaxf(fmaxf(a.y, bq.y), fmaxf(cq.y, dq.y));
  m.z = fmaxf(fmaxf(a.z, bq.z), fmaxf(cq.z, dq.z)); m.w = fmaxf(fmaxf(a.w, bq.w), fmaxf(cq.w, dq.w));
  st2(dv, dlo, ((long long)b * gd.S + (yo + gd.row0) * gd.P + xo) * d_cpitch + d_coff + c, m);
}

// Backward of [LeakyReLU -> (shift) -> max-pool]: routes g = ga (+ gb) to the arg-max of each window
// (first maximum in row-major window order wins, as in ATen's max_pool2d; a winning zero-halo element
// swallows the gradient), multiplies by LeakyReLU'(act) and writes dZ = d(loss)/d(pre-activation) for
// the full-resolution tensor (zeros elsewhere).  One thread = one window x one channel.
__global__ void pool_bwd_kernel(const float* __restrict__ act, const float* __restrict__ act_lo, Geom ga_, int a_cpitch, int a_coff,
                                const float* __restrict__ g1, int g1_cpitch, int g1_coff,
                                const float* __restrict__ g2, int g2_cpitch, int g2_coff, Geom gp,
                                float* __restrict__ dv, float* __restrict__ dlo, int d_cpitch, int d_coff,
                                int C, int blind, float* __restrict__ colsum_partial) {
  // grid-stride with blockDim.x a multiple of C: a thread keeps its channel, so the column sums of dZ (the bias gradient of
  // the conv in front of the pool) accumulate in a register; partial [gridDim.x][C], fixed-order second stage.
  extern __shared__ float sm_pool[];
  const long long n = (long long)gp.B * gp.H * gp.W * C;
  float csum = 0.f;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C); long long t = idx / C;
    const int xo = (int)(t % gp.W); t /= gp.W;
    const int yo = (int)(t % gp.H); const int b = (int)(t / gp.H);
    const long long pflat = (long long)b * gp.S + (yo + gp.row0) * gp.P + xo;
    float g = __ldg(g1 + pflat * g1_cpitch + g1_coff + c);
    if (g2) g += __ldg(g2 + pflat * g2_cpitch + g2_coff + c);
    const int y0 = 2 * yo - (blind ? 1 : 0);
    const long long f0 = (long long)b * ga_.S + (y0 + ga_.row0) * ga_.P + 2 * xo;
    const long long fl[4] = {f0, f0 + 1, f0 + ga_.P, f0 + ga_.P + 1};
    float best = -INFINITY; int arg = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a = __ldg(act + fl[k] * a_cpitch + a_coff + c) + __ldg(act_lo + fl[k] * a_cpitch + a_coff + c);
      if (a > best || a != a) { best = a; arg = k; }
    }
    const bool halo_row = blind && yo == 0;   // window rows (-1, 0): elements 0,1 are padding
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (halo_row && k < 2) continue;        // never write the halo
      float o = 0.f;
      if (k == arg) o = best > 0.f ? g : SSDN_LRELU_SLOPE * g;
      csum += o;
      const long long di = fl[k] * d_cpitch + d_coff + c;
      tf32_split(o, dv[di], dlo[di]);
    }
  }
  if (colsum_partial) {
    sm_pool[threadIdx.x] = csum;
    __syncthreads();
    if (threadIdx.x < C) {
      float t = 0.f;
      for (int k = threadIdx.x; k < blockDim.x; k += C) t += sm_pool[k];
      colsum_partial[(long long)blockIdx.x * C + threadIdx.x] = t;
    }
  }
}

// Backward of [LeakyReLU -> nearest 2x upsample]: dZ[b,y,x,c] = LeakyReLU'(act) * sum of the 2x2 block of g.
// The forward activation is read from its upsampled copy (geometry gg, pixel (2y, 2x)).
__global__ void up_bwd_kernel(const float* __restrict__ g, Geom gg, int g_cpitch, int g_coff,
                              const float* __restrict__ act, int a_cpitch, int a_coff, Geom gl,
                              float* __restrict__ dv, float* __restrict__ dlo, int d_cpitch, int d_coff, int C,
                              float* __restrict__ colsum_partial) {
  // grid-stride with blockDim.x a multiple of C/4 (see pool_bwd_kernel): fused column sums of the produced dZ
  extern __shared__ float4 sm_up[];
  const int c4n = C / 4;
  const long long n = (long long)gl.B * gl.H * gl.W * c4n;
  float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % c4n) * 4; long long t = idx / c4n;
    const int x = (int)(t % gl.W); t /= gl.W;
    const int y = (int)(t % gl.H); const int b = (int)(t / gl.H);
    const long long s0 = ((long long)b * gg.S + (2 * y + gg.row0) * gg.P + 2 * x) * g_cpitch + g_coff + c;
    const float4 a = *reinterpret_cast<const float4*>(g + s0);
    const float4 bq = *reinterpret_cast<const float4*>(g + s0 + g_cpitch);
    const float4 cq = *reinterpret_cast<const float4*>(g + s0 + (long long)gg.P * g_cpitch);
    const float4 dq = *reinterpret_cast<const float4*>(g + s0 + (long long)(gg.P + 1) * g_cpitch);
    const long long lf = (long long)b * gl.S + (y + gl.row0) * gl.P + x;
    const float4 av = *reinterpret_cast<const float4*>(
        act + ((long long)b * gg.S + (2 * y + gg.row0) * gg.P + 2 * x) * a_cpitch + a_coff + c);
    float4 s;
    s.x = (a.x + bq.x) + (cq.x + dq.x); s.y = (a.y + bq.y) + (cq.y + dq.y);
    s.z = (a.z + bq.z) + (cq.z + dq.z); s.w = (a.w + bq.w) + (cq.w + dq.w);
    s.x = av.x > 0.f ? s.x : SSDN_LRELU_SLOPE * s.x; s.y = av.y > 0.f ? s.y : SSDN_LRELU_SLOPE * s.y;
    s.z = av.z > 0.f ? s.z : SSDN_LRELU_SLOPE * s.z; s.w = av.w > 0.f ? s.w : SSDN_LRELU_SLOPE * s.w;
    csum.x += s.x; csum.y += s.y; csum.z += s.z; csum.w += s.w;
    st2(dv, dlo, lf * d_cpitch + d_coff + c, s);
  }
  if (colsum_partial) {
    sm_up[threadIdx.x] = csum;
    __syncthreads();
    if (threadIdx.x < c4n) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = threadIdx.x; k < blockDim.x; k += c4n) { const float4 v = sm_up[k]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
      reinterpret_cast<float4*>(colsum_partial + (long long)blockIdx.x * C)[threadIdx.x] = t;
    }
  }
}
constexpr int kFusedColsumBlock = 240;    // multiple of 48 (pool channels), 24 and 12 (upsample channel quads)
constexpr int kFusedColsumGrid = 592;

// column sums of a dense NCHW tensor: partial[n][c] = sum over h, w (the bias gradient of the last conv is the sum of
// d(loss)/d(output) itself); grid = (C, N)
__global__ void nchw_colsum_kernel(const float* __restrict__ x, int C, int HW, float* __restrict__ partial) {
  __shared__ float sm[32];
  const float* row = x + ((long long)blockIdx.y * C + blockIdx.x) * HW;
  float acc = 0.f;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) acc += __ldg(row + i);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
    partial[(long long)blockIdx.y * C + blockIdx.x] = t;
  }
}

// ---------------------------------------------------------------------------- weights
// Builds the K-major weight slab [n_tile][chunk][tap][plane][N][16] (plane 0 = hi, 1 = lo of the tf32 split) from
// PyTorch-layout weights W[cout][cin][taps].  transpose == 0 (forward):  slab[n][k] = W[n][k][tap]
//                                           transpose == 1 (data-grad): slab[n][k] = W[k][n][tap]   (n = cin, k = cout)
__global__ void weight_prep_kernel(const float* __restrict__ w, float* __restrict__ slab, int cout, int cin, int ntaps,
                                   int n_valid, int k_valid, int n_tiles, int n_chunks, int N, int transpose, int CW) {
  const long long total = (long long)n_tiles * n_chunks * ntaps * N * CW;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int kk = (int)(idx % CW); long long t = idx / CW;
  const int n = (int)(t % N); t /= N;                       // t = slab index ((nt * n_chunks + ch) * ntaps + tap)
  const int tap = (int)(t % ntaps); const long long tc = t / ntaps;
  const int ch = (int)(tc % n_chunks); const int nt = (int)(tc / n_chunks);
  const int ng = nt * N + n, k = ch * CW + kk;
  float val = 0.f;
  if (ng < n_valid && k < k_valid) {
    const long long wi = transpose ? ((long long)k * cin + ng) * ntaps + tap : ((long long)ng * cin + k) * ntaps + tap;
    val = __ldg(w + wi);
  }
  const long long o = ((t * 2) * N + n) * CW + kk;
  tf32_split(val, slab[o], slab[o + (long long)N * CW]);
}

// All weight slabs of a network in ONE launch: blockIdx.y selects the job, blockIdx.x strides over its elements.
struct WeightPrepJob { const float* w; float* slab; int cout, cin, ntaps, n_valid, k_valid, n_tiles, n_chunks, N, transpose, CW; };
constexpr int kMaxPrepJobs = 48;
struct WeightPrepJobs { WeightPrepJob j[kMaxPrepJobs]; };
__global__ void weight_prep_batched_kernel(const __grid_constant__ WeightPrepJobs jobs) {
  const WeightPrepJob& q = jobs.j[blockIdx.y];
  const int CW = q.CW;
  const long long total = (long long)q.n_tiles * q.n_chunks * q.ntaps * q.N * CW;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(idx % CW); long long t = idx / CW;
    const int n = (int)(t % q.N); t /= q.N;
    const int tap = (int)(t % q.ntaps); const long long tc = t / q.ntaps;
    const int ch = (int)(tc % q.n_chunks); const int nt = (int)(tc / q.n_chunks);
    const int ng = nt * q.N + n, k = ch * CW + kk;
    float val = 0.f;
    if (ng < q.n_valid && k < q.k_valid) {
      const long long wi = q.transpose ? ((long long)k * q.cin + ng) * q.ntaps + tap : ((long long)ng * q.cin + k) * q.ntaps + tap;
      val = __ldg(q.w + wi);
    }
    const long long o = ((t * 2) * q.N + n) * CW + kk;
    tf32_split(val, q.slab[o], q.slab[o + (long long)q.N * CW]);
  }
}

// ---------------------------------------------------------------------------- bias gradient
// db[c] = sum over flat pixels of dZ[flat][c] (hi + lo planes).  Two deterministic stages: kColsumBlocks strips of rows,
// then a fixed-order reduction of the strip partials.
constexpr int kColsumBlocks = 296;
__global__ void colsum_stage1_kernel(const float* __restrict__ dz, const float* __restrict__ dz_lo, long long rows, int cpitch, int coff, int C,
                                     float* __restrict__ partial) {
  extern __shared__ float sm[];   // [warps][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * per, r1 = r0 + per < rows ? r0 + per : rows;
  for (int c = lane; c < C; c += 32) {
    float acc = 0.f;
    for (long long r = r0 + warp; r < r1; r += nwarps) acc += __ldg(dz + r * cpitch + coff + c) + __ldg(dz_lo + r * cpitch + coff + c);
    sm[warp * C + c] = acc;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int w = 0; w < nwarps; ++w) acc += sm[w * C + c];
    partial[(long long)blockIdx.x * C + c] = acc;
  }
}
// Fast path of stage 1 for dense tensors (cpitch == C, C % 4 == 0): the planes are flat float4 arrays in which a thread
// that advances by a multiple of C/4 float4s always stays in the same channel quad, so every access is coalesced and many
// independent loads are in flight.  blockDim.x must be a multiple of C/4.
__global__ void colsum_flat_kernel(const float4* __restrict__ hi, const float4* __restrict__ lo, long long total_f4, int C4,
                                   float* __restrict__ partial) {
  extern __shared__ float4 sm4[];
  const long long stride = (long long)gridDim.x * blockDim.x;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < total_f4; i += 4 * stride) {
    float4 a0 = __ldg(hi + i), a1 = __ldg(hi + i + stride), a2 = __ldg(hi + i + 2 * stride), a3 = __ldg(hi + i + 3 * stride);
    float4 b0 = __ldg(lo + i), b1 = __ldg(lo + i + stride), b2 = __ldg(lo + i + 2 * stride), b3 = __ldg(lo + i + 3 * stride);
    acc.x += ((a0.x + b0.x) + (a1.x + b1.x)) + ((a2.x + b2.x) + (a3.x + b3.x));
    acc.y += ((a0.y + b0.y) + (a1.y + b1.y)) + ((a2.y + b2.y) + (a3.y + b3.y));
    acc.z += ((a0.z + b0.z) + (a1.z + b1.z)) + ((a2.z + b2.z) + (a3.z + b3.z));
    acc.w += ((a0.w + b0.w) + (a1.w + b1.w)) + ((a2.w + b2.w) + (a3.w + b3.w));
  }
  for (; i < total_f4; i += stride) {
    const float4 a = __ldg(hi + i), b = __ldg(lo + i);
    acc.x += a.x + b.x; acc.y += a.y + b.y; acc.z += a.z + b.z; acc.w += a.w + b.w;
  }
  sm4[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < C4) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = threadIdx.x; k < blockDim.x; k += C4) { const float4 v = sm4[k]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
    reinterpret_cast<float4*>(partial + (long long)blockIdx.x * C4 * 4)[threadIdx.x] = t;
  }
}
// Launch helper: picks the flat fast path when the layout allows it.
static inline void colsum_launch(const float* hi, const float* lo, long long rows, int cpitch, int coff, int C, float* partial, float* out,
                                 cudaStream_t st);

// grid = ceil(C / 32), block = (32, 32): thread (lane, w) sums partials w, w+32, ... of channel blockIdx.x*32 + lane
__global__ void colsum_stage2_kernel(const float* __restrict__ partial, int nblk, int C, float* __restrict__ out, int accumulate) {
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

static inline void colsum_launch(const float* hi, const float* lo, long long rows, int cpitch, int coff, int C, float* partial, float* out,
                                 cudaStream_t st) {
  if (cpitch == C && coff == 0 && C % 4 == 0 && C / 4 <= 256) {
    const int C4 = C / 4, block = 256 / C4 * C4;
    colsum_flat_kernel<<<kColsumBlocks, block, block * sizeof(float4), st>>>(reinterpret_cast<const float4*>(hi), reinterpret_cast<const float4*>(lo),
                                                                             rows * C4, C4, partial);
  } else {
    colsum_stage1_kernel<<<kColsumBlocks, 256, 8 * C * sizeof(float), st>>>(hi, lo, rows, cpitch, coff, C, partial);
  }
  colsum_stage2_launch(partial, kColsumBlocks, C, out, st);
}

}  // namespace pw
