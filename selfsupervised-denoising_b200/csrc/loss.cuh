// HBM-bound per-pixel kernels of the SSDN pipeline (ssdn/ssdn/denoiser.py:222-397): Gaussian
// posterior mean, negative log-likelihood loss and their analytic gradients; MSE / masked-MSE losses
// (denoiser.py:140-180, utils/n2v_loss.py); sigma-estimator head (spatial mean + softplus); flat Adam.
// All tensors here are dense NCHW fp32 (the network boundary).  The 3x3 covariance algebra is done in
// closed form in fp64 registers (a few hundred flops per pixel; the kernels stay HBM-bound), per-sample
// reductions are two-stage and fixed-order (deterministic, no atomics).
#pragma once
#include "common.cuh"

namespace lossk {

constexpr int kBlock = 256;

struct Sym3 { double a00, a01, a02, a11, a12, a22; };
__device__ __forceinline__ double det3(const Sym3& m) {
  return m.a00 * (m.a11 * m.a22 - m.a12 * m.a12) - m.a01 * (m.a01 * m.a22 - m.a12 * m.a02) +
         m.a02 * (m.a01 * m.a12 - m.a11 * m.a02);
}
__device__ __forceinline__ Sym3 inv3(const Sym3& m, double det) {
  const double r = 1.0 / det;
  Sym3 o;
  o.a00 = (m.a11 * m.a22 - m.a12 * m.a12) * r; o.a01 = (m.a02 * m.a12 - m.a01 * m.a22) * r;
  o.a02 = (m.a01 * m.a12 - m.a02 * m.a11) * r; o.a11 = (m.a00 * m.a22 - m.a02 * m.a02) * r;
  o.a12 = (m.a01 * m.a02 - m.a00 * m.a12) * r; o.a22 = (m.a00 * m.a11 - m.a01 * m.a01) * r;
  return o;
}
__device__ __forceinline__ void mv3(const Sym3& m, double x, double y, double z, double& ox, double& oy, double& oz) {
  ox = m.a00 * x + m.a01 * y + m.a02 * z; oy = m.a01 * x + m.a11 * y + m.a12 * z; oz = m.a02 * x + m.a12 * y + m.a22 * z;
}

// sigma mapping: known -> max(raw, 1e-3) (denoiser.py:279-282); else softplus(raw - 4) + 1e-3 (:272-275)
__device__ __forceinline__ float sigma_map(float raw, int known) {
  if (known) return fmaxf(raw, 1e-3f);
  const float x = raw - 4.0f;
  return (x > 20.f ? x : log1pf(expf(x))) + 1e-3f;
}
__device__ __forceinline__ float sigma_map_grad(float raw) {
  const float x = raw - 4.0f;
  return x > 20.f ? 1.f : 1.f / (1.f + expf(-x));
}

__device__ __forceinline__ float block_sum(float v, float* sm) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sm[w] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0) for (int i = 0; i < (blockDim.x >> 5); ++i) t += sm[i];
  return t;   // valid in thread 0
}

// ---------------------------------------------------------------- SSDN posterior / NLL, forward
// grid = (blocks_per_sample, N).  partial[n][blk] = sum of per-pixel losses of that strip.
// Noise model.  Gaussian (poisson == 0): sigma_c is a per-sample constant, sigma_map(raw).  Poisson (denoiser.py:285-297,
// the signal-dependent approximation of Hasinoff 2012): sigma_c = sqrt(max(mu_c, 1e-3) * k_c) per pixel, with k_c = 1 / raw
// (raw = the known lambda) or k_c = softplus(raw - 4) + 1e-3 (learned); the per-pixel noise level is then an OUTPUT of the
// forward kernel (noise_std_px, [n][hw]) and the loss gradient reaches mu through sigma as well.
__device__ __forceinline__ float noise_coeff(float raw, int known, int poisson) {
  if (!poisson) return sigma_map(raw, known);
  return known ? 1.0f / raw : sigma_map(raw, 0);
}

template <int C>
__global__ void posterior_fwd_kernel(const float* __restrict__ net_out, const float* __restrict__ noisy,
                                     const float* __restrict__ sigma_raw, int cs, int known, int poisson, int diag, int HW,
                                     float* __restrict__ pme, float* __restrict__ model_std, float* __restrict__ noise_std_px,
                                     float* __restrict__ partial) {
  pdl_wait();
  __shared__ float sm[32];
  // diag (denoiser.py:236-243, cfg DIAGONAL_COVARIANCE): the network gives C diagonal factors instead of the triangular one
  const int CO = diag ? 2 * C : C + C * (C + 1) / 2;
  const int n = blockIdx.y;
  const float* no = net_out + (long long)n * CO * HW;
  const float* yy = noisy + (long long)n * C * HW;
  float sg[3], kc[3];
#pragma unroll
  for (int c = 0; c < C; ++c) { kc[c] = noise_coeff(sigma_raw[n * cs + (cs == 1 ? 0 : c)], known, poisson); sg[c] = kc[c]; }
  float acc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    if (C == 1) {
      const float mu = no[i], a = no[HW + i], y = yy[i];
      if (poisson) { sg[0] = sqrtf(fmaxf(mu, 1e-3f) * kc[0]); if (noise_std_px) noise_std_px[(long long)n * HW + i] = sg[0]; }
      const float sx = a * a, sn = sg[0] * sg[0], sy = sx + sn, d = y - mu;
      float l = d * d / sy + logf(sy);
      if (!known) l -= 0.1f * sg[0];
      acc += l;
      pme[(long long)n * HW + i] = (y * sx + mu * sn) / (sx + sn);
      model_std[(long long)n * HW + i] = sqrtf(sx);
    } else {
      double mu[3], y[3], a[6];
#pragma unroll
      for (int c = 0; c < 3; ++c) { mu[c] = no[c * HW + i]; y[c] = yy[c * HW + i]; }
      if (poisson) {
#pragma unroll
        for (int c = 0; c < 3; ++c) sg[c] = sqrtf(fmaxf((float)mu[c], 1e-3f) * kc[c]);
        if (noise_std_px) noise_std_px[(long long)n * HW + i] = (float)pow((double)sg[0] * sg[0] * sg[1] * sg[1] * sg[2] * sg[2], 1.0 / 6.0);
      }
      if (diag) {          // Sigma_x = diag(d0^2, d1^2, d2^2): U with only its diagonal
        a[0] = no[3 * HW + i]; a[3] = no[4 * HW + i]; a[5] = no[5 * HW + i]; a[1] = a[2] = a[4] = 0.0;
      } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) a[c] = no[(3 + c) * HW + i];
      }
      // Sigma_x = U U^T with U = [[a0,a1,a2],[0,a3,a4],[0,0,a5]]   (denoiser.py:246-255)
      Sym3 sx;
      sx.a00 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2]; sx.a01 = a[1] * a[3] + a[2] * a[4]; sx.a02 = a[2] * a[5];
      sx.a11 = a[3] * a[3] + a[4] * a[4]; sx.a12 = a[4] * a[5]; sx.a22 = a[5] * a[5];
      const double n0 = (double)sg[0] * sg[0], n1 = (double)sg[1] * sg[1], n2 = (double)sg[2] * sg[2];
      Sym3 sy = sx; sy.a00 += n0; sy.a11 += n1; sy.a22 += n2;
      const double dsy = det3(sy);
      const Sym3 syi = inv3(sy, dsy);
      const double d0 = y[0] - mu[0], d1 = y[1] - mu[1], d2 = y[2] - mu[2];
      double v0, v1, v2; mv3(syi, d0, d1, d2, v0, v1, v2);
      const double quad = d0 * v0 + d1 * v1 + d2 * v2;
      double l = 0.5 * log(fmax(dsy, 0.0)) + 0.5 * quad;
      if (!known) l -= 0.1 * ((double)sg[0] + sg[1] + sg[2]) / 3.0;
      acc += (float)l;
      // posterior mean (denoiser.py:366-372), literal formulation with the 1e-6 I regularisers
      const double eps = 1e-6;
      Sym3 sxe = sx; sxe.a00 += eps; sxe.a11 += eps; sxe.a22 += eps;
      const Sym3 sxi = inv3(sxe, det3(sxe));
      const double i0 = 1.0 / (n0 + eps), i1 = 1.0 / (n1 + eps), i2 = 1.0 / (n2 + eps);
      Sym3 c1 = sxi; c1.a00 += i0 + eps; c1.a11 += i1 + eps; c1.a22 += i2 + eps;
      const Sym3 c1i = inv3(c1, det3(c1));
      double t0, t1, t2; mv3(sxi, mu[0], mu[1], mu[2], t0, t1, t2);
      t0 += i0 * y[0]; t1 += i1 * y[1]; t2 += i2 * y[2];
      double p0, p1, p2; mv3(c1i, t0, t1, t2, p0, p1, p2);
      float* pp = pme + (long long)n * 3 * HW + i;
      pp[0] = (float)p0; pp[HW] = (float)p1; pp[2 * HW] = (float)p2;
      model_std[(long long)n * HW + i] = (float)pow(fmax(det3(sx), 0.0), 1.0 / 6.0);
    }
  }
  const float t = block_sum(acc, sm);
  if (threadIdx.x == 0) partial[n * gridDim.x + blockIdx.x] = t;
}

// loss[n] = sum(partial[n][:]) / HW ; noise_std_out[n] = (prod sigma_c^2)^(1/6) (RGB) or sigma (mono)
__global__ void posterior_finalize_kernel(const float* __restrict__ partial, int nblk, int HW, const float* __restrict__ sigma_raw,
                                          int cs, int known, int C, int N, float* __restrict__ loss, float* __restrict__ noise_std) {
  pdl_wait();
  // noise_std == NULL for Poisson noise (per-pixel levels were written by the forward kernel)
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int b = 0; b < nblk; ++b) s += partial[n * nblk + b];
  loss[n] = s / (float)HW;
  if (noise_std) {
    if (C == 1) noise_std[n] = sigma_map(sigma_raw[n * cs], known);
    else {
      double p = 1.0;
      for (int c = 0; c < 3; ++c) { const double sgm = sigma_map(sigma_raw[n * cs + (cs == 1 ? 0 : c)], known); p *= sgm * sgm; }
      noise_std[n] = (float)pow(fmax(p, 0.0), 1.0 / 6.0);
    }
  }
}

// ---------------------------------------------------------------- SSDN posterior / NLL, backward
// d(net_out) for  L = sum_n gloss[n] * loss[n];  dsig_partial[n][blk][c] = strip sums of dL/d(sigma_c)
template <int C>
__global__ void posterior_bwd_kernel(const float* __restrict__ net_out, const float* __restrict__ noisy,
                                     const float* __restrict__ sigma_raw, int cs, int known, int poisson, int diag, int HW,
                                     const float* __restrict__ gloss, float* __restrict__ dnet, float* __restrict__ dsig_partial) {
  pdl_wait();
  // ds[c] accumulates d(loss)/d(sigma_c) (Gaussian; the -0.1 regulariser is added by the finalize kernel) or
  // d(loss)/d(k_c) including the regulariser (Poisson: it depends on the pixel)
  __shared__ float sm[32];
  const int CO = diag ? 2 * C : C + C * (C + 1) / 2;
  const int n = blockIdx.y;
  const float* no = net_out + (long long)n * CO * HW;
  const float* yy = noisy + (long long)n * C * HW;
  float* dn = dnet + (long long)n * CO * HW;
  const float scale = gloss[n] / (float)HW;
  float sg[3], kc[3];
#pragma unroll
  for (int c = 0; c < C; ++c) { kc[c] = noise_coeff(sigma_raw[n * cs + (cs == 1 ? 0 : c)], known, poisson); sg[c] = kc[c]; }
  float ds[3] = {0.f, 0.f, 0.f};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    if (C == 1) {
      const float mu = no[i], a = no[HW + i], y = yy[i];
      const float base = fmaxf(mu, 1e-3f);
      if (poisson) sg[0] = sqrtf(base * kc[0]);
      const float sy = a * a + sg[0] * sg[0], d = y - mu;
      const float dsy = -d * d / (sy * sy) + 1.f / sy;
      float dmu = -2.f * d / sy;
      if (poisson) {
        const float reg = known ? 0.f : 0.1f / (2.f * sg[0]);      // d(-0.1 sigma)/d(sigma^2)
        if (mu > 1e-3f) dmu += (dsy - reg) * kc[0];
        ds[0] += scale * (dsy - reg) * base;
      } else {
        ds[0] += scale * dsy * 2.f * sg[0];
      }
      dn[i] = scale * dmu;
      dn[HW + i] = scale * dsy * 2.f * a;
    } else {
      double mu[3], y[3], a[6];
#pragma unroll
      for (int c = 0; c < 3; ++c) { mu[c] = no[c * HW + i]; y[c] = yy[c * HW + i]; }
      if (poisson) {
#pragma unroll
        for (int c = 0; c < 3; ++c) sg[c] = sqrtf(fmaxf((float)mu[c], 1e-3f) * kc[c]);
      }
      if (diag) {
        a[0] = no[3 * HW + i]; a[3] = no[4 * HW + i]; a[5] = no[5 * HW + i]; a[1] = a[2] = a[4] = 0.0;
      } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) a[c] = no[(3 + c) * HW + i];
      }
      Sym3 sy;
      sy.a00 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + (double)sg[0] * sg[0]; sy.a01 = a[1] * a[3] + a[2] * a[4]; sy.a02 = a[2] * a[5];
      sy.a11 = a[3] * a[3] + a[4] * a[4] + (double)sg[1] * sg[1]; sy.a12 = a[4] * a[5]; sy.a22 = a[5] * a[5] + (double)sg[2] * sg[2];
      const double dsy = det3(sy);
      const Sym3 syi = inv3(sy, dsy);
      double v0, v1, v2; mv3(syi, y[0] - mu[0], y[1] - mu[1], y[2] - mu[2], v0, v1, v2);
      const double ld = dsy > 0.0 ? 0.5 : 0.0;   // the log-det term has no gradient once clamped at 0
      // G = dL/dSigma_y = ld * Sigma_y^-1 - 0.5 v v^T   (symmetric)
      Sym3 G;
      G.a00 = ld * syi.a00 - 0.5 * v0 * v0; G.a01 = ld * syi.a01 - 0.5 * v0 * v1; G.a02 = ld * syi.a02 - 0.5 * v0 * v2;
      G.a11 = ld * syi.a11 - 0.5 * v1 * v1; G.a12 = ld * syi.a12 - 0.5 * v1 * v2; G.a22 = ld * syi.a22 - 0.5 * v2 * v2;
      const double s = scale;
      double dmu[3] = {-v0, -v1, -v2};
      if (poisson) {
        const double gd[3] = {G.a00, G.a11, G.a22};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double reg = known ? 0.0 : (0.1 / 3.0) / (2.0 * sg[c]);       // d(-0.1 mean_c sigma_c)/d(sigma_c^2)
          if (mu[c] > 1e-3) dmu[c] += (gd[c] - reg) * kc[c];
          ds[c] += (float)(s * (gd[c] - reg) * fmax(mu[c], 1e-3));
        }
      }
      dn[i] = (float)(s * dmu[0]); dn[HW + i] = (float)(s * dmu[1]); dn[2 * HW + i] = (float)(s * dmu[2]);
      // dL/dU = 2 G U restricted to the upper triangle, U = [[a0,a1,a2],[0,a3,a4],[0,0,a5]]
      if (diag) {
        dn[3 * HW + i] = (float)(2.0 * s * (G.a00 * a[0])); dn[4 * HW + i] = (float)(2.0 * s * (G.a11 * a[3])); dn[5 * HW + i] = (float)(2.0 * s * (G.a22 * a[5]));
      } else {
      dn[3 * HW + i] = (float)(2.0 * s * (G.a00 * a[0]));
      dn[4 * HW + i] = (float)(2.0 * s * (G.a00 * a[1] + G.a01 * a[3]));
      dn[5 * HW + i] = (float)(2.0 * s * (G.a00 * a[2] + G.a01 * a[4] + G.a02 * a[5]));
      dn[6 * HW + i] = (float)(2.0 * s * (G.a01 * a[1] + G.a11 * a[3]));
      dn[7 * HW + i] = (float)(2.0 * s * (G.a01 * a[2] + G.a11 * a[4] + G.a12 * a[5]));
      dn[8 * HW + i] = (float)(2.0 * s * (G.a02 * a[2] + G.a12 * a[4] + G.a22 * a[5]));
      }
      if (!poisson) { ds[0] += (float)(s * 2.0 * G.a00 * sg[0]); ds[1] += (float)(s * 2.0 * G.a11 * sg[1]); ds[2] += (float)(s * 2.0 * G.a22 * sg[2]); }
    }
  }
  if (dsig_partial) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float t = block_sum(ds[c], sm);
      if (threadIdx.x == 0) dsig_partial[(n * gridDim.x + blockIdx.x) * 3 + c] = t;
    }
  }
}

// d(sigma_raw)[n][c'] from the strip partials: adds the -0.1 regulariser and the softplus chain rule.
__global__ void posterior_bwd_finalize_kernel(const float* __restrict__ dsig_partial, int nblk, const float* __restrict__ sigma_raw,
                                              int cs, int C, int N, int poisson, const float* __restrict__ gloss,
                                              float* __restrict__ dsigma_raw) {
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * cs) return;
  const int n = idx / cs, c = idx % cs;
  float s = 0.f;
  const float reg = poisson ? 0.f : 0.1f * gloss[n];      // Poisson: the regulariser is per pixel, already in the partials
  if (cs == 1) { for (int b = 0; b < nblk; ++b) for (int k = 0; k < C; ++k) s += dsig_partial[(n * nblk + b) * 3 + k]; s -= reg; }
  else { for (int b = 0; b < nblk; ++b) s += dsig_partial[(n * nblk + b) * 3 + c]; s -= reg / (float)C; }
  dsigma_raw[idx] = s * sigma_map_grad(sigma_raw[idx]);
}

// ---------------------------------------------------------------- per-sample spatial mean (sigma estimator head)
__global__ void spatial_mean_kernel(const float* __restrict__ x, int HW, float* __restrict__ out) {
  pdl_wait();   // grid = N*C
  __shared__ float sm[32];
  const float* p = x + (long long)blockIdx.x * HW;
  float acc = 0.f;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) acc += p[i];
  const float t = block_sum(acc, sm);
  if (threadIdx.x == 0) out[blockIdx.x] = t / (float)HW;
}
__global__ void spatial_mean_bwd_kernel(const float* __restrict__ g, int HW, long long total, float* __restrict__ dx) {
  pdl_wait();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < total) dx[i] = g[i / HW] / (float)HW;
}

// ---------------------------------------------------------------- MSE per sample (denoiser.py:153-154) and its gradient
__global__ void mse_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int CHW, float* __restrict__ partial) {
  pdl_wait();
  __shared__ float sm[32];   // grid = (blocks_per_sample, N)
  const float* pa = a + (long long)blockIdx.y * CHW; const float* pb = b + (long long)blockIdx.y * CHW;
  float acc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < CHW; i += gridDim.x * blockDim.x) { const float d = pa[i] - pb[i]; acc += d * d; }
  const float t = block_sum(acc, sm);
  if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = t;
}
__global__ void mean_finalize_kernel(const float* __restrict__ partial, int nblk, float denom, int N, float* __restrict__ out) {
  pdl_wait();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int b = 0; b < nblk; ++b) s += partial[n * nblk + b];
  out[n] = s / denom;
}
__global__ void mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ gloss, int CHW,
                               long long total, float* __restrict__ da) {
  pdl_wait();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < total) da[i] = 2.f * (a[i] - b[i]) * gloss[i / CHW] / (float)CHW;
}

// ---------------------------------------------------------------- masked MSE (utils/n2v_loss.py:6-17, denoiser.py:176-178)
// coords: [K][2] int64 of the FIRST sample, indexing [:, :, c0, c1] (row = first coordinate) exactly as the reference.
// One thread per (n, c): sequential over the K coordinates (duplicates accumulate deterministically).
__global__ void masked_mse_fwd_kernel(const float* __restrict__ out, const float* __restrict__ ref, const long long* __restrict__ coords,
                                      int K, int NC, int H, int W, float* __restrict__ per_nc) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NC) return;
  const float* po = out + (long long)i * H * W; const float* pr = ref + (long long)i * H * W;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) { const long long o = coords[2 * k] * W + coords[2 * k + 1]; const float d = pr[o] - po[o]; acc += d * d; }
  per_nc[i] = acc;
}
__global__ void masked_mse_reduce_kernel(const float* __restrict__ per_nc, int N, int C, float* __restrict__ loss) {
  pdl_wait();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += per_nc[n * C + c];
  loss[n] = s / (float)C;
}
__global__ void masked_mse_bwd_kernel(const float* __restrict__ out, const float* __restrict__ ref, const long long* __restrict__ coords,
                                      int K, int NC, int C, int H, int W, const float* __restrict__ gloss, float* __restrict__ dout) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // dout must be zero-filled beforehand
  if (i >= NC) return;
  const float* po = out + (long long)i * H * W; const float* pr = ref + (long long)i * H * W;
  float* pd = dout + (long long)i * H * W;
  const float s = gloss[i / C] / (float)C;
  for (int k = 0; k < K; ++k) { const long long o = coords[2 * k] * W + coords[2 * k + 1]; pd[o] += -2.f * (pr[o] - po[o]) * s; }
}

// ---------------------------------------------------------------- Adam over a flat parameter buffer
// torch.optim.Adam(betas=(0.9, 0.99), eps=1e-8) as used by train.py:100-107; grad_scale folds the 1/world_size of the
// data-parallel all-reduce.  bc1 = 1 - beta1^t, bc2_sqrt = sqrt(1 - beta2^t) are computed on the host in fp64.
__device__ __forceinline__ void adam_update(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                            long long i, float step_size, float beta1, float beta2, float eps, float bc2_sqrt, float grad_scale) {
  const float gr = g[i] * grad_scale;
  const float mm = m[i] * beta1 + (1.f - beta1) * gr;       // exp_avg.mul_(b1).add_(g, alpha=1-b1)   [lerp form equals this to 1 ulp]
  const float vv = v[i] * beta2 + (1.f - beta2) * gr * gr;  // exp_avg_sq.mul_(b2).addcmul_(g, g, value=1-b2)
  m[i] = mm; v[i] = vv;
  const float denom = sqrtf(vv) / bc2_sqrt + eps;
  p[i] = p[i] - step_size * (mm / denom);
}
// skip[0 .. n_skip): "this step's gradients are stale" flags left behind by the networks' backward passes (summed over ranks
// by the gradient all-reduce): any non-zero flag turns the whole update into a no-op
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float step_size, float beta1, float beta2, float eps, float bc2_sqrt, float grad_scale,
                            const float* __restrict__ skip, int n_skip) {
  pdl_wait();
  for (int j = 0; j < n_skip; ++j) if (__ldg(skip + j) != 0.f) return;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  adam_update(p, g, m, v, i, step_size, beta1, beta2, eps, bc2_sqrt, grad_scale);
}
// The same update with the step-dependent scalars read from DEVICE memory (hyper = {lr / bias_correction1, beta1, beta2, eps,
// sqrt(bias_correction2), grad_scale}): a CUDA graph that contains this launch can be replayed for every step.
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                long long n, const float* __restrict__ hyper, const float* __restrict__ skip, int n_skip) {
  pdl_wait();
  for (int j = 0; j < n_skip; ++j) if (__ldg(skip + j) != 0.f) return;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  adam_update(p, g, m, v, i, __ldg(hyper), __ldg(hyper + 1), __ldg(hyper + 2), __ldg(hyper + 3), __ldg(hyper + 4), __ldg(hyper + 5));
}

}  // namespace lossk
