// On-GPU input pipeline: random crop + uint8 -> float + synthetic Gaussian noise, from an image cache in HBM.
// Reference: train.py:756-760 (RandomCrop), datasets/noise_wrapper.py:98-163 (prepare_input), utils/noise.py:14-63
// (add_gaussian).  At B200 step rates the reference's 4-worker PIL / h5py loader is the bottleneck by orders of magnitude;
// this kernel produces a batch in a few microseconds.  Randomness is counter-based (Philox4x32-10 keyed by seed, indexed
// by step / sample / pixel): reproducible and independent of launch geometry; parity with the CPU generator is
// statistical, not bit-wise.  HBM-bound: c bytes read + 8c bytes written per output pixel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace inpk {

struct Philox { uint32_t x[4]; };

__device__ __forceinline__ Philox philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0, hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return Philox{{c0, c1, c2, c3}};
}
__device__ __forceinline__ float u01(uint32_t x) { return ((float)x + 1.0f) * 2.3283064365386963e-10f; }   // (0, 1]
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  const float r = sqrtf(-2.0f * __logf(u01(a))), t = 6.283185307179586f * u01(b);
  float s, c;
  __sincosf(t, &s, &c);
  z0 = r * c; z1 = r * s;
}

// one thread per output pixel, all (<= 4) channels.  stream_id separates independent draws of the same step (the second
// noisy realisation used as the Noise2Noise reference).
__global__ void noisy_crops_kernel(const unsigned char* __restrict__ images, int n_images, int C, int H, int W, const int* __restrict__ order,
                                   int n, int patch, uint32_t seed_lo, uint32_t seed_hi, uint32_t step_lo, uint32_t step_hi, uint32_t stream_id,
                                   float sigma_lo, float sigma_hi, int clip, float* __restrict__ clean, float* __restrict__ noisy,
                                   float* __restrict__ sigma) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int pp = patch * patch;
  if (idx >= (long long)n * pp) return;
  const int s = (int)(idx / pp), pix = (int)(idx - (long long)s * pp);
  const int y = pix / patch, x = pix - y * patch;
  // per-sample draws (identical in every thread of the sample): crop origin - shared by all streams of a step, so that a
  // Noise2Noise pair sees the same crop - and the noise level of each channel
  const Philox cs = philox4x32_10(step_lo, step_hi, (uint32_t)s, 0xFFFFFFFFu, seed_lo, seed_hi);
  const int oy = (int)(((unsigned long long)cs.x[0] * (unsigned)(H - patch + 1)) >> 32);
  const int ox = (int)(((unsigned long long)cs.x[1] * (unsigned)(W - patch + 1)) >> 32);
  const Philox ss = philox4x32_10(step_lo, step_hi, (uint32_t)s, 0xFFFFFFFEu - stream_id, seed_lo, seed_hi);
  const int img = order ? order[s] : (int)((step_lo * (unsigned)n + (unsigned)s) % (unsigned)n_images);
  const Philox pn = philox4x32_10(step_lo, step_hi ^ (stream_id << 24), (uint32_t)s, (uint32_t)pix, seed_lo, seed_hi);
  float z[4];
  box_muller(pn.x[0], pn.x[1], z[0], z[1]);
  box_muller(pn.x[2], pn.x[3], z[2], z[3]);
  const unsigned char* src = images + ((long long)img * C * H + (oy + y)) * W + (ox + x);
  for (int c = 0; c < C; ++c) {
    const float sg = sigma_hi > sigma_lo ? sigma_lo + (sigma_hi - sigma_lo) * u01(ss.x[c & 3]) : sigma_lo;
    const float v = (float)src[(long long)c * H * W] / 255.0f;        // torchvision ToTensor: uint8 / 255
    float nv = v + z[c & 3] * sg;
    if (clip) nv = fminf(fmaxf(nv, 0.0f), 1.0f);
    const long long o = ((long long)s * C + c) * pp + pix;
    if (clean) clean[o] = v;
    noisy[o] = nv;
    if (sigma && pix == 0) sigma[s * C + c] = sg;
  }
}

// Poisson(1) by inversion on one 32-bit uniform: k = #{i : x >= floor(CDF(i) * 2^32)}, CDF(i) = sum_{j<=i} e^-1 / j!.
// The tail beyond k = 12 has probability 6e-11 < 2^-32 and is folded into k = 12.
__device__ __forceinline__ int poisson1(uint32_t x) {
  constexpr uint32_t T[12] = {0x5E2D58D8u, 0xBC5AB1B1u, 0xEB715E1Du, 0xFB239797u, 0xFF1025F5u, 0xFFD90F3Bu,
                              0xFFFA8B71u, 0xFFFF540Cu, 0xFFFFED1Fu, 0xFFFFFE21u, 0xFFFFFFD4u, 0xFFFFFFFCu};
  int k = 0;
#pragma unroll
  for (int i = 0; i < 12; ++i) k += x >= T[i] ? 1 : 0;
  return k;
}

// The reference's Poisson styles (utils/noise.py:66-109, 'poisson30', 'poisson5_50'): noisy = (clean * lam + K) / lam with
// K ~ Poisson(1) drawn per element - the rate of the generator is the constant 1 in the reference, NOT clean * lam; kept as
// is - and lam = lam_lo, or U(lam_lo, lam_hi) per sample AND channel for the range styles (leading axis of a CHW image).
// Same thread mapping, crop and image draws as noisy_crops_kernel, so both kernels cut identical crops at a given
// (seed, step, sample).
__global__ void poisson_crops_kernel(const unsigned char* __restrict__ images, int n_images, int C, int H, int W, const int* __restrict__ order,
                                     int n, int patch, uint32_t seed_lo, uint32_t seed_hi, uint32_t step_lo, uint32_t step_hi, uint32_t stream_id,
                                     float lam_lo, float lam_hi, int clip, float* __restrict__ clean, float* __restrict__ noisy,
                                     float* __restrict__ lam_out) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int pp = patch * patch;
  if (idx >= (long long)n * pp) return;
  const int s = (int)(idx / pp), pix = (int)(idx - (long long)s * pp);
  const int y = pix / patch, x = pix - y * patch;
  const Philox cs = philox4x32_10(step_lo, step_hi, (uint32_t)s, 0xFFFFFFFFu, seed_lo, seed_hi);
  const int oy = (int)(((unsigned long long)cs.x[0] * (unsigned)(H - patch + 1)) >> 32);
  const int ox = (int)(((unsigned long long)cs.x[1] * (unsigned)(W - patch + 1)) >> 32);
  const Philox ss = philox4x32_10(step_lo, step_hi, (uint32_t)s, 0xFFFFFFFEu - stream_id, seed_lo, seed_hi);
  const int img = order ? order[s] : (int)((step_lo * (unsigned)n + (unsigned)s) % (unsigned)n_images);
  const Philox pn = philox4x32_10(step_lo, step_hi ^ (stream_id << 24), (uint32_t)s, (uint32_t)pix, seed_lo, seed_hi);
  const unsigned char* src = images + ((long long)img * C * H + (oy + y)) * W + (ox + x);
  for (int c = 0; c < C; ++c) {
    const float lam = lam_hi > lam_lo ? lam_lo + (lam_hi - lam_lo) * u01(ss.x[c & 3]) : lam_lo;
    const float v = (float)src[(long long)c * H * W] / 255.0f;
    float nv = (v * lam + (float)poisson1(pn.x[c & 3])) / lam;          // mul_, add_, div_ of the reference, in that order
    if (clip) nv = fminf(fmaxf(nv, 0.0f), 1.0f);
    const long long o = ((long long)s * C + c) * pp + pix;
    if (clean) clean[o] = v;
    noisy[o] = nv;
    if (lam_out && pix == 0) lam_out[s * C + c] = lam;
  }
}

// Noise2Void masking (utils/n2v_ups.py:7-49, "uniform pixel selection"): one stratified coordinate per 8 x 8 box
// (get_stratified_coords: box = round(sqrt(100 / 1.5)) = 8), each replaced by another pixel of the same image whose column is
// drawn from [min(x - r, 0), min(x + r, W - 1)) \ {x} and row from [min(y - r, 0), min(y + r, H - 1)) \ {y} - the reference's
// `min` where `max` was meant is kept, negative indices wrap like Python's.  One thread per (sample, box); replacements read
// the UNMASKED image (the reference updates in place, so a source pixel that is itself an earlier mask position differs -
// probability ~1.5 %): statistical, not bit, parity.  coords[n][box] = (x, y) as int64, the order of the reference's list.
__global__ void n2v_mask_kernel(const float* __restrict__ noisy, float* __restrict__ masked, long long* __restrict__ coords, int n, int C,
                                int H, int W, int box, int radius, uint32_t seed_lo, uint32_t seed_hi, uint32_t step_lo, uint32_t step_hi) {
  const int by = H / box, bx = W / box, nb = by * bx;        // the reference's outer loop runs over the x axis (shape[0] = W)
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * nb) return;
  const int s = idx / nb, b = idx - s * nb;
  const int i = b / by, j = b - i * by;                       // i: box along x, j: box along y
  Philox r = philox4x32_10(step_lo, step_hi, (uint32_t)s, 0x80000000u + (uint32_t)b, seed_lo, seed_hi);
  const int x = i * box + (int)(((unsigned long long)r.x[0] * (unsigned)box) >> 32);
  const int y = j * box + (int)(((unsigned long long)r.x[1] * (unsigned)box) >> 32);
  auto draw = [&](int centre, int size, uint32_t salt) {
    const int lo = min(centre - radius, 0), hi = min(centre + radius, size - 1);      // torch.randint(lo, hi): [lo, hi)
    int v = centre;
    for (uint32_t t = 0; v == centre && t < 64; ++t) {
      const Philox q = philox4x32_10(step_lo ^ salt, step_hi + t, (uint32_t)s, 0xC0000000u + (uint32_t)b, seed_lo, seed_hi);
      v = lo + (int)(((unsigned long long)q.x[0] * (unsigned)(hi - lo)) >> 32);
    }
    return v < 0 ? v + size : v;                                                      // Python-style negative index
  };
  const int rx = draw(x, W, 0x5bd1e995u), ry = draw(y, H, 0x1b873593u);
  for (int c = 0; c < C; ++c) {
    const long long base = ((long long)s * C + c) * H * W;
    masked[base + (long long)y * W + x] = noisy[base + (long long)ry * W + rx];
  }
  coords[((long long)s * nb + b) * 2 + 0] = x;
  coords[((long long)s * nb + b) * 2 + 1] = y;
}

}  // namespace inpk
