// C-ABI: whole-network forward/backward, pipeline losses, index operators, flat Adam.
// Everything here is asynchronous on `stream` unless noted.  See include/ssdn_b200.h.
#include "input_pipeline.cuh"
#include "loss.cuh"
#include "net.cuh"
#include "../../include/ssdn_b200.h"

using namespace eng;

// ------------------------------------------------------------------------------------ network
extern "C" int ssdn_net_create(int n, int cin, int cout, int h, int w, int blindspot, void** handle) {
  if (!handle) return fail(-1, "null handle");
  if (n <= 0 || cin <= 0 || cout <= 0) return fail(-1, "bad batch/channel count");
  if (h % 32 || w % 32 || h <= 0 || w <= 0) return fail(-1, "input height and width must be positive multiples of 32 (got %dx%d)", h, w);
  if (blindspot && h != w) return fail(-1, "blind-spot network needs square inputs (got %dx%d)", h, w);
  if (cin > 16 || cout > 16) return fail(-1, "at most 16 image channels are supported");
  *handle = new net::Net(n, cin, cout, h, w, blindspot != 0, num_sms());
  return 0;
}
extern "C" void ssdn_net_destroy(void* handle) { delete (net::Net*)handle; }
extern "C" size_t ssdn_net_workspace_bytes(void* handle) { return ((net::Net*)handle)->ws_bytes; }
extern "C" size_t ssdn_net_param_count(void* handle) { return ((net::Net*)handle)->n_params; }
extern "C" int ssdn_net_bind(void* handle, void* ws, size_t bytes, void* stream) {
  return ((net::Net*)handle)->bind(ws, bytes, (cudaStream_t)stream);
}
extern "C" int ssdn_net_forward(void* handle, const float* params, const float* x, float* out, int training, void* stream) {
  return ((net::Net*)handle)->forward(params, x, out, (cudaStream_t)stream, training != 0);
}
extern "C" int ssdn_net_backward(void* handle, const float* params, const float* dout, float* grads, float* stale_out, void* stream) {
  return ((net::Net*)handle)->backward(params, dout, grads, stale_out, (cudaStream_t)stream);
}
// Synchronises the stream and reports the operand-scale status of the last passes: status3 = {forward stale, backward stale,
// stale passes since bind}.  A stale pass ran with fp16 operand scales outside their accurate band and must be repeated
// (the scales it left behind are the right ones).
extern "C" int ssdn_net_scale_status(void* handle, int* status3, void* stream) {
  net::Net* nn = (net::Net*)handle;
  if (!nn->ws) return fail(-5, "network workspace not bound");
  SSDN_CUDA(cudaMemcpyAsync(status3, nn->scales.status, 3 * sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  SSDN_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}
// Test hook: the scale exponent in use and the bit pattern of the running maximum of every slot (96 each).
extern "C" int ssdn_net_debug_scales(void* handle, int* k96, unsigned* amax96, void* stream) {
  net::Net* nn = (net::Net*)handle;
  if (!nn->ws) return fail(-5, "network workspace not bound");
  SSDN_CUDA(cudaMemcpyAsync(k96, nn->scales.k, net::kSlots * sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  SSDN_CUDA(cudaMemcpyAsync(amax96, nn->scales.amax, net::kSlots * sizeof(unsigned), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  SSDN_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}
// Synchronises the stream and reports device-side pipeline errors (bounded mbarrier waits that expired).
extern "C" int ssdn_net_check(void* handle, void* stream) {
  net::Net* nn = (net::Net*)handle;
  if (!nn->ws) return fail(-5, "network workspace not bound");
  int h = 0;
  SSDN_CUDA(cudaMemcpyAsync(&h, nn->flag, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  SSDN_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (h) {
    cudaMemsetAsync(nn->flag, 0, 4, (cudaStream_t)stream);       // report once: the next check sees only new trouble
    return fail(-4, "kernel pipeline timeout (role %d)", h);
  }
  return 0;
}
// Debug/test helper: copies channels [0, c) of a named internal buffer (plane 0 = value = (hi + lo) * 2^-k, 1 = lo, 2 = hi as
// stored) into a dense NCHW tensor [B][c][H][W] of that buffer's own geometry.  With out == NULL only dims is filled.
extern "C" int ssdn_net_debug_read(void* handle, const char* name, int plane, int c, float* out, int* dims, void* stream) {
  net::Net* nn = (net::Net*)handle;
  const net::Buf* b = nn->find_buf(name);
  if (!b) return fail(-1, "unknown buffer '%s'", name);
  if (plane != 0 && !b->hi) return fail(-1, "buffer '%s' is a single fp32 plane", name);
  if (c > b->cpitch) return fail(-1, "buffer '%s' has only %d channels", name, b->cpitch);
  if (dims) { dims[0] = b->g.B; dims[1] = c; dims[2] = b->g.H; dims[3] = b->g.W; }
  if (!out) return 0;
  const long long n = (long long)b->g.B * c * b->g.H * b->g.W;
  if (b->hi) pw::unpack_nchw_kernel<<<pw::grid_for(n), pw::kBlock, 0, (cudaStream_t)stream>>>(b->hi, b->lo, b->sc.k, plane, out, b->g.B, c, b->g.H, b->g.W, b->g, b->cpitch, 0);
  else pw::unpack_nchw_f32_kernel<<<pw::grid_for(n), pw::kBlock, 0, (cudaStream_t)stream>>>(b->v, out, b->g.B, c, b->g.H, b->g.W, b->g, b->cpitch, 0);
  SSDN_CUDA(cudaGetLastError());
  return 0;
}
// Test helper: overwrites channels [0, c) of a named internal buffer (both planes, with the buffer's current scale) from a
// dense NCHW tensor [B][c][H][W].  Used by the parity tests to run the backward pass on the oracle's forward activations.
extern "C" int ssdn_net_debug_write(void* handle, const char* name, int c, const float* src, void* stream) {
  net::Net* nn = (net::Net*)handle;
  net::Buf* b = const_cast<net::Buf*>(nn->find_buf(name));
  if (!b) return fail(-1, "unknown buffer '%s'", name);
  if (c > b->cpitch) return fail(-1, "buffer '%s' has only %d channels", name, b->cpitch);
  const long long n = (long long)b->g.B * c * b->g.H * b->g.W;
  if (b->hi) {
    ScaleRef sc = b->sc; sc.amax = nullptr;
    pw::pack_nchw_kernel<<<pw::grid_for(n), pw::kBlock, 0, (cudaStream_t)stream>>>(src, b->hi, b->lo, b->g.B, c, b->g.H, b->g.W, b->g, b->cpitch, 0, 0, sc);
  } else {
    pw::pack_nchw_f32_kernel<<<pw::grid_for(n), pw::kBlock, 0, (cudaStream_t)stream>>>(src, b->v, b->g.B, c, b->g.H, b->g.W, b->g, b->cpitch, 0);
  }
  if (b->mask)
    pw::mask_from_planes_kernel<<<pw::grid_for(b->g.total() * b->mask_words), pw::kBlock, 0, (cudaStream_t)stream>>>(b->hi, b->g.total(), b->cpitch,
                                                                                                                  b->mask, b->mask_words);
  SSDN_CUDA(cudaGetLastError());
  return 0;
}
// Per-launch timing of every engine kernel (bench.py roofline).  profile_begin() switches CUDA-event bracketing of every launch
// on (everything then runs on one stream); profile_end() synchronises and returns, per kind (ProfKind, conv_igemm.cuh), the
// launch count, the summed device time in ms, the summed algorithmic FLOPs and the summed algorithmic HBM bytes:
// out[kind*4 + {0,1,2,3}] for K_COUNT = 15 kinds.
extern "C" int ssdn_profile_begin(void) { profiler().recs.clear(); profiler().on = true; return 0; }
static std::vector<double> g_last_records;   // (kind, ms, flops, bytes) per launch of the last profiled region
extern "C" int ssdn_profile_kinds(void) { return K_COUNT; }
extern "C" int ssdn_profile_end(double* out) {
  LaunchProfiler& pr = profiler();
  pr.on = false;
  for (int i = 0; i < 4 * K_COUNT; ++i) out[i] = 0;
  SSDN_CUDA(cudaDeviceSynchronize());
  g_last_records.clear();
  for (auto& r : pr.recs) {
    float ms = 0; cudaEventElapsedTime(&ms, r.a, r.b);
    out[r.kind * 4 + 0] += 1; out[r.kind * 4 + 1] += ms; out[r.kind * 4 + 2] += r.flops; out[r.kind * 4 + 3] += r.bytes;
    g_last_records.push_back(r.kind); g_last_records.push_back(ms); g_last_records.push_back(r.flops); g_last_records.push_back(r.bytes);
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  pr.recs.clear();
  return 0;
}
// Per-launch records of the last profiled region, in launch order: out[4*i + {0,1,2,3}] = {kind, ms, algorithmic FLOPs, bytes}.
// Returns the number of records (copies at most max_records).
extern "C" int ssdn_profile_records(double* out, int max_records) {
  const int n = (int)(g_last_records.size() / 4);
  for (int i = 0; i < n && i < max_records; ++i) for (int k = 0; k < 4; ++k) out[4 * i + k] = g_last_records[4 * i + k];
  return n;
}
extern "C" int ssdn_net_kernel_launches(void* handle, int training) {
  net::Net* nn = (net::Net*)handle;
  const int nl = (int)nn->layers.size();
  int fwd = nl /*conv*/ + 5 /*pool*/ + 1 /*pack: rotation stack + im2col of the first conv*/ + 2 /*weight scales, weight slabs*/ + 1 /*scale finish*/;
  int bwd = 3 /*loss-gradient scale, pack, column sums*/ + nl /*wgrad*/ + 2 /*split-K reductions*/ + 1 /*all bias reductions*/ + (nl - 1) /*dgrad*/ +
            5 + 5 /*pool, upsample*/ + 1 /*scale finish*/;
  return training ? fwd + bwd : fwd;
}

// ------------------------------------------------------------------------------------ index operators (standalone)
namespace {
__global__ void rot4_stack_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int C, int H, int W) {
  const long long n = 4LL * B * C * H * W;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int j = (int)(idx % W); long long t = idx / W;
  const int i = (int)(t % H); t /= H;
  const int c = (int)(t % C); const int bo = (int)(t / C);
  const int r = bo / B, b = bo % B;
  int si = i, sj = j;
  if (r == 1) { si = j; sj = W - 1 - i; } else if (r == 2) { si = H - 1 - i; sj = W - 1 - j; } else if (r == 3) { si = H - 1 - j; sj = i; }
  y[idx] = x[(((long long)b * C + c) * H + si) * W + sj];
}
// y[n, r*C + c, i, j] = rotate(shift_down_1(x[r*N + n]), inv_angle_r)[c, i, j]
__global__ void shift_unrot_concat_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int C, int H, int W) {
  const long long n = 4LL * N * C * H * W;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int j = (int)(idx % W); long long t = idx / W;
  const int i = (int)(t % H); t /= H;
  const int cc = (int)(t % (4 * C)); const int b = (int)(t / (4 * C));
  const int r = cc / C, c = cc % C;
  int p, q;   // position in the shifted branch image
  if (r == 0) { p = i; q = j; } else if (r == 1) { p = H - 1 - j; q = i; } else if (r == 2) { p = H - 1 - i; q = W - 1 - j; } else { p = j; q = W - 1 - i; }
  y[idx] = p == 0 ? 0.f : x[((((long long)r * N + b) * C + c) * H + (p - 1)) * W + q];
}
}  // namespace

extern "C" int ssdn_rot4_stack(const float* x, float* y, int n, int c, int h, int w, void* stream) {
  if (h != w) return fail(-1, "rot4_stack needs square images");
  rot4_stack_kernel<<<pw::grid_for(4LL * n * c * h * w), pw::kBlock, 0, (cudaStream_t)stream>>>(x, y, n, c, h, w);
  SSDN_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int ssdn_shift_unrot_concat(const float* x, float* y, int n, int c, int h, int w, void* stream) {
  if (h != w) return fail(-1, "shift_unrot_concat needs square images");
  shift_unrot_concat_kernel<<<pw::grid_for(4LL * n * c * h * w), pw::kBlock, 0, (cudaStream_t)stream>>>(x, y, n, c, h, w);
  SSDN_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------ losses
static inline int loss_blocks(int hw) { int b = (hw + 511) / 512; return b < 1 ? 1 : (b > 64 ? 64 : b); }

extern "C" size_t ssdn_loss_workspace_bytes(int n, int c) { return (size_t)n * 64 * 4 * sizeof(float) + (size_t)n * c * sizeof(float) + 1024; }

extern "C" int ssdn_posterior_forward(void* ws, const float* net_out, const float* noisy, const float* sigma_raw, int n, int c, int h,
                                      int w, int cs, int sigma_known, int noise_model, float* pme, float* loss, float* model_std,
                                      float* noise_std, void* stream) {
  if (c != 1 && c != 3) return fail(-1, "num_channels must be 1 or 3");
  if (cs != 1 && cs != c) return fail(-1, "sigma must have 1 or C values per sample");
  if (noise_model & ~3) return fail(-1, "noise_model: bit 0 = Poisson noise, bit 1 = diagonal covariance");
  const int poisson = noise_model & 1, diag = (noise_model >> 1) & 1;
  cudaStream_t st = (cudaStream_t)stream;
  const int hw = h * w, nblk = loss_blocks(hw);
  float* partial = (float*)ws;
  dim3 grid(nblk, n);
  float* ns_px = poisson ? noise_std : nullptr;     // Poisson: noise_std is [n][h][w]; Gaussian: [n]
  // per pixel: net_out (c + c(c+1)/2, or 2c with a diagonal covariance) + noisy c in, pme c + model_std 1 out
  const double fbytes = (double)n * hw * 4.0 * ((diag ? 2 * c : c + c * (c + 1) / 2) + c + c + 1);
  if (c == 1) SSDN_PROF(K_POSTERIOR_FWD, 0, fbytes, st, (lossk::posterior_fwd_kernel<1><<<grid, lossk::kBlock, 0, st>>>(net_out, noisy, sigma_raw, cs, sigma_known, poisson, diag, hw, pme, model_std, ns_px, partial)));
  else SSDN_PROF(K_POSTERIOR_FWD, 0, fbytes, st, (lossk::posterior_fwd_kernel<3><<<grid, lossk::kBlock, 0, st>>>(net_out, noisy, sigma_raw, cs, sigma_known, poisson, diag, hw, pme, model_std, ns_px, partial)));
  lossk::posterior_finalize_kernel<<<(n + 127) / 128, 128, 0, st>>>(partial, nblk, hw, sigma_raw, cs, sigma_known, c, n, loss,
                                                                     poisson ? nullptr : noise_std);
  SSDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ssdn_posterior_backward(void* ws, const float* net_out, const float* noisy, const float* sigma_raw, const float* gloss,
                                       int n, int c, int h, int w, int cs, int sigma_known, int noise_model, float* dnet, float* dsigma_raw,
                                       void* stream) {
  if (c != 1 && c != 3) return fail(-1, "num_channels must be 1 or 3");
  if (noise_model & ~3) return fail(-1, "noise_model: bit 0 = Poisson noise, bit 1 = diagonal covariance");
  const int poisson = noise_model & 1, diag = (noise_model >> 1) & 1;
  cudaStream_t st = (cudaStream_t)stream;
  const int hw = h * w, nblk = loss_blocks(hw);
  float* partial = (float*)ws;
  dim3 grid(nblk, n);
  float* dp = (dsigma_raw && !sigma_known) ? partial : nullptr;
  // per pixel: net_out + noisy in, d(net_out) out
  const double bbytes = (double)n * hw * 4.0 * (2 * (diag ? 2 * c : c + c * (c + 1) / 2) + c);
  if (c == 1) SSDN_PROF(K_POSTERIOR_BWD, 0, bbytes, st, (lossk::posterior_bwd_kernel<1><<<grid, lossk::kBlock, 0, st>>>(net_out, noisy, sigma_raw, cs, sigma_known, poisson, diag, hw, gloss, dnet, dp)));
  else SSDN_PROF(K_POSTERIOR_BWD, 0, bbytes, st, (lossk::posterior_bwd_kernel<3><<<grid, lossk::kBlock, 0, st>>>(net_out, noisy, sigma_raw, cs, sigma_known, poisson, diag, hw, gloss, dnet, dp)));
  if (dp) lossk::posterior_bwd_finalize_kernel<<<(n * cs + 127) / 128, 128, 0, st>>>(partial, nblk, sigma_raw, cs, c, n, poisson, gloss, dsigma_raw);
  SSDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ssdn_spatial_mean_forward(const float* x, int rows, int hw, float* out, void* stream) {
  lossk::spatial_mean_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(x, hw, out);
  SSDN_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int ssdn_spatial_mean_backward(const float* g, int rows, int hw, float* dx, void* stream) {
  const long long total = (long long)rows * hw;
  lossk::spatial_mean_bwd_kernel<<<pw::grid_for(total), pw::kBlock, 0, (cudaStream_t)stream>>>(g, hw, total, dx);
  SSDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ssdn_mse_forward(void* ws, const float* a, const float* b, int n, int chw, float* loss, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = loss_blocks(chw);
  dim3 grid(nblk, n);
  lossk::mse_fwd_kernel<<<grid, lossk::kBlock, 0, st>>>(a, b, chw, (float*)ws);
  lossk::mean_finalize_kernel<<<(n + 127) / 128, 128, 0, st>>>((float*)ws, nblk, (float)chw, n, loss);
  SSDN_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int ssdn_mse_backward(const float* a, const float* b, const float* gloss, int n, int chw, float* da, void* stream) {
  const long long total = (long long)n * chw;
  lossk::mse_bwd_kernel<<<pw::grid_for(total), pw::kBlock, 0, (cudaStream_t)stream>>>(a, b, gloss, chw, total, da);
  SSDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ssdn_masked_mse_forward(void* ws, const float* out, const float* ref, const long long* coords, int k, int n, int c, int h,
                                       int w, float* loss, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  float* per_nc = (float*)ws;
  lossk::masked_mse_fwd_kernel<<<(n * c + 127) / 128, 128, 0, st>>>(out, ref, coords, k, n * c, h, w, per_nc);
  lossk::masked_mse_reduce_kernel<<<(n + 127) / 128, 128, 0, st>>>(per_nc, n, c, loss);
  SSDN_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int ssdn_masked_mse_backward(const float* out, const float* ref, const long long* coords, int k, const float* gloss, int n,
                                        int c, int h, int w, float* dout, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SSDN_CUDA(cudaMemsetAsync(dout, 0, (size_t)n * c * h * w * sizeof(float), st));
  lossk::masked_mse_bwd_kernel<<<(n * c + 127) / 128, 128, 0, st>>>(out, ref, coords, k, n * c, c, h, w, gloss, dout);
  SSDN_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------ optimiser
#include <math.h>
extern "C" int ssdn_adam_step(float* p, const float* g, float* m, float* v, long long count, double lr, double beta1, double beta2,
                              double eps, long long step, double grad_scale, const float* skip, int n_skip, void* stream) {
  if (step < 1) return fail(-1, "Adam step count starts at 1");
  if (n_skip < 0 || n_skip > 8) return fail(-1, "at most 8 skip flags");
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  // 16 bytes read + 12 written per parameter
  SSDN_PROF(K_ADAM, 0, 28.0 * count, (cudaStream_t)stream,
            (lossk::adam_kernel<<<pw::grid_for(count), pw::kBlock, 0, (cudaStream_t)stream>>>(p, g, m, v, count, (float)(lr / bc1), (float)beta1, (float)beta2,
                                                                                            (float)eps, (float)sqrt(bc2), (float)grad_scale, skip, skip ? n_skip : 0)));
  SSDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ssdn_adam_step_dev(float* p, const float* g, float* m, float* v, long long count, const float* hyper6, const float* skip,
                                  int n_skip, void* stream) {
  if (!hyper6) return fail(-1, "null hyper-parameter buffer");
  if (n_skip < 0 || n_skip > 8) return fail(-1, "at most 8 skip flags");
  SSDN_PROF(K_ADAM, 0, 28.0 * count, (cudaStream_t)stream,
            (lossk::adam_dev_kernel<<<pw::grid_for(count), pw::kBlock, 0, (cudaStream_t)stream>>>(p, g, m, v, count, hyper6, skip, skip ? n_skip : 0)));
  SSDN_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------ on-GPU input pipeline
extern "C" int ssdn_noisy_crops(const unsigned char* images, int n_images, int c, int h, int w, const int* order, int n, int patch,
                                unsigned long long seed, unsigned long long step, int stream_id, float sigma_lo, float sigma_hi, int clip,
                                float* clean, float* noisy, float* sigma, void* stream) {
  if (!images || !noisy) return fail(-1, "null image cache / output");
  if (c < 1 || c > 4) return fail(-1, "1..4 image channels are supported");
  if (n_images <= 0 || n <= 0 || patch <= 0 || patch > h || patch > w) return fail(-1, "patch %d does not fit %dx%d images", patch, h, w);
  if (sigma_lo < 0.f || sigma_hi < 0.f) return fail(-1, "negative noise level");
  const long long total = (long long)n * patch * patch;
  inpk::noisy_crops_kernel<<<pw::grid_for(total), pw::kBlock, 0, (cudaStream_t)stream>>>(
      images, n_images, c, h, w, order, n, patch, (uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)step, (uint32_t)(step >> 32),
      (uint32_t)stream_id, sigma_lo, sigma_hi, clip, clean, noisy, sigma);
  SSDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ssdn_poisson_crops(const unsigned char* images, int n_images, int c, int h, int w, const int* order, int n, int patch,
                                  unsigned long long seed, unsigned long long step, int stream_id, float lam_lo, float lam_hi, int clip,
                                  float* clean, float* noisy, float* lam, void* stream) {
  if (!images || !noisy) return fail(-1, "null image cache / output");
  if (c < 1 || c > 4) return fail(-1, "1..4 image channels are supported");
  if (n_images <= 0 || n <= 0 || patch <= 0 || patch > h || patch > w) return fail(-1, "patch %d does not fit %dx%d images", patch, h, w);
  if (!(lam_lo > 0.f) || !(lam_hi > 0.f)) return fail(-1, "the Poisson scale must be positive");
  const long long total = (long long)n * patch * patch;
  inpk::poisson_crops_kernel<<<pw::grid_for(total), pw::kBlock, 0, (cudaStream_t)stream>>>(
      images, n_images, c, h, w, order, n, patch, (uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)step, (uint32_t)(step >> 32),
      (uint32_t)stream_id, lam_lo, lam_hi, clip, clean, noisy, lam);
  SSDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int ssdn_n2v_mask(const float* noisy, float* masked, long long* coords, int n, int c, int h, int w, int subpatch_size,
                             unsigned long long seed, unsigned long long step, void* stream) {
  if (!noisy || !masked || !coords) return fail(-1, "null pointer");
  if (subpatch_size % 2 == 0) return fail(-1, "subpatch_size must be odd");
  const int box = 8;                                  // round(sqrt(100 / 1.5)), utils/n2v_ups.py:72-73
  if (h % box || w % box) return fail(-1, "the on-GPU Noise2Void mask needs height and width that are multiples of %d", box);
  cudaStream_t st = (cudaStream_t)stream;
  if (masked != noisy) SSDN_CUDA(cudaMemcpyAsync(masked, noisy, (size_t)n * c * h * w * sizeof(float), cudaMemcpyDeviceToDevice, st));
  const int total = n * (h / box) * (w / box);
  inpk::n2v_mask_kernel<<<pw::grid_for(total), pw::kBlock, 0, st>>>(noisy, masked, coords, n, c, h, w, box, subpatch_size / 2, (uint32_t)seed,
                                                                      (uint32_t)(seed >> 32), (uint32_t)step, (uint32_t)(step >> 32));
  SSDN_CUDA(cudaGetLastError());
  return 0;
}
