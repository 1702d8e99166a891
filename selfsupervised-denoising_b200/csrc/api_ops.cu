// C-ABI: operator-level entry points (one reference operator per call).  See include/ssdn_b200.h.
// These are synchronous on `stream` (they end with a stream sync to report device-side errors);
// the whole-network entry points in api_net.cu are asynchronous.
#include "engine.cuh"
#include "../../include/ssdn_b200.h"

using namespace eng;

extern "C" const char* ssdn_b200_last_error(void) { return err_buf(); }
extern "C" int ssdn_b200_version(void) { return 100; }

namespace {

struct ConvOpLayout {
  Geom g; int cin_pitch; int cout_padded, N; ConvTaps taps;
  size_t act_halves, slab_halves;
};

ConvOpLayout conv_layout(int n, int cin, int h, int w, int cout, int ksize, bool blind, bool dgrad) {
  ConvOpLayout L;
  L.g = make_geom(n, h, w, ksize == 3);
  L.cin_pitch = round_up(cin, 8);
  L.cout_padded = round_up(cout, 16);
  L.N = pick_n(L.cout_padded);
  L.taps = make_taps(ksize, blind, dgrad, L.g.P);
  L.act_halves = ((size_t)L.g.total() * L.cin_pitch + 127) / 128 * 128;
  L.slab_halves = conv_weight_slab_halves(cin, L.cout_padded, L.taps.n);
  return L;
}

// scale slots of a single-operator call (both operands are leaves: exact scales, computed before packing)
pw::ScaleState take_scales(Arena& a) {
  pw::ScaleState st{};
  st.k = a.take<int>(4); st.k_next = a.take<int>(4); st.amax = a.take<unsigned>(4); st.status = a.take<int>(8); st.counter = a.take<unsigned>(1);
  return st;
}
cudaError_t zero_scales(const pw::ScaleState& st, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(st.k, 0, 16, s);
  if (e == cudaSuccess) e = cudaMemsetAsync(st.k_next, 0, 16, s);
  if (e == cudaSuccess) e = cudaMemsetAsync(st.amax, 0, 16, s);
  if (e == cudaSuccess) e = cudaMemsetAsync(st.counter, 0, 4, s);
  return e;
}

// y[n,cout,h,w] = act( conv(x[n,cin,h,w], slab) + bias ); the slab orientation decides fwd vs dgrad
int run_conv(void* ws, size_t ws_bytes, const float* x, const float* w, const float* bias, float* y, int n, int cin,
             int h, int wd, int cout, int ksize, bool blind, bool dgrad, int lrelu_act, int w_cout, int w_cin,
             cudaStream_t st) {
  ConvOpLayout L = conv_layout(n, cin, h, wd, cout, ksize, blind, dgrad);
  Arena a(ws, ws_bytes);
  __half* planes = a.take<__half>(2 * L.act_halves + 256);
  __half* slab = a.take<__half>(L.slab_halves);
  pw::ScaleState sc = take_scales(a);
  int* flag = a.take<int>(1);
  if (!ws) return (int)0;
  if (!a.ok()) return fail(-3, "workspace too small: need %zu bytes, have %zu", a.off, ws_bytes);
  __half* ah = planes; __half* al = planes + L.act_halves;
  SSDN_CUDA(cudaMemsetAsync(planes, 0, (2 * L.act_halves + 256) * sizeof(__half), st));
  SSDN_CUDA(cudaMemsetAsync(flag, 0, 4, st));
  SSDN_CUDA(zero_scales(sc, st));
  const long long ne = (long long)n * cin * h * wd;
  pw::leaf_scale_kernel<<<64, 256, 0, st>>>(x, ne, sc, 0, 0, 0);
  pw::WeightScaleJobs wj{}; wj.j[0] = {w, w_cout * w_cin * ksize * ksize, 1};
  pw::weight_scale_kernel<<<dim3(pw::kWeightScaleBlocks, 1), 256, 0, st>>>(wj, sc, 0, 0);
  ScaleRef xs{sc.k, nullptr};
  pw::pack_nchw_kernel<<<pw::grid_for(ne), pw::kBlock, 0, st>>>(x, ah, al, n, cin, h, wd, L.g, L.cin_pitch, 0, 0, xs);
  const bool wide = conv_is_wide(cin, L.taps.n);
  int nc, kl; conv_chunks(cin, &nc, &kl, wide);
  const int n_tiles = L.cout_padded / L.N;
  const long long ns = (long long)L.slab_halves / 2 / 8;
  pw::weight_prep_kernel<<<pw::grid_for(ns), pw::kBlock, 0, st>>>(w, slab, w_cout, w_cin, L.taps.n, cout, cin, n_tiles, nc,
                                                                  L.N, dgrad ? 1 : 0, wide ? 64 : 32, sc, 1);
  ConvDst d{};
  d.v = y; d.hi = d.lo = nullptr; d.cpitch = 0; d.coff = 0; d.g = L.g; d.map = MAP_NCHW;
  d.flags = (bias ? EP_BIAS : 0) | (lrelu_act ? EP_LRELU : 0); d.cvalid = cout; d.nimg = n; d.bias = bias;
  ConvPlan plan;
  int r = conv_plan_init(&plan, L.g, ah, al, L.cin_pitch, 0, cin, slab, L.cout_padded, L.N, L.taps, d, flag, num_sms(), sc.k, sc.k + 1);
  if (r) return fail(r, "conv_plan_init failed (%d)", r);
  SSDN_CUDA(conv_launch(plan, st));
  int hflag = 0;
  SSDN_CUDA(cudaMemcpyAsync(&hflag, flag, 4, cudaMemcpyDeviceToHost, st));
  SSDN_CUDA(cudaStreamSynchronize(st));
  if (hflag) return fail(-4, "conv kernel pipeline timeout (role %d)", hflag);
  return 0;
}

size_t conv_ws_bytes(int n, int cin, int h, int w, int cout, int ksize, bool blind, bool dgrad) {
  ConvOpLayout L = conv_layout(n, cin, h, w, cout, ksize, blind, dgrad);
  Arena a(nullptr, 0);
  a.take<__half>(2 * L.act_halves + 256); a.take<__half>(L.slab_halves);
  take_scales(a);
  a.take<int>(1);
  return a.off;
}

}  // namespace

extern "C" size_t ssdn_conv2d_workspace_bytes(int n, int cin, int h, int w, int cout, int ksize) {
  size_t a = conv_ws_bytes(n, cin, h, w, cout, ksize, true, false), b = conv_ws_bytes(n, cout, h, w, cin, ksize, true, true);
  return a > b ? a : b;
}

extern "C" int ssdn_conv2d_forward(void* ws, size_t ws_bytes, const float* x, const float* w, const float* bias, float* y,
                                   int n, int cin, int h, int wd, int cout, int ksize, int blind, int lrelu_act,
                                   void* stream) {
  if (ksize != 1 && ksize != 3) return fail(-1, "ksize must be 1 or 3");
  if (n <= 0 || cin <= 0 || cout <= 0 || h <= 0 || wd <= 0) return fail(-1, "bad shape");
  return run_conv(ws, ws_bytes, x, w, bias, y, n, cin, h, wd, cout, ksize, blind != 0, false, lrelu_act, cout, cin,
                  (cudaStream_t)stream);
}

extern "C" int ssdn_conv2d_backward_data(void* ws, size_t ws_bytes, const float* dy, const float* w, float* dx, int n,
                                         int cin, int h, int wd, int cout, int ksize, int blind, void* stream) {
  if (ksize != 1 && ksize != 3) return fail(-1, "ksize must be 1 or 3");
  // the data-gradient is a convolution of dy (cout channels) producing cin channels
  return run_conv(ws, ws_bytes, dy, w, nullptr, dx, n, cout, h, wd, cin, ksize, blind != 0, true, 0, cout, cin,
                  (cudaStream_t)stream);
}

static size_t wgrad_op_layout(Arena& a, const Geom& g, int cin, int cout, int ntaps, int ks, __half** xp, size_t* xh, __half** yp, size_t* yh,
                              float** partial, float** colpart, pw::ScaleState* sc, int** flag) {
  *xh = ((size_t)g.total() * round_up(cin, 8) + 127) / 128 * 128;
  *yh = ((size_t)g.total() * round_up(cout, 8) + 127) / 128 * 128;
  *xp = a.take<__half>(2 * *xh + 256); *yp = a.take<__half>(2 * *yh + 256);
  *partial = a.take<float>(wgrad_partial_floats(ks, ntaps, cout, cin));
  *colpart = a.take<float>((size_t)1024 * round_up(cout, 8));
  *sc = take_scales(a);
  *flag = a.take<int>(1);
  return a.off;
}

extern "C" size_t ssdn_conv2d_backward_weight_workspace_bytes(int n, int cin, int h, int w, int cout, int ksize) {
  Geom g = make_geom(n, h, w, ksize == 3);
  const int ks = wgrad_pick_ksplit(g.total(), cout, cin, ksize * ksize, 148, g.P);
  Arena a(nullptr, 0);
  __half *xp, *yp; size_t xh, yh; float *partial, *colpart; pw::ScaleState sc; int* flag;
  return wgrad_op_layout(a, g, cin, cout, ksize * ksize, ks, &xp, &xh, &yp, &yh, &partial, &colpart, &sc, &flag) + 4096;
}

/* dw[cout][cin][k][k], db[cout] (either may be NULL) from x and dy = d(loss)/d(conv output). */
extern "C" int ssdn_conv2d_backward_weight(void* ws, size_t ws_bytes, const float* x, const float* dy, float* dw, float* db,
                                           int n, int cin, int h, int wd, int cout, int ksize, int blind, void* stream) {
  if (ksize != 1 && ksize != 3) return fail(-1, "ksize must be 1 or 3");
  cudaStream_t st = (cudaStream_t)stream;
  Geom g = make_geom(n, h, wd, ksize == 3);
  const int xp = round_up(cin, 8), yp = round_up(cout, 8), ntaps = ksize * ksize;
  const int ks = wgrad_pick_ksplit(g.total(), cout, cin, ntaps, 148, g.P);
  Arena a(ws, ws_bytes);
  __half *xpl, *ypl; size_t xh, yh; float *partial, *colpart; pw::ScaleState sc; int* flag;
  wgrad_op_layout(a, g, cin, cout, ntaps, ks, &xpl, &xh, &ypl, &yh, &partial, &colpart, &sc, &flag);
  if (!a.ok()) return fail(-3, "workspace too small: need %zu bytes, have %zu", a.off, ws_bytes);
  SSDN_CUDA(cudaMemsetAsync(xpl, 0, (2 * xh + 256) * sizeof(__half), st));
  SSDN_CUDA(cudaMemsetAsync(ypl, 0, (2 * yh + 256) * sizeof(__half), st));
  SSDN_CUDA(cudaMemsetAsync(flag, 0, 4, st));
  SSDN_CUDA(zero_scales(sc, st));
  const long long nx = (long long)n * cin * h * wd, ny = (long long)n * cout * h * wd;
  pw::leaf_scale_kernel<<<64, 256, 0, st>>>(x, nx, sc, 0, 0, 0);
  pw::leaf_scale_kernel<<<64, 256, 0, st>>>(dy, ny, sc, 1, 0, 0);
  ScaleRef xs{sc.k, nullptr}, ys{sc.k + 1, nullptr};
  pw::pack_nchw_kernel<<<pw::grid_for(nx), pw::kBlock, 0, st>>>(x, xpl, xpl + xh, n, cin, h, wd, g, xp, 0, 0, xs);
  pw::pack_nchw_kernel<<<pw::grid_for(ny), pw::kBlock, 0, st>>>(dy, ypl, ypl + yh, n, cout, h, wd, g, yp, 0, 0, ys);
  if (dw) {
    WgradPlan plan;
    int r = wgrad_plan_init(&plan, g.total(), ypl, ypl + yh, yp, 0, cout, xpl, xpl + xh, xp, 0, cin, make_taps(ksize, blind != 0, false, g.P), ks,
                            partial, flag, num_sms(), sc.k + 1, sc.k);
    if (r) return fail(r, "wgrad_plan_init failed (%d)", r);
    SSDN_CUDA(wgrad_launch(plan, st));
    wgradk::wgrad_reduce_launch(partial, ks, ntaps, cout, cin, plan.p.cin_pitch, dw, 0, st);
  }
  if (db) {
    const long long rows = g.total();
    pw::colsum_launch(ypl, ypl + yh, sc.k + 1, rows, yp, 0, cout, colpart, db, st);
  }
  int hflag = 0;
  SSDN_CUDA(cudaMemcpyAsync(&hflag, flag, 4, cudaMemcpyDeviceToHost, st));
  SSDN_CUDA(cudaStreamSynchronize(st));
  if (hflag) return fail(-4, "wgrad kernel pipeline timeout (role %d)", hflag);
  return 0;
}

// ------------------------------------------------------------------------------------ shifted max-pool (operator level)
// nn.Sequential(Shift2d((1, 0)), nn.MaxPool2d(2)) / nn.MaxPool2d(2) - models/noise_network.py:64-67 - through the network's own
// pool kernels: x [n][c][h][w] -> y [n][c][h/2][w/2], c a multiple of 8.  Synchronous.
extern "C" size_t ssdn_maxpool2_workspace_bytes(int n, int c, int h, int w) {
  Geom gs = make_geom(n, h, w, true), gd = make_geom(n, h / 2, w / 2, true);
  Arena a(nullptr, 0);
  a.take<__half>(2 * (((size_t)gs.total() * c + 127) / 128 * 128) + 256); a.take<__half>(2 * (((size_t)gs.total() * c + 127) / 128 * 128) + 256);
  a.take<__half>(2 * (((size_t)gd.total() * c + 127) / 128 * 128) + 256); a.take<float>((size_t)gd.total() * c);
  take_scales(a);
  return a.off + 4096;
}
// forward: y = maxpool(shift(x)).  backward (dy != NULL): also dz = d(loss)/dz for x = LeakyReLU(z) given dy = d(loss)/dy, i.e. the
// backward of [LeakyReLU -> (shift) -> max-pool] as the network runs it (first maximum wins ties, a winning padding zero
// swallows the gradient).  y or (dy, dz) may be NULL.
extern "C" int ssdn_maxpool2(void* ws, size_t ws_bytes, const float* x, float* y, const float* dy, float* dz, int n, int c, int h, int w, int blind,
                             void* stream) {
  if (c % 8 || h % 2 || w % 2 || n <= 0 || c <= 0) return fail(-1, "maxpool2 needs c %% 8 == 0 and even h, w");
  cudaStream_t st = (cudaStream_t)stream;
  Geom gs = make_geom(n, h, w, true), gd = make_geom(n, h / 2, w / 2, true);
  const size_t sh = ((size_t)gs.total() * c + 127) / 128 * 128, dh = ((size_t)gd.total() * c + 127) / 128 * 128;
  Arena a(ws, ws_bytes);
  __half* xs = a.take<__half>(2 * sh + 256); __half* zs = a.take<__half>(2 * sh + 256); __half* ys = a.take<__half>(2 * dh + 256);
  float* g1 = a.take<float>((size_t)gd.total() * c);
  pw::ScaleState sc = take_scales(a);
  if (!a.ok()) return fail(-3, "workspace too small: need %zu bytes, have %zu", a.off, ws_bytes);
  SSDN_CUDA(cudaMemsetAsync(xs, 0, (2 * sh + 256) * sizeof(__half), st));
  SSDN_CUDA(cudaMemsetAsync(zs, 0, (2 * sh + 256) * sizeof(__half), st));
  SSDN_CUDA(cudaMemsetAsync(ys, 0, (2 * dh + 256) * sizeof(__half), st));
  SSDN_CUDA(zero_scales(sc, st));
  const long long ne = (long long)n * c * h * w, no = ne / 4;
  pw::leaf_scale_kernel<<<64, 256, 0, st>>>(x, ne, sc, 0, 0, 0);
  ScaleRef xsc{sc.k, nullptr}, ysc{sc.k, nullptr}, zsc{sc.k + 1, nullptr};      // the pooled tensor shares x's scale (max <= max)
  pw::pack_nchw_kernel<<<pw::grid_for(ne), pw::kBlock, 0, st>>>(x, xs, xs + sh, n, c, h, w, gs, c, 0, 0, xsc);
  if (y) {
    pw::pool_fwd_kernel<<<pw::grid_for(no / 8), pw::kBlock, 0, st>>>(xs, xs + sh, gs, c, 0, xsc, ys, ys + dh, gd, c, 0, ysc, c, blind);
    pw::unpack_nchw_kernel<<<pw::grid_for(no), pw::kBlock, 0, st>>>(ys, ys + dh, sc.k, 0, y, n, c, h / 2, w / 2, gd, c, 0);
  }
  if (dy && dz) {
    pw::leaf_scale_kernel<<<64, 256, 0, st>>>(dy, no, sc, 1, 0, 0);
    pw::pack_nchw_f32_kernel<<<pw::grid_for(no), pw::kBlock, 0, st>>>(dy, g1, n, c, h / 2, w / 2, gd, c, 0);
    const int grid = (int)std::min<long long>(pw::kFusedColsumGrid, (no / 8 + pw::kFusedColsumBlock - 1) / pw::kFusedColsumBlock);
    pw::pool_bwd_kernel<<<grid, pw::kFusedColsumBlock, pw::kFusedColsumBlock * 8 * sizeof(float), st>>>(
        xs, xs + sh, gs, c, 0, g1, c, 0, nullptr, 0, 0, gd, zs, zs + sh, c, 0, zsc, c, blind, nullptr);
    pw::unpack_nchw_kernel<<<pw::grid_for(ne), pw::kBlock, 0, st>>>(zs, zs + sh, sc.k + 1, 0, dz, n, c, h, w, gs, c, 0);
  }
  SSDN_CUDA(cudaGetLastError());
  SSDN_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// ------------------------------------------------------------------------------------ measured tensor peak
#include "peak_kernel.cuh"
// Full-chip sustained tcgen05.mma rate (peak_kernel.cuh): kind f16 (f16 != 0) or tf32, the better of cta_group::1 / ::2,
// measured for `seconds` per configuration on `stream`'s device.  out4 = {TFLOP/s, FLOP/clk/SM, SM MHz (from clock64), seconds}.
extern "C" int ssdn_tensor_peak(int f16, double seconds, double* out4, void* stream) {
  if (!out4 || !(seconds > 0) || seconds > 30) return fail(-1, "bad arguments");
  peakk::PeakResult best{}; best.tflops = 0;
  for (int pair = 0; pair <= 1; ++pair) {
    peakk::PeakResult r{};
    int rc = peakk::measure(f16 ? 1 : 0, pair, num_sms(), seconds, (cudaStream_t)stream, &r);
    if (rc) return fail(-2, "tensor peak measurement failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError()));
    if (r.tflops > best.tflops) best = r;
  }
  out4[0] = best.tflops; out4[1] = best.flop_per_clk_sm; out4[2] = best.sm_mhz; out4[3] = best.seconds;
  return 0;
}
