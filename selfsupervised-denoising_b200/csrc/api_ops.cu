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
  size_t act_floats, slab_floats;
};

ConvOpLayout conv_layout(int n, int cin, int h, int w, int cout, int ksize, bool blind, bool dgrad) {
  ConvOpLayout L;
  L.g = make_geom(n, h, w, ksize == 3);
  L.cin_pitch = round_up(cin, 4);
  L.cout_padded = round_up(cout, 16);
  L.N = pick_n(L.cout_padded);
  L.taps = make_taps(ksize, blind, dgrad, L.g.P);
  L.act_floats = (size_t)L.g.total() * L.cin_pitch;
  L.slab_floats = conv_weight_slab_floats(cin, L.cout_padded, L.taps.n);
  return L;
}

// y[n,cout,h,w] = act( conv(x[n,cin,h,w], slab) + bias ); the slab orientation decides fwd vs dgrad
int run_conv(void* ws, size_t ws_bytes, const float* x, const float* w, const float* bias, float* y, int n, int cin,
             int h, int wd, int cout, int ksize, bool blind, bool dgrad, int lrelu_act, int w_cout, int w_cin,
             cudaStream_t st) {
  ConvOpLayout L = conv_layout(n, cin, h, wd, cout, ksize, blind, dgrad);
  Arena a(ws, ws_bytes);
  float* av = a.take<float>(L.act_floats); float* al = a.take<float>(L.act_floats);
  float* slab = a.take<float>(L.slab_floats);
  int* flag = a.take<int>(1);
  if (!ws) return (int)0;
  if (!a.ok()) return fail(-3, "workspace too small: need %zu bytes, have %zu", a.off, ws_bytes);
  SSDN_CUDA(cudaMemsetAsync(av, 0, L.act_floats * 4, st));
  SSDN_CUDA(cudaMemsetAsync(al, 0, L.act_floats * 4, st));
  SSDN_CUDA(cudaMemsetAsync(flag, 0, 4, st));
  const long long ne = (long long)n * cin * h * wd;
  pw::pack_nchw_kernel<<<pw::grid_for(ne), pw::kBlock, 0, st>>>(x, av, al, n, cin, h, wd, L.g, L.cin_pitch, 0, 0);
  const bool wide = conv_is_wide(cin, L.taps.n);
  int nc, kl; conv_chunks(cin, &nc, &kl, wide);
  const int n_tiles = L.cout_padded / L.N;
  const long long ns = (long long)L.slab_floats / 2;
  pw::weight_prep_kernel<<<pw::grid_for(ns), pw::kBlock, 0, st>>>(w, slab, w_cout, w_cin, L.taps.n, cout, cin, n_tiles, nc,
                                                                  L.N, dgrad ? 1 : 0, wide ? 32 : 16);
  ConvDst d{};
  d.v = y; d.lo = nullptr; d.cpitch = 0; d.coff = 0; d.g = L.g; d.map = MAP_NCHW;
  d.flags = (bias ? EP_BIAS : 0) | (lrelu_act ? EP_LRELU : 0); d.cvalid = cout; d.nimg = n; d.bias = bias;
  ConvPlan plan;
  int r = conv_plan_init(&plan, L.g, av, al, L.cin_pitch, 0, cin, slab, L.cout_padded, L.N, L.taps, d, flag, num_sms());
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
  a.take<float>(L.act_floats); a.take<float>(L.act_floats); a.take<float>(L.slab_floats);
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

extern "C" size_t ssdn_conv2d_backward_weight_workspace_bytes(int n, int cin, int h, int w, int cout, int ksize) {
  Geom g = make_geom(n, h, w, ksize == 3);
  const int ks = wgrad_pick_ksplit(g.total(), cout, cin, ksize * ksize, 148);
  Arena a(nullptr, 0);
  a.take<float>((size_t)g.total() * round_up(cin, 4) * 2); a.take<float>((size_t)g.total() * round_up(cout, 4) * 2);
  a.take<float>(wgrad_partial_floats(ks, ksize * ksize, cout, cin)); a.take<float>((size_t)1024 * round_up(cout, 4)); a.take<int>(1);
  return a.off + 4096;
}

/* dw[cout][cin][k][k], db[cout] (either may be NULL) from x and dy = d(loss)/d(conv output). */
extern "C" int ssdn_conv2d_backward_weight(void* ws, size_t ws_bytes, const float* x, const float* dy, float* dw, float* db,
                                           int n, int cin, int h, int wd, int cout, int ksize, int blind, void* stream) {
  if (ksize != 1 && ksize != 3) return fail(-1, "ksize must be 1 or 3");
  cudaStream_t st = (cudaStream_t)stream;
  Geom g = make_geom(n, h, wd, ksize == 3);
  const int xp = round_up(cin, 4), yp = round_up(cout, 4), ntaps = ksize * ksize;
  const int ks = wgrad_pick_ksplit(g.total(), cout, cin, ntaps, num_sms());
  Arena a(ws, ws_bytes);
  const size_t xf = (size_t)g.total() * xp, yf = (size_t)g.total() * yp;
  float* xv = a.take<float>(xf * 2); float* xl = xv + xf;
  float* yv = a.take<float>(yf * 2); float* yl = yv + yf;
  float* partial = a.take<float>(wgrad_partial_floats(ks, ntaps, cout, cin));
  float* colpart = a.take<float>((size_t)1024 * yp);
  int* flag = a.take<int>(1);
  if (!a.ok()) return fail(-3, "workspace too small: need %zu bytes, have %zu", a.off, ws_bytes);
  SSDN_CUDA(cudaMemsetAsync(xv, 0, xf * 8, st));
  SSDN_CUDA(cudaMemsetAsync(yv, 0, yf * 8, st));
  SSDN_CUDA(cudaMemsetAsync(flag, 0, 4, st));
  pw::pack_nchw_kernel<<<pw::grid_for((long long)n * cin * h * wd), pw::kBlock, 0, st>>>(x, xv, xl, n, cin, h, wd, g, xp, 0, 0);
  pw::pack_nchw_kernel<<<pw::grid_for((long long)n * cout * h * wd), pw::kBlock, 0, st>>>(dy, yv, yl, n, cout, h, wd, g, yp, 0, 0);
  if (dw) {
    WgradPlan plan;
    int r = wgrad_plan_init(&plan, g.total(), yv, yl, yp, 0, cout, xv, xl, xp, 0, cin, make_taps(ksize, blind != 0, false, g.P), ks,
                            partial, flag, num_sms());
    if (r) return fail(r, "wgrad_plan_init failed (%d)", r);
    SSDN_CUDA(wgrad_launch(plan, st));
    wgradk::wgrad_reduce_launch(partial, ks, ntaps, cout, cin, plan.p.cin_pitch, dw, 0, st);
  }
  if (db) {
    const long long rows = g.total();
    pw::colsum_launch(yv, yl, rows, yp, 0, cout, colpart, db, st);
  }
  int hflag = 0;
  SSDN_CUDA(cudaMemcpyAsync(&hflag, flag, 4, cudaMemcpyDeviceToHost, st));
  SSDN_CUDA(cudaStreamSynchronize(st));
  if (hflag) return fail(-4, "wgrad kernel pipeline timeout (role %d)", hflag);
  return 0;
}
