/* ssdn_b200 — C ABI of the B200-native blind-spot denoising engine (libssdn_b200.so).
 *
 * The reference (COMP6248-Reproducability-Challenge/selfsupervised-denoising) has no FFI: its
 * boundary is the Python class API of ssdn/ssdn/models/noise_network.py and ssdn/ssdn/denoiser.py.
 * Each entry point below names the reference operator / method it replaces (file:line relative to
 * the reference repository).  Conventions:
 *   - all pointers are raw DEVICE pointers to float32 unless stated otherwise; boundary tensors are
 *     dense NCHW; `stream` is a cudaStream_t passed as void*;
 *   - functions return 0 on success, < 0 on error (-1 invalid argument, -2 CUDA error, -3 workspace
 *     too small, -4 device-side pipeline timeout, -5 not bound); ssdn_b200_last_error() returns a
 *     thread-local description.  No C++ exception crosses the boundary;
 *   - the library never allocates or frees caller memory: scratch comes from caller-owned
 *     workspaces sized by the *_workspace_bytes queries (PyTorch's caching allocator stays in charge);
 *   - one handle per (process, device), single caller thread; work is enqueued on the given stream.
 * There is no CPU implementation behind any of these functions.
 */
#ifndef SSDN_B200_H
#define SSDN_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* ssdn_b200_last_error(void);
int ssdn_b200_version(void);

/* ---- ShiftConv2d / nn.Conv2d (+ LeakyReLU 0.1) — models/noise_network.py:241-260, :70-156 ------------------
 * blind != 0: half-plane "shift" convolution (output row h sees input rows h-2..h for ksize 3);
 * blind == 0: ordinary 'same' convolution.  ksize in {1, 3}.  w is [cout][cin][k][k], bias [cout] or NULL.
 * These three operator-level calls are synchronous with respect to `stream`. */
size_t ssdn_conv2d_workspace_bytes(int n, int cin, int h, int w, int cout, int ksize);
int ssdn_conv2d_forward(void* ws, size_t ws_bytes, const float* x, const float* w, const float* bias, float* y,
                        int n, int cin, int h, int wd, int cout, int ksize, int blind, int lrelu_act, void* stream);
/* dx = d(loss)/dx given dy = d(loss)/d(conv output)  (autograd: convolution_backward, data gradient). */
int ssdn_conv2d_backward_data(void* ws, size_t ws_bytes, const float* dy, const float* w, float* dx, int n, int cin,
                              int h, int wd, int cout, int ksize, int blind, void* stream);
/* dw [cout][cin][k][k] and db [cout] (either may be NULL)  (autograd: convolution_backward, weight + bias gradient). */
size_t ssdn_conv2d_backward_weight_workspace_bytes(int n, int cin, int h, int w, int cout, int ksize);
int ssdn_conv2d_backward_weight(void* ws, size_t ws_bytes, const float* x, const float* dy, float* dw, float* db,
                                int n, int cin, int h, int wd, int cout, int ksize, int blind, void* stream);

/* ---- shifted max-pool — models/noise_network.py:64-67 --------------------------------------------------------
 * y = MaxPool2d(2)(Shift2d((1,0))(x)) (blind != 0) or MaxPool2d(2)(x): x [n][c][h][w] -> y [n][c][h/2][w/2], c % 8 == 0.
 * With dy / dz non-NULL also the backward of [LeakyReLU(0.1) -> (shift) -> max-pool] as the network runs it: x is the
 * ACTIVATION LeakyReLU(z); dz = d(loss)/dz given dy (first maximum wins ties as in ATen's max_pool2d; a winning padding
 * zero swallows the gradient).  Synchronous. */
size_t ssdn_maxpool2_workspace_bytes(int n, int c, int h, int w);
int ssdn_maxpool2(void* ws, size_t ws_bytes, const float* x, float* y, const float* dy, float* dz, int n, int c, int h, int w, int blind,
                  void* stream);

/* ---- index operators ------------------------------------------------------------------------------------
 * ssdn.utils.rotate x4 + torch.cat(dim=0) — models/noise_network.py:187-189, utils/data.py:42-67.
 *   y[r*n + b] = rotate(x[b], 90*r), x [n][c][h][w] (h == w), y [4n][c][h][w].  Bit exact. */
int ssdn_rot4_stack(const float* x, float* y, int n, int c, int h, int w, void* stream);
/* Shift2d((1,0)) + chunk(4) + rotate back (0, 270, 180, 90) + cat(dim=1) — models/noise_network.py:213-222.
 *   x [4n][c][h][w] -> y [n][4c][h][w].  Bit exact. */
int ssdn_shift_unrot_concat(const float* x, float* y, int n, int c, int h, int w, void* stream);

/* ---- NoiseNetwork — models/noise_network.py:48-226 ----------------------------------------------------------
 * A handle is an execution plan for one fixed (n, cin, cout, h, w, blindspot).  `params` / `grads` are flat fp32
 * buffers in nn.Module.parameters() order (for every conv: weight [cout][cin][k][k] then bias), see
 * ssdn_net_param_count.  h and w must be multiples of 32 (input_wh_mul, :228-238); blind-spot needs h == w.
 *   forward : NoiseNetwork.forward (:186-226); `training` != 0 also prepares what backward needs.
 *   backward: autograd of forward w.r.t. all parameters given dout = d(loss)/d(out); overwrites grads.
 *             Must follow the forward whose activations it differentiates (they live in the workspace).
 * Operand scales: the tensor-core kernels read every activation / gradient tensor as two fp16 planes scaled by a per-tensor
 * power of two that is derived from the tensor's maximum in the PREVIOUS pass (weights and the loss gradient: from the
 * current one).  A pass during which some maximum left the accurate band is "stale": its results must not be used and the
 * pass must simply be repeated (the first passes of a new plan usually are).  backward() writes the step's verdict to
 * `stale_out` (DEVICE float, may be NULL; 1.0f = forward or backward stale) so that the caller can append it to the
 * gradient buffer it all-reduces and hand it to ssdn_adam_step as a skip flag; ssdn_net_scale_status synchronises and
 * returns {forward stale, backward stale, stale passes since bind}. */
int ssdn_net_create(int n, int cin, int cout, int h, int w, int blindspot, void** handle);
void ssdn_net_destroy(void* handle);
size_t ssdn_net_workspace_bytes(void* handle);
size_t ssdn_net_param_count(void* handle);
int ssdn_net_bind(void* handle, void* workspace, size_t workspace_bytes, void* stream);
int ssdn_net_forward(void* handle, const float* params, const float* x, float* out, int training, void* stream);
int ssdn_net_backward(void* handle, const float* params, const float* dout, float* grads, float* stale_out, void* stream);
int ssdn_net_scale_status(void* handle, int* status3, void* stream);
/* Synchronises `stream` and reports device-side errors (a bounded mbarrier wait that expired). */
int ssdn_net_check(void* handle, void* stream);
/* Number of kernels one forward (+ backward when training != 0) launches. */
int ssdn_net_kernel_launches(void* handle, int training);
/* Test hooks: copy the first c channels of a named internal buffer (plane 0 = value, 1 / 2 = the fp16 lo / hi plane as
 * stored) to / from a dense [B][c][H][W] tensor; with out == NULL only dims[4] = {B, c, H, W} is filled.
 * ssdn_net_debug_scales: the 96 scale exponents in use and the bit patterns of the running maxima. */
int ssdn_net_debug_scales(void* handle, int* k96, unsigned* amax96, void* stream);
int ssdn_net_debug_read(void* handle, const char* name, int plane, int c, float* out, int* dims, void* stream);
int ssdn_net_debug_write(void* handle, const char* name, int c, const float* src, void* stream);
/* Per-launch CUDA-event timing of every engine kernel (bench.py roofline).  Kinds, in order: conv_fwd, conv_dgrad, wgrad,
 * wgrad_reduce, pool_fwd, pool_bwd, up_bwd, pack, weight_prep, bias, scale, posterior_fwd, posterior_bwd, adam, other
 * (ssdn_profile_kinds() = 15).  profile_end: out[kind*4 + {0,1,2,3}] = {launches, ms, algorithmic FLOPs, algorithmic HBM bytes}.
 * profile_records: per launch of the last profiled region, in launch order, out[4*i + {0,1,2,3}] = {kind, ms, FLOPs, bytes};
 * returns the count. */
int ssdn_profile_begin(void);
int ssdn_profile_kinds(void);
int ssdn_profile_end(double* out);
int ssdn_profile_records(double* out, int max_records);

/* ---- Denoiser._ssdn_pipeline maths — denoiser.py:222-397 ------------------------------------------------------
 * net_out [n][c + c(c+1)/2][h][w] (mean, then the triangular factor of Sigma_x), noisy [n][c][h][w], c in {1, 3};
 * sigma_raw [n][cs], cs in {1, c}.
 * Gaussian noise (poisson == 0): sigma_known != 0 -> sigma = max(raw, 1e-3) (:279-282), else
 *   sigma = softplus(raw - 4) + 1e-3 (:272-275) and the -0.1 sigma regulariser is applied (:333, :360-363).
 * noise_model: bit 0 = Poisson noise, bit 1 = diagonal covariance (cfg DIAGONAL_COVARIANCE, :213, :236-243: net_out then has
 *   c + c channels - the mean and c diagonal factors d, Sigma_x = diag(d^2); the reference's own branch raises a TypeError at
 *   :240 (`c00.shape()`), this is what it evidently means - see oracle/ssdn_oracle.py:ssdn_posterior).
 * Poisson noise (noise_model & 1, :285-297): sigma = sqrt(max(mu, 1e-3) * k) per pixel and channel with k = 1 / raw
 *   (raw = the known lambda) or k = softplus(raw - 4) + 1e-3 (learned, regularised); the loss gradient then also
 *   reaches mu through sigma.
 * Outputs: pme [n][c][h][w] posterior mean (:328-330, :366-372), loss [n] per-sample mean NLL (:323-363, :388),
 * model_std [n][h][w] (:331, :375-377), noise_std (:332, :378-380): [n] for Gaussian noise (may be NULL), [n][h][w] for
 * Poisson noise. */
size_t ssdn_loss_workspace_bytes(int n, int c);
int ssdn_posterior_forward(void* ws, const float* net_out, const float* noisy, const float* sigma_raw, int n, int c, int h, int w,
                           int cs, int sigma_known, int noise_model, float* pme, float* loss, float* model_std, float* noise_std,
                           void* stream);
/* Gradient of sum_n gloss[n] * loss[n]: dnet like net_out; dsigma_raw [n][cs] (NULL or sigma_known: not computed). */
int ssdn_posterior_backward(void* ws, const float* net_out, const float* noisy, const float* sigma_raw, const float* gloss, int n,
                            int c, int h, int w, int cs, int sigma_known, int noise_model, float* dnet, float* dsigma_raw, void* stream);
/* torch.mean(x, dim=(2,3)) of the sigma-estimator output — denoiser.py:263-265.  x [rows][hw] -> out [rows]. */
int ssdn_spatial_mean_forward(const float* x, int rows, int hw, float* out, void* stream);
int ssdn_spatial_mean_backward(const float* g, int rows, int hw, float* dx, void* stream);

/* ---- Denoiser._mse_pipeline loss — denoiser.py:153-154: loss[n] = mean over chw of (a - b)^2 -------------------- */
int ssdn_mse_forward(void* ws, const float* a, const float* b, int n, int chw, float* loss, void* stream);
int ssdn_mse_backward(const float* a, const float* b, const float* gloss, int n, int chw, float* da, void* stream);
/* ---- loss_mask_mse — utils/n2v_loss.py:6-17 + denoiser.py:176-178.  coords: DEVICE int64 [k][2] (the FIRST sample's
 * coordinate list, used for the whole batch, indexing [:, :, c0, c1]); loss[n] = mean_c sum_k (ref - out)^2. */
int ssdn_masked_mse_forward(void* ws, const float* out, const float* ref, const long long* coords, int k, int n, int c, int h, int w,
                            float* loss, void* stream);
int ssdn_masked_mse_backward(const float* out, const float* ref, const long long* coords, int k, const float* gloss, int n, int c,
                             int h, int w, float* dout, void* stream);

/* ---- optim.Adam(betas=[0.9, 0.99]) over a flat buffer — train.py:100-107, :202 ------------------------------------
 * In place on p/m/v; `step` is the 1-based step count; the gradient is multiplied by grad_scale first (1/world_size
 * after the data-parallel all-reduce).  skip: n_skip (<= 8) DEVICE floats or NULL - if any is non-zero the update is a no-op
 * (the stale flags of ssdn_net_backward, summed over ranks by the same all-reduce as the gradients). */
int ssdn_adam_step(float* p, const float* g, float* m, float* v, long long count, double lr, double beta1, double beta2, double eps,
                   long long step, double grad_scale, const float* skip, int n_skip, void* stream);

/* The same update with the step-dependent scalars in DEVICE memory: hyper6 = {lr / (1 - beta1^step), beta1, beta2, eps,
 * sqrt(1 - beta2^step), grad_scale}.  A CUDA graph holding this launch can be replayed for every step (ssdn.train.GraphedTrainStep). */
int ssdn_adam_step_dev(float* p, const float* g, float* m, float* v, long long count, const float* hyper6, const float* skip, int n_skip,
                       void* stream);

/* ---- measured tensor roofline (bench.py) ------------------------------------------------------------------------------
 * Full-chip sustained tcgen05.mma rate, kind::f16 (f16 != 0) or kind::tf32: every SM issues back-to-back M = 128 / 256,
 * N = 256 MMAs on shared-memory-resident operands for `seconds` per configuration.  Synchronous.
 * out4 = {TFLOP/s, FLOP/clk/SM, SM clock in MHz seen by clock64, seconds timed}. */
int ssdn_tensor_peak(int f16, double seconds, double* out4, void* stream);

/* ---- on-GPU input pipeline — train.py:756-760 (RandomCrop), datasets/noise_wrapper.py:98-163, utils/noise.py:14-63 ---------
 * images: DEVICE uint8 cache [n_images][c][h][w] (c <= 4).  Output sample i takes image order[i] (DEVICE int32 [n]) or, with
 * order == NULL, image (step * n + i) mod n_images; a uniformly random patch x patch crop; clean = u8 / 255 (ToTensor);
 * noisy = clean + N(0, sigma^2) with sigma = sigma_lo, or U(sigma_lo, sigma_hi) per sample AND channel when
 * sigma_hi > sigma_lo (the range styles draw per leading axis of a CHW image: noise.py:55-56); clipped to [0, 1] if clip.
 * Randomness is Philox4x32-10 keyed by seed and indexed by (step, stream_id, sample, pixel): reproducible; stream_id
 * selects an independent noise realisation of the SAME crops (the Noise2Noise reference).  Outputs dense NCHW fp32:
 * clean (may be NULL), noisy [n][c][patch][patch], sigma [n][c] (may be NULL).  Statistical parity with the CPU generator. */
int ssdn_noisy_crops(const unsigned char* images, int n_images, int c, int h, int w, const int* order, int n, int patch,
                     unsigned long long seed, unsigned long long step, int stream_id, float sigma_lo, float sigma_hi, int clip,
                     float* clean, float* noisy, float* sigma, void* stream);

/* The same crops with the reference's Poisson styles ('poisson30', 'poisson5_50') — utils/noise.py:66-109 (add_poisson):
 * noisy = (clean * lam + K) / lam with K ~ Poisson(1) per element (the reference's generator has the constant rate 1, kept),
 * lam = lam_lo, or U(lam_lo, lam_hi) per sample AND channel when lam_hi > lam_lo; clipped to [0, 1] if clip.  lam [n][c]
 * (may be NULL) is what the data loader reports as INPUT_NOISE_VALUES.  Same crop / image draws as ssdn_noisy_crops at equal
 * (seed, step); other arguments as there. */
int ssdn_poisson_crops(const unsigned char* images, int n_images, int c, int h, int w, const int* order, int n, int patch,
                       unsigned long long seed, unsigned long long step, int stream_id, float lam_lo, float lam_hi, int clip,
                       float* clean, float* noisy, float* lam, void* stream);

/* Noise2Void "uniform pixel selection" — utils/n2v_ups.py:7-96 (manipulate, get_stratified_coords).  noisy / masked
 * [n][c][h][w] DEVICE fp32 (h, w multiples of 8; masked may alias noisy only if a stale read of a masked pixel is acceptable -
 * pass a separate buffer), coords DEVICE int64 [n][(h/8)*(w/8)][2] = (x, y) per 8x8 box in the reference's list order.
 * The reference's index quirks (min for max, negative wrap) are kept; parity is statistical. */
int ssdn_n2v_mask(const float* noisy, float* masked, long long* coords, int n, int c, int h, int w, int subpatch_size,
                  unsigned long long seed, unsigned long long step, void* stream);

#ifdef __cplusplus
}
#endif
#endif
