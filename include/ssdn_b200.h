/* ssdn_b200 — C ABI of the B200-native blind-spot denoising engine.
 *
 * The reference (COMP6248-Reproducability-Challenge/selfsupervised-denoising) has no FFI: its
 * boundary is the Python class API of ssdn/ssdn/models/noise_network.py and ssdn/ssdn/denoiser.py.
 * Each entry point below names the reference operator / method it replaces.  All pointers are raw
 * DEVICE pointers to float32 unless stated; tensors at the boundary are dense NCHW; `stream` is a
 * cudaStream_t.  Functions return 0 on success and a negative code on error, in which case
 * ssdn_b200_last_error() returns a thread-local description.  The library never allocates or frees
 * caller memory: scratch comes from caller-owned workspaces sized by the *_workspace_bytes queries.
 */
#ifndef SSDN_B200_H
#define SSDN_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* ssdn_b200_last_error(void);
int ssdn_b200_version(void);

/* ---- ShiftConv2d / nn.Conv2d (+ LeakyReLU) — models/noise_network.py:241-260, :70-156 ----------
 * blind != 0: half-plane "shift" convolution (output row h sees input rows h-2..h for ksize 3);
 * blind == 0: ordinary 'same' convolution.  ksize in {1, 3}.  w is [cout][cin][k][k], bias [cout] or NULL.
 * Synchronous with respect to `stream`. */
size_t ssdn_conv2d_workspace_bytes(int n, int cin, int h, int w, int cout, int ksize);
int ssdn_conv2d_forward(void* ws, size_t ws_bytes, const float* x, const float* w, const float* bias, float* y,
                        int n, int cin, int h, int wd, int cout, int ksize, int blind, int lrelu_act, void* stream);
/* dx = d(loss)/dx given dy = d(loss)/d(conv output) (autograd of the op above, convolution_backward dgrad). */
int ssdn_conv2d_backward_data(void* ws, size_t ws_bytes, const float* dy, const float* w, float* dx, int n, int cin,
                              int h, int wd, int cout, int ksize, int blind, void* stream);

/* dw [cout][cin][k][k] and db [cout] (either may be NULL): convolution_backward wgrad + bias reduction. */
size_t ssdn_conv2d_backward_weight_workspace_bytes(int n, int cin, int h, int w, int cout, int ksize);
int ssdn_conv2d_backward_weight(void* ws, size_t ws_bytes, const float* x, const float* dy, float* dw, float* db,
                                int n, int cin, int h, int wd, int cout, int ksize, int blind, void* stream);

#ifdef __cplusplus
}
#endif
#endif
