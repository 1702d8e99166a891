// Host-side helpers shared by the C-ABI entry points: error reporting, workspace carving,
// conv tap tables, and thin launch wrappers around the kernels.
#pragma once
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "conv_igemm.cuh"
#include "pointwise.cuh"
#include "wgrad_igemm.cuh"   // uses the LaunchProfiler declared in conv_igemm.cuh

namespace eng {

inline char* err_buf() { static thread_local char buf[512] = {0}; return buf; }
inline int fail(int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(err_buf(), 512, fmt, ap); va_end(ap);
  return code;
}
#define SSDN_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return eng::fail(-2, "CUDA error '%s' at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

inline int num_sms() {      // of the CURRENT device (a process may drive plans on several devices)
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

struct Arena {           // bump allocator over a caller-owned workspace (256-byte aligned blocks)
  char* base; size_t cap, off;
  Arena(void* b, size_t c) : base((char*)b), cap(c), off(0) {}
  template <class T> T* take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) / 256 * 256;
    T* p = base ? (T*)(base + off) : nullptr;
    off += bytes;
    return p;
  }
  bool ok() const { return !base || off <= cap; }
};

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// 3x3 / 1x1 stencil offsets in weight order tap = kh*k + kw.
//   blind: rows (h-2, h-1, h) [ShiftConv2d, models/noise_network.py:241-260]; else rows (h-1, h, h+1).
//   dgrad: the data-gradient reads the output gradient at the mirrored offsets.
inline ConvTaps make_taps(int ksize, bool blind, bool dgrad, int P) {
  ConvTaps t; t.n = ksize * ksize;
  const int sh = ksize == 1 ? 0 : (blind ? 2 : 1), sw = ksize / 2;
  for (int kh = 0; kh < ksize; ++kh)
    for (int kw = 0; kw < ksize; ++kw) {
      int off = (kh - sh) * P + (kw - sw);
      t.off[kh * ksize + kw] = dgrad ? -off : off;
    }
  return t;
}

inline int pick_n(int cout_padded) {
  if (cout_padded <= 128) return cout_padded;
  for (int n = 128; n >= 16; n -= 16) if (cout_padded % n == 0) return n;
  return 16;
}

}  // namespace eng
