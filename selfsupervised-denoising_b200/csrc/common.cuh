// Shared definitions for the ssdn_b200 engine.
//
// DATA LAYOUT IN HBM ("padded-flat NHWC, two planes")
//   Every activation / activation-gradient tensor of the U-Net is stored channels-last as a 2-D
//   array [flat pixel][channel] of fp32.  Each image of height H and width W occupies
//   S = (H + 2) * (W + 1) consecutive flat pixels: two all-zero halo rows above the image and one
//   all-zero halo column after every row (pitch P = W + 1).  Pixel (b, y, x) lives at
//       flat = b * S + (y + 2) * P + x.
//   With this layout a 3x3 tap (dy, dx) of ANY of the convolutions on the path (half-plane
//   "shift" conv, plain conv, and both of their data-gradients) is the constant flat offset
//   dy * P + dx, for every image of the batch at once, and zero padding is simply what is stored
//   in the halo.  Kernels never write halo pixels, so they stay zero for the life of the buffer.
//   1x1 convolutions (the output head) use the same code with H+0 rows / pitch W ("dense" geometry).
//
//   Each tensor has two planes:  v  = the fp32 value,  lo = v - trunc_tf32(v).
//   The tensor cores truncate fp32 operands to tf32 (measured, profiles/r01_umma_probe_full.log),
//   so  a*b ~= v_a*v_b (hardware: hi*hi) + lo_a*v_b + v_a*lo_b  reproduces fp32-grade products
//   ("3xTF32", error ~6e-7) without ever materialising a separate "hi" plane.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SSDN_LRELU_SLOPE 0.1f

struct Geom {            // geometry of one padded-flat tensor
  int B, H, W;           // images, height, width
  int P, S;              // row pitch (W+1 or W), image stride in flat pixels
  int row0;              // halo rows above each image (2 or 0)
  __host__ __device__ long long total() const { return (long long)B * S; }
};

static inline Geom make_geom(int B, int H, int W, bool padded) {
  Geom g; g.B = B; g.H = H; g.W = W;
  g.P = padded ? W + 1 : W; g.row0 = padded ? 2 : 0; g.S = (H + g.row0) * g.P;
  return g;
}

__device__ __forceinline__ float tf32_lo(float v) {
  return v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
}
__device__ __forceinline__ float lrelu(float v) { return v > 0.f ? v : SSDN_LRELU_SLOPE * v; }

// Destination mapping modes of the conv epilogue
enum : int {
  MAP_IDENT = 0,      // same (b, y, x) in the destination geometry
  MAP_UP2 = 1,        // nearest 2x upsample: (b, y, x) -> the 2x2 block at (2y, 2x)
  MAP_UNROT = 2,      // Shift2d((1,0)) + un-rotate branch b / nimg into channel block (forward head input)
  MAP_UNROT_INV = 3,  // inverse of MAP_UNROT (data-gradient of the first head conv -> 4 branch images)
  MAP_NCHW = 4        // dense NCHW fp32 output (network output / unit tests)
};
enum : int {
  EP_BIAS = 1,        // add bias[c]
  EP_LRELU = 2,       // LeakyReLU(0.1)
  EP_ACT_GRAD = 4,    // multiply by LeakyReLU'(act) where act is the forward activation at the destination
  EP_WRITE_LO = 8,    // also write the lo plane
  EP_ACT_AT_SRC = 16  // EP_ACT_GRAD reads the activation at the SOURCE pixel / true GEMM channel
};

struct ConvDst {
  float* v; float* lo;           // destination planes
  int cpitch, coff;              // channels per destination pixel, channel offset of this conv's output
  Geom g;                        // destination geometry
  int map, flags;
  int cvalid;                    // number of real output channels (<= n_tiles * N)
  int nimg;                      // images per rotation group (MAP_UNROT / MAP_UNROT_INV)
  const float* bias;
  const float* act; int act_cpitch, act_coff;   // forward activation for EP_ACT_GRAD (destination geometry)
};
