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
//   Each GEMM operand tensor has two planes:  hi = round_tf32(v)  and  lo = round_tf32(v - hi)  (round to nearest).
//   hi + lo reproduces the fp32 value v to 2^-24 relative, and
//       a*b ~= hi_a*hi_b + lo_a*hi_b + hi_a*lo_b          (dropped lo*lo term <= 2^-24)
//   gives fp32-grade products from three tf32 tensor-core MMAs ("3xTF32"); both planes are exactly representable in
//   tf32, so the tensor core's own fp32->tf32 truncation (measured: profiles/r01_umma_probe_full.log) is a no-op.
//   Kernels that need the value itself (pooling, reductions) read hi + lo; sign tests read hi alone.
//   Tensors that are only consumed by pointwise kernels ("raw" gradients) have a single plain fp32 plane.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SSDN_LRELU_SLOPE 0.1f

// The tensor core's fp32 accumulator TRUNCATES (rounds toward zero): every tcgen05.mma accumulate step shrinks the magnitude
// of the running sum by about half an ulp, so a 3xTF32 result carries a BIAS of -1.5e-8 x (number of MMA instructions per
// output) relative to its magnitude - measured, stable to +-5 % across layer shapes and zero-mean data
// (tests/dev_conv_accuracy.py, profiles/r01_conv_accuracy.log): -4.9e-6 for a 96->96 3x3 layer, 1e-4 once compounded over
// the 20 layers.  The epilogues multiply by 1 + SSDN_ACC_BETA * n_mma: the remaining error is the zero-mean part (half the
// rms per layer, and it compounds as a square root instead of linearly).  SSDN_ACC_COMP=0 disables it (to re-measure).
#define SSDN_ACC_BETA 1.525e-8f

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

__device__ __forceinline__ float tf32_rn(float v) {
  uint32_t o;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(o) : "f"(v));
  return __uint_as_float(o);
}
__device__ __forceinline__ void tf32_split(float v, float& hi, float& lo) {
  hi = tf32_rn(v);
  lo = tf32_rn(v - hi);
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
  EP_ACT_GRAD = 4,    // multiply by LeakyReLU'(act): the sign bits of the forward activation come from mask_in
  EP_WRITE_LO = 8,    // destination is a GEMM operand: write the (hi, lo) tf32 split instead of the plain value
  EP_ACT_AT_SRC = 16  // EP_ACT_GRAD indexes mask_in by the SOURCE pixel / true GEMM channel (instead of the destination's)
};

// LeakyReLU sign masks: one bit per (pixel, channel) of an activation tensor, word w of a pixel holds channels
// [32w, 32w+32).  The forward epilogue that produces an activation writes them (mask_out); the data-gradient epilogue
// whose output is multiplied by LeakyReLU'(that activation) reads them (mask_in) - 1/32 of the bytes of re-reading the
// activation, fetched before the accumulator is ready.
struct ConvDst {
  float* v; float* lo;           // destination planes: (hi, lo) when EP_WRITE_LO, else v = plain fp32
  int cpitch, coff;              // channels per destination pixel, channel offset of this conv's output
  Geom g;                        // destination geometry
  int map, flags;
  int cvalid;                    // number of real output channels (<= n_tiles * N)
  int nimg;                      // images per rotation group (MAP_UNROT / MAP_UNROT_INV)
  const float* bias;
  uint32_t* mask_out; int mask_out_words;        // EP_LRELU: sign bits of the written activation, [destination pixel][words]
  const uint32_t* mask_in; int mask_in_words;    // EP_ACT_GRAD: sign bits of the forward activation, [pixel][words]
  float* colsum; int colsum_pitch;               // optional: per-(CTA, epilogue warp) column sums of the written values
                                                 // [gridDim.x * 8][colsum_pitch] (bias gradient of the consuming layer)
};
