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
//   Each GEMM operand tensor has two fp16 planes of the value scaled by a per-tensor power of two 2^k:
//       hi = rn_f16(v * 2^k),  lo = rn_f16(v * 2^k - hi)        (round to nearest, saturating)
//   hi + lo reproduces v * 2^k to 2^-23 relative as long as |v| * 2^k stays inside fp16's range with room below it for
//   the lo term (profiles/r01_split_numerics.txt, profiles/r02_f16_probe.log), and
//       a*b ~= hi_a*hi_b + lo_a*hi_b + hi_a*lo_b          (dropped lo*lo term <= 2^-22)
//   gives fp32-grade products from three kind::f16 tensor-core MMAs of K = 16 with fp32 accumulation in TMEM - half the
//   instructions and half the operand bytes of the 3xTF32 scheme this replaces.  The consumer's epilogue multiplies the
//   accumulator by 2^-(k_a + k_b).
//   SCALES (ScaleRef below): every operand tensor owns a slot {k, k_next, amax} in device memory.  Whoever writes the planes
//   reads k and folds max|v| of what it wrote into amax (one atomicMax per warp at the end of the kernel).  Leaves (weights,
//   the loss gradient) get their exact scale from a reduction that runs before they are packed; every other tensor uses
//   the scale derived from its maximum in the PREVIOUS pass ("delayed scaling", net.cuh: scale_finish / scale_begin), aimed
//   at 2^11: five binades of headroom before fp16 saturates and thirteen before accuracy is lost.  A pass whose maxima
//   left that band reports itself stale; the host re-runs it (first use of a plan) and the optimiser step is predicated on
//   the flag, so a stale gradient never reaches the weights.
//   Kernels that need the value itself (pooling, reductions) read (hi + lo) * 2^-k; sign tests read hi alone.
//   Tensors that are only consumed by pointwise kernels ("raw" gradients) have a single plain fp32 plane.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define SSDN_LRELU_SLOPE 0.1f

// Programmatic dependent launch: every engine kernel is launched with programmaticStreamSerializationAllowed, so that it may
// become resident (and run its prologue: barrier init, TMEM allocation, tensor-map prefetch) while the kernel before it on
// the stream is still running on other SMs; pdl_wait() then blocks until that kernel has completed and its writes are visible.
// It is the FIRST thing a kernel does before touching global memory, so stream-order semantics are unchanged.  On the small
// pyramid levels (grids of 12 .. 90 CTAs, 15 - 25 us per kernel) this hides most of the launch latency and the prologue.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// The tensor core's fp32 accumulator TRUNCATES (rounds toward zero): every tcgen05.mma accumulate step shrinks the magnitude
// of the running sum by a fraction of an ulp, so a split-operand result carries a BIAS proportional to the number of MMA
// instructions per output.  Its size depends on how the partial sums evolve, i.e. on the data: measured per MMA instruction
// (kind::f16, tests/dev_conv_accuracy.py and tests/dev_trained_bias.py -> profiles/r02_conv_accuracy*.log, r02_trained_bias*.log)
//   -1.4e-8 .. -1.6e-8  zero-mean synthetic activations and weights, every layer shape
//   -0.35e-8 .. -1.3e-8 the 20 layers of the TRAINED checkpoint on their own activations (smallest on the 144-channel concat inputs)
//   -3.8e-8             all-positive activations AND weights (no layer of the network looks like that).
// The epilogues multiply by 1 + SSDN_ACC_BETA * n_mma with the midpoint of the first two ranges: the residual is then at most
// 0.65e-8 per MMA either way - below 1.6e-6 for the deepest accumulation of the network (243 MMAs) and below 2e-5 if the
// residuals of all 20 layers lined up - where the uncompensated bias is 4e-5 summed over the depth of the network.
// tests/test_gpu_accuracy.py pins both numbers on the trained checkpoint.  SSDN_ACC_COMP=0 disables it (to re-measure).
#define SSDN_ACC_BETA 0.95e-8f

struct Geom {            // geometry of one padded-flat tensor
  int B, H, W;           // images, height, width
  int P, S;              // row pitch (W+1 or W), image stride in flat pixels
  int row0;              // halo rows above each image (2 or 0)
  __host__ __device__ long long total() const { return (long long)B * S; }
};

#include <cstdlib>
#include <utility>
static inline bool pdl_enabled() { static const bool on = !(getenv("SSDN_PDL") && atoi(getenv("SSDN_PDL")) == 0); return on; }
// Launch priority of the kernels on the CRITICAL path (everything except the weight gradients, which run on the side stream and
// only have to be done by the end of the backward pass): when CTAs of both are pending, the critical path's are placed first and
// the weight gradients fill the SMs the small pyramid levels leave idle.  A launch attribute (captured into CUDA graphs), relative
// to the device's range (0 = least urgent).  An experiment knob (SSDN_MAIN_PRIORITY=-2), see below.
// MEASURED (tools/sweep_bench.sh, interleaved runs): priority -2 / -5 on the critical path makes the step 5 % SLOWER (3.92 vs 3.71 ms) -
// the weight gradients are then pushed to the end of the backward pass, where nothing overlaps them - so the default is 0 (off).
static inline int main_priority() { static const int p = getenv("SSDN_MAIN_PRIORITY") ? atoi(getenv("SSDN_MAIN_PRIORITY")) : 0; return p; }
inline bool& launching_background() { static thread_local bool b = false; return b; }   // set around the side-stream launches
static inline int launch_attrs(cudaLaunchAttribute* a, bool cluster2) {                   // fills a[0..], returns the count
  int n = 0;
  if (cluster2) { a[n].id = cudaLaunchAttributeClusterDimension; a[n].val.clusterDim.x = 2; a[n].val.clusterDim.y = 1; a[n].val.clusterDim.z = 1; ++n; }
  if (pdl_enabled()) { a[n].id = cudaLaunchAttributeProgrammaticStreamSerialization; a[n].val.programmaticStreamSerializationAllowed = 1; ++n; }
  if (main_priority() != 0 && !launching_background()) { a[n].id = cudaLaunchAttributePriority; a[n].val.priority = main_priority(); ++n; }
  return n;
}
// kernel<<<grid, block, smem, stream>>>(args...) with the programmatic-stream-serialization attribute (see pdl_wait)
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute a[3];
  cfg.attrs = a; cfg.numAttrs = launch_attrs(a, false);
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

static inline Geom make_geom(int B, int H, int W, bool padded) {
  Geom g; g.B = B; g.H = H; g.W = W;
  g.P = padded ? W + 1 : W; g.row0 = padded ? 2 : 0; g.S = (H + g.row0) * g.P;
  return g;
}

// ---------------------------------------------------------------- operand scales
constexpr int kScaleTarget = 11;      // a tensor's maximum is aimed at [2^11, 2^12) in its fp16 planes
constexpr int kScaleHiLimit = 16;     // floor(log2(max)) + k >= 16: values beyond fp16's largest finite number were saturated
constexpr int kScaleLoLimit = 2;      // floor(log2(max)) + k <  2: the lo plane has lost bits (graceful, but re-run)
struct ScaleRef {                     // one operand tensor's scale slot (device pointers; null = unscaled test tensor)
  const int* k;                       // exponent in use: planes hold v * 2^k
  unsigned* amax;                     // bit pattern of max|v| written during the current pass (atomicMax target)
};
__device__ __forceinline__ float exp2_int(int k) {          // 2^k as a float, k clamped to the normal range
  k = k < -126 ? -126 : (k > 127 ? 127 : k);
  return __int_as_float((127 + k) << 23);
}
__device__ __forceinline__ int floor_log2_bits(unsigned bits) { return (int)(bits >> 23) - 127; }   // of a positive normal float
// scale exponent that puts a maximum with bit pattern `bits` into [2^target, 2^(target+1)); 0 when the tensor is all zero
__device__ __forceinline__ int scale_for_amax(unsigned bits) { return bits ? kScaleTarget - floor_log2_bits(bits) : 0; }
__device__ __forceinline__ void amax_commit(unsigned* amax, float m) {   // whole warp: one atomic per warp
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && amax && m > 0.f) atomicMax(amax, __float_as_uint(m));
}

// ---------------------------------------------------------------- fp16 two-term split
// packs (a -> low half, b -> high half), round to nearest, saturating to +-65504 instead of infinity
__device__ __forceinline__ uint32_t cvt_f16x2_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float2 f16x2_to_float2(uint32_t r) { return __half22float2(*reinterpret_cast<const __half2*>(&r)); }
// (a, b) already multiplied by the tensor's scale -> packed hi pair and packed lo pair
__device__ __forceinline__ void f16_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = cvt_f16x2_sat(a, b);
  const float2 h = f16x2_to_float2(hi);
  lo = cvt_f16x2_sat(a - h.x, b - h.y);
}
__device__ __forceinline__ void f16_split1(float a, __half& hi, __half& lo) {
  uint32_t h, l; f16_split2(a, 0.f, h, l);
  hi = __ushort_as_half((unsigned short)(h & 0xffffu)); lo = __ushort_as_half((unsigned short)(l & 0xffffu));
}
// 8 consecutive channels: scaled floats -> one 16-byte vector per plane
__device__ __forceinline__ void f16_split8(const float (&f)[8], uint4& hi, uint4& lo) {
  f16_split2(f[0], f[1], hi.x, lo.x); f16_split2(f[2], f[3], hi.y, lo.y);
  f16_split2(f[4], f[5], hi.z, lo.z); f16_split2(f[6], f[7], hi.w, lo.w);
}
// 8 consecutive channels of both planes -> hi + lo as floats (still carrying the tensor's scale)
__device__ __forceinline__ void f16_join8(const uint4& hi, const uint4& lo, float (&f)[8]) {
  const uint32_t* h = reinterpret_cast<const uint32_t*>(&hi); const uint32_t* l = reinterpret_cast<const uint32_t*>(&lo);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 a = f16x2_to_float2(h[i]), b = f16x2_to_float2(l[i]);
    f[2 * i] = a.x + b.x; f[2 * i + 1] = a.y + b.y;
  }
}
__device__ __forceinline__ float f16_join1(__half hi, __half lo) { return __half2float(hi) + __half2float(lo); }
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
  EP_WRITE_LO = 8,    // destination is a GEMM operand: write the scaled (hi, lo) fp16 split instead of the plain value
  EP_ACT_AT_SRC = 16  // EP_ACT_GRAD indexes mask_in by the SOURCE pixel / true GEMM channel (instead of the destination's)
};

// LeakyReLU sign masks: one bit per (pixel, channel) of an activation tensor, word w of a pixel holds channels
// [32w, 32w+32).  The forward epilogue that produces an activation writes them (mask_out); the data-gradient epilogue
// whose output is multiplied by LeakyReLU'(that activation) reads them (mask_in) - 1/32 of the bytes of re-reading the
// activation, fetched before the accumulator is ready.
struct ConvDst {
  float* v;                      // plain fp32 destination (raw gradients, NCHW output) when not EP_WRITE_LO
  __half* hi; __half* lo;        // EP_WRITE_LO: the two fp16 planes of a GEMM operand, values scaled by 2^(*scale.k)
  ScaleRef scale;                //    scale slot of the destination tensor
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
