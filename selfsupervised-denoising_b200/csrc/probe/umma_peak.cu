// Hardware probe #4 (developer tool, also the source of bench.py's tensor-roofline denominator): FULL-CHIP sustained
// tcgen05.mma rate.  Every SM runs one CTA whose elected thread issues back-to-back M = 128 (cta_group::1) or M = 256
// (cta_group::2, clusters of two) MMAs of N = 256 on shared-memory-resident operands into two alternating TMEM
// accumulators - no TMA, no epilogue: what the tensor pipe of the whole chip sustains at the clocks it runs at.
// Measured per kind (f16 = fp16 operands, K = 16; tf32, K = 8) with CUDA events around launches of >= 5 ms.
//   usage: umma_peak [seconds per configuration, default 1.0] [out.json]
// Output: one line per configuration and, if a path is given, a JSON file
//   {"f16_tflops": .., "tf32_tflops": .., "f16_flop_per_clk_sm": .., "tf32_flop_per_clk_sm": .., "sm_clock_mhz_est": ..}
// (the best of cta_group 1 / 2 per kind).  bench.py runs this binary's library twin (ssdn_tensor_peak, api_ops.cu) live.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../umma.cuh"
#include "../peak_kernel.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2);} } while (0)

int main(int argc, char** argv) {
  const double secs = argc > 1 ? atof(argv[1]) : 1.0;
  CK(cudaSetDevice(0));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);
  double best[2] = {0, 0}, best_fpc[2] = {0, 0}, clk_mhz = 0;
  for (int f16 = 1; f16 >= 0; --f16)
    for (int pair = 0; pair <= 1; ++pair) {
      peakk::PeakResult r;
      int rc = peakk::measure(f16, pair, sms, secs, 0, &r);
      if (rc) { printf("kind %s cta_group::%d: error %d\n", f16 ? "f16" : "tf32", pair + 1, rc); continue; }
      printf("kind::%-4s cta_group::%d  M=%d N=256: %8.1f TFLOP/s over %.2f s (%d launches), %.0f FLOP/clk/SM (dense %d), SM clock from clock64 %.0f MHz\n",
             f16 ? "f16" : "tf32", pair + 1, pair ? 256 : 128, r.tflops, r.seconds, r.launches, r.flop_per_clk_sm, f16 ? 8192 : 4096, r.sm_mhz);
      if (r.tflops > best[f16]) { best[f16] = r.tflops; best_fpc[f16] = r.flop_per_clk_sm; clk_mhz = r.sm_mhz; }
    }
  if (argc > 2) {
    FILE* f = fopen(argv[2], "w");
    if (f) {
      fprintf(f, "{\"f16_tflops\": %.1f, \"tf32_tflops\": %.1f, \"f16_flop_per_clk_sm\": %.0f, \"tf32_flop_per_clk_sm\": %.0f, \"sm_clock_mhz_est\": %.0f, \"sms\": %d, "
                 "\"how\": \"csrc/probe/umma_peak.cu: all SMs issue back-to-back tcgen05.mma M=128/256 N=256 on resident smem operands, CUDA events, %.1f s per configuration\"}\n",
              best[1], best[0], best_fpc[1], best_fpc[0], clk_mhz, sms, secs);
      fclose(f);
    }
  }
  return 0;
}
