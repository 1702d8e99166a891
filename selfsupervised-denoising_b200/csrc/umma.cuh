// sm_100a primitives used by the implicit-GEMM kernels: mbarrier, TMA (cp.async.bulk.tensor),
// TMEM allocation, tcgen05.mma (kind::f16; kind::tf32 kept for the probes), tcgen05.ld, and the shared-memory / instruction
// descriptor encoders. Everything here is inline PTX; no CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

// developer instrumentation: accumulate the clocks a call (an mbarrier wait) takes
#define SSDN_TIMED(acc, call) do { const long long t0_ = clock64(); call; acc += clock64() - t0_; } while (0)

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of a fully converged warp.  Role loops run WARP-UNIFORMLY and only the single-thread instructions
// (tcgen05.mma / commit, TMA, expect_tx) are predicated with this: operands then stay in uniform registers.
// Running the whole loop under `if (lane == 0)` instead makes ptxas move every descriptor through R2UR inside an
// elect "waterfall" loop - measured 184 clk per MMA instead of <= 48 (profiles/r01_umma_rate.log).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint (ns): the warp may sleep in hardware up to that long instead of spinning through the
// issue slots of its scheduler (which the epilogue warps need)
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
// Bounded spin: a kernel bug must not hang the GPU box.  On expiry the CTA-wide abort word (shared memory) is set;
// every later wait of the CTA then returns at once, so the kernel drains quickly (its results are garbage and the
// host sees the error flag).  The return value is deliberately NOT used for control flow by the callers: loop
// structure stays warp-uniform, which is what lets ptxas keep MMA / TMA operands in uniform registers.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t abort_addr, int* error_flag, int code, uint32_t hint_ns = 0) {
  uint32_t aborted;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(aborted) : "r"(abort_addr) : "memory");
  if (aborted) return;
#ifdef UMMA_UNBOUNDED_WAIT
  while (!mbar_try_wait(bar, parity)) {}
#else
  if (hint_ns) {
#pragma unroll 1
    for (uint32_t i = 0; i < (1u << 22); ++i)
      if (mbar_try_wait_hint(bar, parity, hint_ns)) return;
  } else
#pragma unroll 1      // ptxas otherwise unrolls the spin 64x at every call site: the kernels are instruction-cache-bound enough
  for (uint32_t i = 0; i < (1u << 22); ++i)
    if (mbar_try_wait(bar, parity)) return;
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(abort_addr), "r"(1u) : "memory");
  atomicExch(error_flag, code);
#endif
}
// Simple bounded wait for the probes (returns false on timeout).
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t i = 0; i < (1u << 22); ++i)
    if (mbar_try_wait(bar, parity)) return true;
  return false;
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem], TF32 inputs, FP32 accumulate. Issued by ONE thread.
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, with the descriptors given as their LOW words (start address >> 4 | LBO << 16, see make_desc_base) and one shared
// HIGH word: issue loops then advance a descriptor with ONE uniform add of a value in 16-byte units instead of
// re-encoding (shift, mask, or) the address for every instruction.
__device__ __forceinline__ void mma_tf32_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f16 (fp16 operands, K = 16 per instruction, fp32 accumulate): the product path.  Same descriptor conventions as
// kind::tf32 (profiles/r02_f16_probe.log), twice the K per instruction at the same issue rate.
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_f16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The same MMA with a COLLECTOR hint for the A operand: the tensor core keeps the A tile it fetched from shared memory in its
// collector buffer (FILL), later MMAs with the SAME A descriptor read it from there instead of shared memory (USE), the last one
// releases it (LASTUSE).  SS-mode MMAs are shared-memory-bandwidth-bound (A 4 KB + B 32 N bytes per instruction at 128 B/clk):
// when consecutive instructions share A - the taps of a weight-gradient stencil, the hi*lo / hi*hi products of a tile - this
// takes most of the A bytes off the shared-memory port.  SASS: UTCHMMA gdesc[..].A_KEEP / .A_REUSE.A_KEEP / .A_REUSE.
enum : int { CU_NONE = 0, CU_FILL = 1, CU_USE = 2, CU_LASTUSE = 3 };
template <int CU, bool PAIR>
__device__ __forceinline__ void mma_f16_lo_cu(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
#define SSDN_MMA_CU(GROUP, SUFFIX)                                                                                      \
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t" \
               "tcgen05.mma.cta_group::" GROUP ".kind::f16" SUFFIX " [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),        \
               "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)                                          \
               : "memory")
#ifdef SSDN_NO_COLLECTOR          // ablation build: the same instruction stream without the hints
  if (PAIR) SSDN_MMA_CU("2", ""); else SSDN_MMA_CU("1", "");
#else
  if (PAIR) {
    if (CU == CU_FILL) SSDN_MMA_CU("2", ".collector::a::fill"); else if (CU == CU_USE) SSDN_MMA_CU("2", ".collector::a::use");
    else if (CU == CU_LASTUSE) SSDN_MMA_CU("2", ".collector::a::lastuse"); else SSDN_MMA_CU("2", "");
  } else {
    if (CU == CU_FILL) SSDN_MMA_CU("1", ".collector::a::fill"); else if (CU == CU_USE) SSDN_MMA_CU("1", ".collector::a::use");
    else if (CU == CU_LASTUSE) SSDN_MMA_CU("1", ".collector::a::lastuse"); else SSDN_MMA_CU("1", "");
  }
#endif
#undef SSDN_MMA_CU
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets lane (lane_base+i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// ---------------------------------------------------------------- CTA pairs (cta_group::2)
// Two CTAs of a cluster on the two SMs of a TPC execute ONE tcgen05.mma of M = 256: each CTA supplies its own 128 rows of
// A and HALF of the N rows of B from its own shared memory (same offsets in both), and receives its 128 accumulator rows
// in its own TMEM.  Per CTA an MMA then reads 4 KB (A) + N/2 x 32 B (B) instead of 4 KB + N x 32 B - which is what takes
// N = 96 tiles off the shared-memory bandwidth limit.  Only the leader (rank 0) issues MMAs and commits.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive / expect_tx on a barrier given by a shared::cluster address (possibly in the peer CTA)
// RELAXED: the only thing this arrival orders is TMEM traffic, which tcgen05.fence::before_thread_sync (issued by the caller)
// and the waiter's tcgen05.fence::after_thread_sync take care of.  The .release.cluster form compiles to MEMBAR.ALL.GPU +
// ERRBAR, i.e. the epilogue warp stalls until every global store it has in flight is acknowledged - 10 % of the epilogue's
// time in the CTA-pair kernels (profiles/r02_epilogue_source_counters.txt).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA load into THIS CTA's shared memory whose completion bytes are signalled on a barrier of the pair's leader
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_tf32_lo_pair(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_f16_lo_pair(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair once all previously issued MMAs completed
__device__ __forceinline__ void mma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (sm_100 format, version 1).
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4   bits [46,48) version = 1
//   bits [49,52) base offset               bits [61,64) layout: 0 none, 2 SW128, 4 SW64, 6 SW32
enum : uint32_t { LAYOUT_NONE = 0, LAYOUT_SW128 = 2, LAYOUT_SW64 = 4, LAYOUT_SW32 = 6 };

__host__ __device__ constexpr uint64_t make_desc_base(uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                      uint32_t layout, uint32_t base_offset = 0) {
  return (uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16) | (uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32) |
         (uint64_t(1) << 46) | (uint64_t(base_offset & 7) << 49) | (uint64_t(layout & 7) << 61);
}
__device__ __forceinline__ uint64_t desc_at(uint64_t base, uint32_t smem_addr) {
  return base | uint64_t((smem_addr >> 4) & 0x3FFF);
}
// Instruction descriptor for kind::tf32 with fp32 accumulate.
//   c_format F32 (1) at [4,6), a/b format TF32 (2) at [7,10)/[10,13), a_major at 15, b_major at 16,
//   N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t M, uint32_t N, uint32_t a_mn_major,
                                                       uint32_t b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major & 1u) << 15) | ((b_mn_major & 1u) << 16) |
         ((N >> 3) << 17) | ((M >> 4) << 24);
}

// Instruction descriptor for kind::f16 with fp16 A / B (format 0) and fp32 accumulate: the field positions of make_idesc_tf32.
__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t M, uint32_t N, uint32_t a_mn_major, uint32_t b_mn_major) {
  return (1u << 4) | ((a_mn_major & 1u) << 15) | ((b_mn_major & 1u) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

}  // namespace umma

// ---------------------------------------------------------------- host: tensor-map encoder
#include <cudaTypedefs.h>
#include <stdio.h>
namespace umma {
inline PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) !=
            cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}
// fp32 tensor, rank<=5. dims/strides innermost first; strides in BYTES for dims 1..rank-1.
inline int encode_f32(CUtensorMap* m, void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle sw) {
  auto fn = get_encode_fn();
  if (!fn) return -1;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  cuuint64_t d[5], s[5];
  cuuint32_t b[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, base, d, s, b, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}
// fp16 tensor, rank<=5. dims/strides innermost first; strides in BYTES for dims 1..rank-1.  Out-of-bounds elements read as zero.
inline int encode_f16(CUtensorMap* m, void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle sw) {
  auto fn = get_encode_fn();
  if (!fn) return -1;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  cuuint64_t d[5], s[5];
  cuuint32_t b[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, base, d, s, b, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}
}  // namespace umma
