// The blind-spot / plain U-Net of the reference (ssdn/ssdn/models/noise_network.py:48-226) as a
// static plan of kernel launches over one caller-owned workspace.
//
// Concatenations, nearest-neighbour upsampling, the final Shift2d + un-rotate + channel concat and
// their backward counterparts are never separate passes: every producer writes straight into the
// (strided) slot of the consumer's concat buffer from its epilogue, and every consumer reads
// channel-slice views through TMA tensor maps.  See common.cuh for the memory layout.
#pragma once
#include <string>
#include <vector>

#include "engine.cuh"

namespace net {

using eng::Arena;
using eng::round_up;

struct Buf {                 // one padded-flat tensor: a GEMM operand (two scaled fp16 planes + a scale slot) or a raw fp32 plane
  __half* hi = nullptr; __half* lo = nullptr;     // operand planes
  float* v = nullptr;                             // raw fp32 plane (gradients that only feed pointwise kernels)
  ScaleRef sc{}; int sid = -1;                    // scale slot of an operand tensor
  Geom g{}; int cpitch = 0;
  uint32_t* mask = nullptr; int mask_words = 0;   // LeakyReLU sign bits [flat pixel][words] (buffers whose derivative a dgrad epilogue applies)
  size_t elems() const { return (size_t)g.total() * cpitch; }
};

// Scale slots of a network (common.cuh): activations, gradients, weights.  Each range is begun / finished by its own pass.
constexpr int kSlotA = 0, kSlotG = 32, kSlotW = 64, kSlots = 96;
constexpr int kInitActScale = 8;                  // first forward pass of a plan: activations are assumed to peak near 2^3
__global__ void scale_init_kernel(pw::ScaleState st, int first, int count, int value) {
  const int i = first + threadIdx.x;
  if (threadIdx.x < count) { st.k[i] = value; st.k_next[i] = value; }
}

struct Layer {               // one convolution of the network
  std::string name;
  int cin, cout, ksize;       // GEMM view: the first conv is (9 * Cin, 48, 1) over the im2col'd input (same flat weight order)
  bool im2col = false;
  size_t w_off, b_off;       // offsets (floats) into the flat parameter / gradient buffers
  // forward
  ConvPlan fwd; __half* slab_f; int n_f; int coutp_f;
  // data gradient (has_dgrad == false for the first conv: nobody consumes d(input))
  bool has_dgrad = false;
  ConvPlan dgrad; __half* slab_d; int n_d; int cinp_d; int dgrad_nvalid;
  bool bias_fused = false; int bias_nblk = 0;   // bias gradient = column sums written by the kernel that produced this layer's dZ
  const float* bias_partial = nullptr;          //   partial[bias_nblk][cout]
  float* bias_buf = nullptr;                    //   this layer's own partial buffer (all bias reductions run as one kernel at the end)
  // weight gradient
  WgradPlan wgrad; int ksplit; float* partial = nullptr;   // this layer's own K-split partial buffer
  const Buf* x; int x_coff;          // conv input (channel slice [x_coff, x_coff+cin))
  const Buf* dz;                     // gradient w.r.t. this conv's pre-activation output
};

struct PoolOp { const Buf* src; const Buf* dst; int dst_coff; };
struct PoolBwdOp { const Buf* act; const Buf* g1; const Buf* g2; int g2_coff; const Buf* dz; Geom gp; float* colsum; int grid; };
struct UpBwdOp { const Buf* g; const Buf* act_up; const Buf* dz; int C; float* colsum; int grid; };
// pool_bwd / up_bwd / the loss-gradient pack also produce a dZ: their fused column sums go to the owning layer's bias_buf

class Net {
 public:
  int N, Cin, Cout, H, W; bool blind; int B;      // B = images inside the network (4N when blind)
  int sms;
  Geom g[6], gh;                                   // padded geometries of the 6 levels, dense head geometry
  std::vector<Layer> layers;                       // in PARAMETER order (noise_network.py registration order)
  size_t n_params = 0;
  // buffers
  Buf cat[6], e[6], e1a, p5, d_a[6], head_in, h1, h2, xcol;
  Buf g_out, dz_h2, dz_h1, dz_db[6], dz_da[6], gcat[6], dz_e[7], dz_e1a, g_p[6];
  float* partial = nullptr; float* colpart = nullptr; int* flag = nullptr;
  pw::ScaleState scales{}; int n_slot_a = 0, n_slot_g = 0;   // operand-scale state (device) and the number of slots per range
  int fwd_runs = 0, bwd_runs = 0;                            // passes launched so far (the first ones seed the scales)
  size_t partial_floats = 0;
  size_t ws_bytes = 0; void* ws = nullptr;
  std::vector<PoolOp> pools; std::vector<PoolBwdOp> pool_bwds; std::vector<UpBwdOp> up_bwds;
  int nin;                                         // head width (384 blind, 96 plain)
  // Weight gradients are off the critical path of the backward pass (nothing downstream reads them before the optimiser):
  // they run on a second stream, forked after the dZ they consume is ready and joined at the end of backward(), so that
  // they fill the SMs the small pyramid levels leave idle and hide each other's launch / drain latency.
  cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr; bool use_side = true;

  Net(int n, int cin, int cout, int h, int w, bool blind_, int sms_) : N(n), Cin(cin), Cout(cout), H(h), W(w), blind(blind_), sms(sms_) {
    B = blind ? 4 * N : N;
    nin = blind ? 384 : 96;
    for (int l = 0; l < 6; ++l) g[l] = make_geom(B, H >> l, W >> l, true);
    gh = make_geom(N, H, W, false);
    build_layers();
    ws_bytes = carve(nullptr, 0);
    use_side = !(getenv("SSDN_WGRAD_STREAM") && atoi(getenv("SSDN_WGRAD_STREAM")) == 0);
  }
  ~Net() {
    if (side) cudaStreamDestroy(side);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
  }
  Net(const Net&) = delete;
  Net& operator=(const Net&) = delete;

  const Buf* find_buf(const std::string& nm) const {
    auto idx = [&](const char* pre) -> int { return (nm.rfind(pre, 0) == 0 && nm.size() == strlen(pre) + 1) ? nm.back() - '0' : -1; };
    int i;
    if (nm == "e1a") return &e1a; if (nm == "xcol") return &xcol; if (nm == "p5") return &p5; if (nm == "head_in") return &head_in;
    if (nm == "h1") return &h1; if (nm == "h2") return &h2; if (nm == "g_out") return &g_out;
    if (nm == "dz_h2") return &dz_h2; if (nm == "dz_h1") return &dz_h1; if (nm == "dz_e1a") return &dz_e1a;
    if ((i = idx("cat")) >= 1 && i <= 5) return &cat[i];
    if ((i = idx("gcat")) >= 1 && i <= 5) return &gcat[i];
    if ((i = idx("e")) >= 1 && i <= 5) return &e[i];
    if ((i = idx("d_a")) >= 1 && i <= 5) return &d_a[i];
    if ((i = idx("dz_db")) >= 1 && i <= 5) return &dz_db[i];
    if ((i = idx("dz_da")) >= 1 && i <= 5) return &dz_da[i];
    if ((i = idx("dz_e")) >= 1 && i <= 6) return &dz_e[i];
    if ((i = idx("g_p")) >= 1 && i <= 5) return &g_p[i];
    return nullptr;
  }

  Layer& L(const std::string& name) { for (auto& l : layers) if (l.name == name) return l; return layers[0]; }

  void build_layers() {
    auto add = [&](const std::string& nm, int ci, int co, int k) {
      Layer l; l.name = nm; l.cin = ci; l.cout = co; l.ksize = k;
      l.w_off = n_params; n_params += (size_t)co * ci * k * k;
      l.b_off = n_params; n_params += co;
      layers.push_back(l);
    };
    // the first conv as a 1x1 GEMM over xcol (pw::pack_im2col3x3_kernel): a [48][Cin][3][3] weight IS a [48][9 Cin] matrix
    add("encode_block_1.0", 9 * Cin, 48, 1); layers.back().im2col = true;
    add("encode_block_1.2", 48, 48, 3);
    for (int i = 2; i <= 6; ++i) add("encode_block_" + std::to_string(i) + ".0", 48, 48, 3);
    add("decode_block_5.0", 96, 96, 3); add("decode_block_5.2", 96, 96, 3);
    for (int i = 4; i >= 2; --i) { add("decode_block_" + std::to_string(i) + ".0", 144, 96, 3); add("decode_block_" + std::to_string(i) + ".2", 96, 96, 3); }
    add("decode_block_1.0", 96 + Cin, 96, 3); add("decode_block_1.2", 96, 96, 3);
    add("output_conv", 96, Cout, 1); add("output_block.0", nin, nin, 1); add("output_block.2", nin, 96, 1);
  }

  // Lays every buffer out in the workspace; with base == nullptr only measures.
  size_t carve(void* base, size_t cap) {
    Arena a(base, cap);
    // scale state first: its addresses go into every operand buffer
    scales.k = a.take<int>(kSlots); scales.k_next = a.take<int>(kSlots); scales.amax = a.take<unsigned>(kSlots);
    scales.status = a.take<int>(8); scales.counter = a.take<unsigned>(1);
    n_slot_a = 0; n_slot_g = 0;
    auto slot = [&](Buf& b, int sid) { b.sid = sid; b.sc.k = scales.k ? scales.k + sid : nullptr; b.sc.amax = scales.amax ? scales.amax + sid : nullptr; };
    // kind: 0 raw fp32 plane, 1 activation operand, 2 gradient operand.  Plane pairs are one allocation (+ slack: the weight
    // gradient's 64-channel TMA blocks may run past the last pixel's channels)
    auto mk = [&](Buf& b, const Geom& gg, int cp, int kind) {
      b.g = gg; b.cpitch = cp; b.hi = b.lo = nullptr; b.v = nullptr;
      if (kind == 0) { b.v = a.take<float>(b.elems()); return; }
      const size_t n = (b.elems() + 127) / 128 * 128;
      __half* p = a.take<__half>(2 * n + 256);
      b.hi = p; b.lo = p ? p + n : nullptr;
      slot(b, kind == 1 ? kSlotA + n_slot_a++ : kSlotG + n_slot_g++);
    };
    auto mkmask = [&](Buf& b) { b.mask_words = (b.cpitch + 31) / 32; b.mask = a.take<uint32_t>((size_t)b.g.total() * b.mask_words); };
    // 16 channels = 32 bytes: every pixel of the largest concat buffer starts on a sector boundary, so the upsampling epilogue
    // that fills it writes whole sectors (with a pitch of 104 channels it wrote 25 % more sectors, half of them partial)
    const int c1 = round_up(96 + Cin, 16);
    mk(cat[1], g[0], c1, 1); mk(e1a, g[0], 48, 1); mk(e[1], g[0], 48, 1);
    mk(xcol, g[0], round_up(9 * Cin, 8), 1);
    mk(cat[2], g[1], 144, 1); mk(e[2], g[1], 48, 1);
    mk(cat[3], g[2], 144, 1); mk(e[3], g[2], 48, 1);
    mk(cat[4], g[3], 144, 1); mk(e[4], g[3], 48, 1);
    mk(cat[5], g[4], 96, 1); mk(e[5], g[4], 48, 1);
    mk(p5, g[5], 48, 1);
    for (int l = 1; l <= 5; ++l) mk(d_a[l], g[l - 1], 96, 1);        // dec{l}a output lives at level l-1
    mk(head_in, gh, nin, 1); mk(h1, gh, nin, 1); mk(h2, gh, 96, 1);
    for (int l = 1; l <= 5; ++l) mkmask(d_a[l]);
    mkmask(e1a); mkmask(head_in); mkmask(h1); mkmask(h2);
    // backward
    mk(g_out, gh, round_up(Cout, 8), 2); mk(dz_h2, gh, 96, 2); mk(dz_h1, gh, nin, 2);
    for (int l = 1; l <= 5; ++l) { mk(dz_db[l], g[l - 1], 96, 2); mk(dz_da[l], g[l - 1], 96, 2); }
    mk(gcat[1], g[0], 96, 0);
    for (int l = 2; l <= 4; ++l) mk(gcat[l], g[l - 1], 144, 0);
    mk(gcat[5], g[4], 96, 0);
    for (int l = 1; l <= 5; ++l) mk(dz_e[l], g[l - 1], 48, 2);       // dZ of the conv that feeds pool l
    mk(dz_e[6], g[5], 48, 2); mk(dz_e1a, g[0], 48, 2);
    for (int l = 1; l <= 5; ++l) mk(g_p[l], g[l], 48, 0);           // d(pool l) coming from the next encoder conv
    // weight slabs
    for (auto& l : layers) {
      const int nt = l.ksize * l.ksize;
      l.coutp_f = round_up(l.cout, 16);
      l.n_f = (l.cout == 384) ? (l.ksize == 1 ? 192 : 96) : eng::pick_n(l.coutp_f);   // 1x1: N = 192 halves the re-reads of A
      size_t f = conv_weight_slab_halves(l.cin, l.coutp_f, nt);
      l.slab_f = a.take<__half>(f);
      l.has_dgrad = (l.name != "encode_block_1.0");
      l.dgrad_nvalid = (l.name == "decode_block_1.0") ? 96 : l.cin;       // d(x) part of the last concat is never used
      l.cinp_d = round_up(l.dgrad_nvalid, 16);
      // data-gradient N tiles: the un-rotating head conv needs one tile per rotation branch (96 channels)
      l.n_d = (l.cinp_d == 384) ? ((l.name == "output_block.0" || l.ksize != 1) ? 96 : 192) : eng::pick_n(l.cinp_d);
      if (l.cinp_d == 144) l.n_d = 144;   // one N=144 tile (T = 1) instead of three smem-bound N=48 tiles
      size_t fd = conv_weight_slab_halves(l.cout, l.cinp_d, nt);
      l.slab_d = a.take<__half>(fd);
    }
    // wgrad partials: one buffer per layer, so that a single batched reduction at the end of backward() finishes them all
    partial_floats = 0;
    for (auto& l : layers) {
      const Geom& gg = l.im2col ? g[0] : (l.ksize == 1) ? gh : g[level_of(l.name)];
      l.ksplit = wgrad_pick_ksplit(gg.total(), l.cout, l.cin, l.ksize * l.ksize, sms, gg.P);
      const size_t pf = wgrad_partial_floats(l.ksplit, l.ksize * l.ksize, l.cout, l.cin);
      l.partial = a.take<float>(pf);
      partial_floats += pf;
    }
    partial = nullptr;
    colpart = a.take<float>((size_t)1024 * 384);
    for (auto& l : layers) l.bias_buf = a.take<float>((size_t)std::max(sms * convk::kEpiWarps, pw::kFusedColsumGrid) * l.cout);
    flag = a.take<int>(64);
    return a.off;
  }

  int level_of(const std::string& nm) const {   // spatial level (0 = full resolution) at which a 3x3 conv runs
    if (nm.rfind("encode_block_", 0) == 0) return nm[13] - '1';
    if (nm.rfind("decode_block_", 0) == 0) return nm[13] - '1';
    return 0;
  }

  // ------------------------------------------------------------------ plan construction
  int bind(void* workspace, size_t bytes, cudaStream_t st) {
    if (bytes < ws_bytes) return eng::fail(-3, "workspace too small: need %zu bytes, have %zu", ws_bytes, bytes);
    ws = workspace;
    carve(workspace, bytes);
    cudaError_t ce = cudaMemsetAsync(workspace, 0, ws_bytes, st);   // zero halos (never written afterwards)
    if (ce != cudaSuccess) return eng::fail(-2, "memset failed: %s", cudaGetErrorString(ce));
    pools.clear(); pool_bwds.clear(); up_bwds.clear();
    launch_pdl(scale_init_kernel, dim3(1), dim3(32), 0, st, scales, kSlotA, n_slot_a, kInitActScale);
    fwd_runs = bwd_runs = 0;
    int r;
    const int act = EP_BIAS | EP_LRELU | EP_WRITE_LO;
    // ---- encoder
    if ((r = plan_fwd(L("encode_block_1.0"), xcol, 0, dst(e1a, 0, MAP_IDENT, act, 48)))) return r;
    if ((r = plan_fwd(L("encode_block_1.2"), e1a, 0, dst(e[1], 0, MAP_IDENT, act, 48)))) return r;
    pools.push_back({&e[1], &cat[2], 96});
    for (int i = 2; i <= 4; ++i) {
      if ((r = plan_fwd(L("encode_block_" + std::to_string(i) + ".0"), cat[i], 96, dst(e[i], 0, MAP_IDENT, act, 48)))) return r;
      pools.push_back({&e[i], &cat[i + 1], i == 4 ? 48 : 96});
    }
    if ((r = plan_fwd(L("encode_block_5.0"), cat[5], 48, dst(e[5], 0, MAP_IDENT, act, 48)))) return r;
    pools.push_back({&e[5], &p5, 0});
    if ((r = plan_fwd(L("encode_block_6.0"), p5, 0, dst(cat[5], 0, MAP_UP2, act, 48)))) return r;
    // ---- decoder
    for (int i = 5; i >= 1; --i) {
      const std::string a_nm = "decode_block_" + std::to_string(i) + ".0", b_nm = "decode_block_" + std::to_string(i) + ".2";
      if ((r = plan_fwd(L(a_nm), cat[i], 0, dst(d_a[i], 0, MAP_IDENT, act, 96)))) return r;
      ConvDst d2 = (i > 1) ? dst(cat[i - 1], 0, MAP_UP2, act, 96)
                           : dst(head_in, 0, blind ? MAP_UNROT : MAP_IDENT, act, 96);
      if ((r = plan_fwd(L(b_nm), d_a[i], 0, d2))) return r;
    }
    // ---- head
    if ((r = plan_fwd(L("output_block.0"), head_in, 0, dst(h1, 0, MAP_IDENT, act, nin)))) return r;
    if ((r = plan_fwd(L("output_block.2"), h1, 0, dst(h2, 0, MAP_IDENT, act, 96)))) return r;
    { ConvDst d = dst(h2, 0, MAP_NCHW, EP_BIAS, Cout); d.v = nullptr; d.hi = d.lo = nullptr; d.g = gh;
      if ((r = plan_fwd(L("output_conv"), h2, 0, d))) return r; }

    // ---- backward: data gradients
    const int gact = EP_ACT_GRAD | EP_WRITE_LO;
    // every dst_act below also yields the bias gradient of the layer that owns the produced dZ (fused column sums)
    auto fused = [&](const char* producer, const char* owner) {
      Layer& o = L(owner); o.bias_fused = true; o.bias_partial = o.bias_buf;
      o.bias_nblk = L(producer).dgrad.grid * 4 * L(producer).dgrad.p.epi_per_quad;
      if ((size_t)o.bias_nblk > (size_t)std::max(sms * convk::kEpiWarps, pw::kFusedColsumGrid)) o.bias_nblk = -1;   // checked below
    };
    // pointwise producers (grid-stride kernels with at most kFusedColsumGrid blocks of kFusedColsumBlock threads)
    auto pw_grid = [&](long long n) { return (int)std::min<long long>(pw::kFusedColsumGrid, (n + pw::kFusedColsumBlock - 1) / pw::kFusedColsumBlock); };
    auto fused_pw = [&](const std::string& owner, int nblk) -> float* {
      Layer& o = L(owner); o.bias_fused = true; o.bias_nblk = nblk; o.bias_partial = o.bias_buf; return o.bias_buf;
    };
    fused_pw("output_conv", N);                                              // nchw_colsum_kernel over d(loss)/d(out)
    if ((r = plan_dgrad(L("output_conv"), g_out, dst_act(dz_h2, MAP_IDENT, gact, 96, h2, &L("output_block.2"))))) return r;
    fused("output_conv", "output_block.2");
    if ((r = plan_dgrad(L("output_block.2"), dz_h2, dst_act(dz_h1, MAP_IDENT, gact, nin, h1, &L("output_block.0"))))) return r;
    fused("output_block.2", "output_block.0");
    { ConvDst d = dst_act(dz_db[1], blind ? MAP_UNROT_INV : MAP_IDENT, gact | EP_ACT_AT_SRC, 96, head_in, &L("decode_block_1.2"));
      if ((r = plan_dgrad(L("output_block.0"), dz_h1, d))) return r; }
    fused("output_block.0", "decode_block_1.2");
    for (int i = 1; i <= 5; ++i) {
      const std::string a_nm = "decode_block_" + std::to_string(i) + ".0", b_nm = "decode_block_" + std::to_string(i) + ".2";
      if ((r = plan_dgrad(L(b_nm), dz_db[i], dst_act(dz_da[i], MAP_IDENT, gact, 96, d_a[i], &L(a_nm))))) return r;
      fused(b_nm.c_str(), a_nm.c_str());
      if ((r = plan_dgrad(L(a_nm), dz_da[i], dst(gcat[i], 0, MAP_IDENT, 0, L(a_nm).dgrad_nvalid)))) return r;
      if (i < 5) {
        const int grid = pw_grid((long long)g[i].B * g[i].H * g[i].W * (96 / 8));
        up_bwds.push_back({&gcat[i], &cat[i], &dz_db[i + 1], 96, fused_pw("decode_block_" + std::to_string(i + 1) + ".2", grid), grid});
      }
    }
    { const int grid = pw_grid((long long)g[5].B * g[5].H * g[5].W * (48 / 8));
      up_bwds.push_back({&gcat[5], &cat[5], &dz_e[6], 48, fused_pw("encode_block_6.0", grid), grid}); }
    if ((r = plan_dgrad(L("encode_block_6.0"), dz_e[6], dst(g_p[5], 0, MAP_IDENT, 0, 48)))) return r;
    for (int i = 5; i >= 2; --i) {
      // dZ of the conv feeding pool i: gradient from the next encoder conv (+ the skip path for i <= 4)
      { const int grid = pw_grid((long long)g[i].B * g[i].H * g[i].W * (48 / 8));
        pool_bwds.push_back({&e[i], &g_p[i], i <= 4 ? &gcat[i + 1] : nullptr, i == 4 ? 48 : 96, &dz_e[i], g[i],
                             fused_pw("encode_block_" + std::to_string(i) + ".0", grid), grid}); }
      if ((r = plan_dgrad(L("encode_block_" + std::to_string(i) + ".0"), dz_e[i], dst(g_p[i - 1], 0, MAP_IDENT, 0, 48)))) return r;
    }
    { const int grid = pw_grid((long long)g[1].B * g[1].H * g[1].W * (48 / 8));
      pool_bwds.push_back({&e[1], &g_p[1], &gcat[2], 96, &dz_e[1], g[1], fused_pw("encode_block_1.2", grid), grid}); }
    if ((r = plan_dgrad(L("encode_block_1.2"), dz_e[1], dst_act(dz_e1a, MAP_IDENT, gact, 48, e1a, &L("encode_block_1.0"))))) return r;
    fused("encode_block_1.2", "encode_block_1.0");

    // ---- backward: weight gradients
    struct WG { const char* nm; const Buf* x; int xoff; const Buf* dz; };
    const WG wg[] = {
        {"encode_block_1.0", &xcol, 0, &dz_e1a}, {"encode_block_1.2", &e1a, 0, &dz_e[1]},
        {"encode_block_2.0", &cat[2], 96, &dz_e[2]}, {"encode_block_3.0", &cat[3], 96, &dz_e[3]},
        {"encode_block_4.0", &cat[4], 96, &dz_e[4]}, {"encode_block_5.0", &cat[5], 48, &dz_e[5]},
        {"encode_block_6.0", &p5, 0, &dz_e[6]},
        {"decode_block_5.0", &cat[5], 0, &dz_da[5]}, {"decode_block_5.2", &d_a[5], 0, &dz_db[5]},
        {"decode_block_4.0", &cat[4], 0, &dz_da[4]}, {"decode_block_4.2", &d_a[4], 0, &dz_db[4]},
        {"decode_block_3.0", &cat[3], 0, &dz_da[3]}, {"decode_block_3.2", &d_a[3], 0, &dz_db[3]},
        {"decode_block_2.0", &cat[2], 0, &dz_da[2]}, {"decode_block_2.2", &d_a[2], 0, &dz_db[2]},
        {"decode_block_1.0", &cat[1], 0, &dz_da[1]}, {"decode_block_1.2", &d_a[1], 0, &dz_db[1]},
        {"output_conv", &h2, 0, &g_out}, {"output_block.0", &head_in, 0, &dz_h1}, {"output_block.2", &h1, 0, &dz_h2}};
    for (const WG& w : wg) {
      Layer& l = L(w.nm);
      l.x = w.x; l.x_coff = w.xoff; l.dz = w.dz;
      const Geom& gg = w.x->g;
      ConvTaps taps = eng::make_taps(l.ksize, blind, false, gg.P);
      if ((r = wgrad_plan_init(&l.wgrad, gg.total(), w.dz->hi, w.dz->lo, w.dz->cpitch, 0, l.cout, w.x->hi, w.x->lo, w.x->cpitch,
                               w.xoff, l.cin, taps, l.ksplit, l.partial, flag, sms, w.dz->sc.k, w.x->sc.k)))
        return eng::fail(r, "wgrad plan for %s failed (%d)", w.nm, r);
      l.wgrad.flops = 2.0 * B0(gg) * l.cin * l.cout * l.ksize * l.ksize;
      // both operands' plane pairs once + the K-split partials
      l.wgrad.bytes = (double)gg.total() * 4.0 * (l.cin + l.cout) + 4.0 * wgrad_partial_floats(l.ksplit, l.ksize * l.ksize, l.cout, l.cin);
    }
    for (auto& l : layers) if (l.bias_fused && l.bias_nblk < 0) return eng::fail(-6, "bias partial buffer of %s too small", l.name.c_str());
    return 0;
  }

  static double B0(const Geom& gg) { return (double)gg.B * gg.H * gg.W; }   // valid pixels of a geometry

  ConvDst dst(Buf& b, int coff, int map, int flags, int cvalid) {
    ConvDst d{}; d.v = b.v; d.hi = b.hi; d.lo = b.lo; d.scale = b.sc; d.cpitch = b.cpitch; d.coff = coff; d.g = b.g; d.map = map; d.flags = flags;
    if (!b.hi) d.flags &= ~EP_WRITE_LO;
    d.cvalid = cvalid; d.nimg = N; d.bias = nullptr;
    if (b.mask && (flags & EP_LRELU) && (map == MAP_IDENT || map == MAP_UNROT)) { d.mask_out = b.mask; d.mask_out_words = b.mask_words; }
    return d;
  }
  // data-gradient destination whose values are multiplied by LeakyReLU'(act) (sign masks of `act`); with_colsum also
  // collects the column sums of what is written (= the bias gradient of the layer whose dZ this is)
  ConvDst dst_act(Buf& b, int map, int flags, int cvalid, Buf& act, Layer* bias_owner) {
    ConvDst d = dst(b, 0, map, flags, cvalid); d.mask_in = act.mask; d.mask_in_words = act.mask_words;
    if (bias_owner) { d.colsum = bias_owner->bias_buf; d.colsum_pitch = (map == MAP_UNROT_INV) ? 96 : cvalid; }
    return d;
  }

  // algorithmic HBM bytes of one conv launch: the source plane pair once (halo included) + what the epilogue writes
  // (plane pairs or fp32, x4 when upsampling) + sign-mask words in / out
  static double conv_bytes(const Geom& sg, int cin, const ConvDst& d, int cout) {
    const double px = B0(sg);
    double b = (double)sg.total() * cin * 4.0;
    const double out_px = px * (d.map == MAP_UP2 ? 4.0 : 1.0);
    b += out_px * cout * 4.0;                         // two fp16 planes or one fp32 plane: 4 bytes per element either way
    if (d.mask_out) b += px * ((cout + 31) / 32) * 4.0;
    if (d.mask_in) b += px * d.mask_in_words * 4.0;
    return b;
  }
  int layer_index(const Layer& l) const { return (int)(&l - layers.data()); }
  const int* k_w(const Layer& l) const { return scales.k ? scales.k + kSlotW + layer_index(l) : nullptr; }

  int plan_fwd(Layer& l, Buf& src, int coff, ConvDst d) {
    ConvTaps taps = eng::make_taps(l.ksize, blind, false, src.g.P);
    int r = conv_plan_init(&l.fwd, src.g, src.hi, src.lo, src.cpitch, coff, l.cin, l.slab_f, l.coutp_f, l.n_f, taps, d,
                           flag, sms, src.sc.k, k_w(l));
    if (r) return eng::fail(r, "forward plan for %s failed (%d)", l.name.c_str(), r);
    l.fwd.flops = 2.0 * B0(src.g) * l.cin * l.cout * l.ksize * l.ksize;
    l.fwd.bytes = conv_bytes(src.g, l.cin, d, l.cout);
    return 0;
  }
  int plan_dgrad(Layer& l, Buf& src, ConvDst d) {
    ConvTaps taps = eng::make_taps(l.ksize, blind, true, src.g.P);
    int r = conv_plan_init(&l.dgrad, src.g, src.hi, src.lo, src.cpitch, 0, l.cout, l.slab_d, l.cinp_d, l.n_d, taps, d,
                           flag, sms, src.sc.k, k_w(l));
    if (r) return eng::fail(r, "dgrad plan for %s failed (%d)", l.name.c_str(), r);
    l.dgrad.flops = 2.0 * B0(src.g) * l.dgrad_nvalid * l.cout * l.ksize * l.ksize;
    l.dgrad.bytes = conv_bytes(src.g, l.cout, d, l.dgrad_nvalid);
    return 0;
  }

  // ------------------------------------------------------------------ execution
  int prep_weights(const float* params, cudaStream_t st, bool with_dgrad) {
    // exact per-layer scales first (weights are leaves), then all slabs in one launch
    pw::WeightScaleJobs sj{};
    int ns = 0;
    for (auto& l : layers) sj.j[ns++] = {params + l.w_off, l.cout * l.cin * l.ksize * l.ksize, kSlotW + layer_index(l)};
    SSDN_PROF(K_WEIGHT_PREP, 0, 4.0 * n_params, st,
              (launch_pdl(pw::weight_scale_kernel, dim3(pw::kWeightScaleBlocks, ns), dim3(256), 0, st, sj, scales, kSlotA, n_slot_a)));
    pw::WeightPrepJobs jobs{};
    int nj = 0;
    for (auto& l : layers) {
      const int nt = l.ksize * l.ksize;
      bool wide = conv_is_wide(l.cin, nt);
      int nc, kl; conv_chunks(l.cin, &nc, &kl, wide);
      jobs.j[nj++] = {params + l.w_off, l.slab_f, l.cout, l.cin, nt, l.cout, l.cin, l.coutp_f / l.n_f, nc, l.n_f, 0, wide ? 64 : 32, kSlotW + layer_index(l)};
      if (with_dgrad && l.has_dgrad) {
        wide = conv_is_wide(l.cout, nt);
        conv_chunks(l.cout, &nc, &kl, wide);
        jobs.j[nj++] = {params + l.w_off, l.slab_d, l.cout, l.cin, nt, l.dgrad_nvalid, l.cout, l.cinp_d / l.n_d, nc, l.n_d, 1, wide ? 64 : 32, kSlotW + layer_index(l)};
      }
    }
    double slab_bytes = 0;
    for (int j = 0; j < nj; ++j) slab_bytes += 2.0 * 2.0 * jobs.j[j].n_tiles * jobs.j[j].n_chunks * jobs.j[j].ntaps * jobs.j[j].N * jobs.j[j].CW;
    // the slabs are independent of the input pack that follows: they are prepared on the side stream (forward() joins it before
    // the first convolution), two latency-bound kernels side by side instead of one after the other
    cudaStream_t ps = st;
    prep_forked = false;
    if (use_side && !profiler().on) {
      int r = ensure_side(); if (r) return r;
      SSDN_CUDA(cudaEventRecord(ev_fork, st));          // the weight scales (and the begun activation pass) are complete
      SSDN_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
      ps = side; prep_forked = true;
    }
    SSDN_PROF(K_WEIGHT_PREP, 0, slab_bytes + 4.0 * n_params * (with_dgrad ? 2 : 1), ps,
              (launch_pdl(pw::weight_prep_batched_kernel, dim3(64, nj), dim3(pw::kBlock), 0, ps, jobs, scales)));
    SSDN_CUDA(cudaGetLastError());
    return 0;
  }
  bool prep_forked = false;
  int ensure_side() {
    if (!side) {
      SSDN_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
      SSDN_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
      SSDN_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    }
    return 0;
  }

  int run_fwd(Layer& l, const float* params, cudaStream_t st, float* nchw_out = nullptr) {
    l.fwd.p.dst.bias = params + l.b_off;
    if (nchw_out) l.fwd.p.dst.v = nchw_out;
    SSDN_CUDA(conv_launch(l.fwd, st, K_CONV_FWD));
    return 0;
  }

  int forward(const float* params, const float* x, float* out, cudaStream_t st, bool training) {
    if (!ws) return eng::fail(-5, "network workspace not bound");
    int r;
    ++fwd_runs;                 // (the pass over the activation slots is begun by prep_weights' first kernel)
    if ((r = prep_weights(params, st, training))) return r;
    {   // the input: rotation stack -> its channel slot of the last concat buffer + the first conv's im2col operand, one kernel
      const int tiles = ((H + pw::kPackTileH - 1) / pw::kPackTileH) * ((W + pw::kPackTileW - 1) / pw::kPackTileW);
      const size_t smem = (size_t)Cin * pw::kPackRegH * pw::kPackRegPitch * sizeof(float);
      if (smem > 48 * 1024) return eng::fail(-7, "input with %d channels does not fit the pack kernel's tile", Cin);
      SSDN_PROF(K_PACK, 0, (double)N * Cin * H * W * 4.0 + (double)B * H * W * (round_up(Cin, 8) + xcol.cpitch) * 4.0, st,
                (launch_pdl(pw::pack_input_kernel, dim3(B * tiles), dim3(256), smem, st, x, cat[1].hi, cat[1].lo, cat[1].cpitch, 96, cat[1].sc,
                            xcol.hi, xcol.lo, xcol.cpitch, xcol.sc, N, Cin, H, W, g[0], blind ? 2 : 1, blind ? 1 : 0)));
    }
    if (prep_forked) { if ((r = join_side(st))) return r; }      // the weight slabs are ready
    size_t pi = 0;
    auto pool = [&]() {
      const PoolOp& p = pools[pi++];
      const long long n = (long long)p.dst->g.B * p.dst->g.H * p.dst->g.W * 6;
      // reads the full-resolution plane pair once, writes the pooled one
      SSDN_PROF(K_POOL_FWD, 0, B0(p.src->g) * 48 * 4.0 + B0(p.dst->g) * 48 * 4.0, st,
                (launch_pdl(pw::pool_fwd_kernel, dim3(pw::grid_for(n)), dim3(pw::kBlock), 0, st, p.src->hi, p.src->lo, p.src->g, p.src->cpitch, 0, p.src->sc, p.dst->hi, p.dst->lo,
                                                                            p.dst->g, p.dst->cpitch, p.dst_coff, p.dst->sc, 48, blind ? 1 : 0)));
    };
    if ((r = run_fwd(L("encode_block_1.0"), params, st))) return r;
    if ((r = run_fwd(L("encode_block_1.2"), params, st))) return r;
    pool();
    for (int i = 2; i <= 5; ++i) { if ((r = run_fwd(L("encode_block_" + std::to_string(i) + ".0"), params, st))) return r; pool(); }
    if ((r = run_fwd(L("encode_block_6.0"), params, st))) return r;
    for (int i = 5; i >= 1; --i) {
      if ((r = run_fwd(L("decode_block_" + std::to_string(i) + ".0"), params, st))) return r;
      if ((r = run_fwd(L("decode_block_" + std::to_string(i) + ".2"), params, st))) return r;
    }
    if ((r = run_fwd(L("output_block.0"), params, st))) return r;
    if ((r = run_fwd(L("output_block.2"), params, st))) return r;
    if ((r = run_fwd(L("output_conv"), params, st, out))) return r;
    SSDN_PROF(K_SCALE, 0, 0, st, (launch_pdl(pw::scale_finish_kernel, dim3(1), dim3(32), 0, st, scales, kSlotA, n_slot_a, 0, nullptr, 0, kSlotW, (int)layers.size())));
    SSDN_CUDA(cudaGetLastError());
    return 0;
  }

  int run_wgrad(Layer& l, float* grads, cudaStream_t st) {
    const int nt = l.ksize * l.ksize;
    // bias gradient: reduce the column-sum partials its dZ producer left behind (main stream: the next producer reuses them)
    if (!l.bias_fused) {
      const long long rows = l.dz->g.total();
      pw::colsum_launch(l.dz->hi, l.dz->lo, l.dz->sc.k, rows, l.dz->cpitch, 0, l.cout, colpart, grads + l.b_off, st);
    }   // else: backward() finishes all bias gradients with one batched reduction after the last producer
    cudaStream_t ws_ = st;
    if (use_side && !profiler().on) {   // per-launch profiling serialises everything on one stream
      { int rs = ensure_side(); if (rs) return rs; }
      SSDN_CUDA(cudaEventRecord(ev_fork, st));          // this layer's dZ (and everything before it) is complete
      SSDN_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
      ws_ = side;
    }
    (void)nt;
    launching_background() = (ws_ != st);      // off the critical path: default launch priority (common.cuh: main_priority)
    cudaError_t we = wgrad_launch(l.wgrad, ws_);     // its K-split partials are reduced by reduce_all_wgrads()
    launching_background() = false;
    SSDN_CUDA(we);
    return 0;
  }
  // dW of the layers [first, last) of `order` (names in launch order) from their K-split partials: one launch on the stream
  // the weight gradients ran on.  backward() calls it twice: for everything but the last two layers while the main stream
  // still has work, and for those two at the very end (a short tail instead of an 80 us one).
  int reduce_wgrads(float* grads, cudaStream_t st, bool last_two) {
    wgradk::WgradReduceJobs jobs{};
    int nj = 0, blocks = 0; double bytes = 0;
    for (auto& l : layers) {
      const bool tail = (l.name == "encode_block_1.0" || l.name == "encode_block_1.2");
      if (tail != last_two) continue;
      const int nt = l.ksize * l.ksize;
      const long long n = (long long)l.cout * l.wgrad.p.cin_pitch * nt;
      jobs.j[nj++] = {l.partial, grads + l.w_off, l.wgrad.p.ksplit, nt, l.cout, l.cin, l.wgrad.p.cin_pitch, blocks};
      blocks += (int)((n / 4 + 31) / 32);
      bytes += (double)n * 4 * (l.wgrad.p.ksplit + 1);
    }
    jobs.n_jobs = nj;
    launching_background() = (st == side);
    SSDN_PROF(K_WGRAD_REDUCE, 0, bytes, st, (launch_pdl(wgradk::wgrad_reduce_batched_kernel, dim3(blocks), dim3(32, 16), 0, st, jobs)));
    launching_background() = false;
    SSDN_CUDA(cudaGetLastError());
    return 0;
  }
  int join_side(cudaStream_t st) {
    if (side) { SSDN_CUDA(cudaEventRecord(ev_join, side)); SSDN_CUDA(cudaStreamWaitEvent(st, ev_join, 0)); }
    return 0;
  }
  int run_dgrad(Layer& l, cudaStream_t st) { SSDN_CUDA(conv_launch(l.dgrad, st, K_CONV_DGRAD)); return 0; }

  // grads: flat buffer with the layout of params; every element is overwritten.  stale_out (optional, device): receives 1.0f
  // when this step's forward or backward pass ran with operand scales outside their band (see common.cuh), else 0.0f.
  int backward(const float* params, const float* dout, float* grads, float* stale_out, cudaStream_t st) {
    if (!ws) return eng::fail(-5, "network workspace not bound");
    int r;
    const double dout_bytes = (double)N * Cout * H * W * 4.0;
    // (the gradient slots were begun by the previous backward pass's scale_finish_kernel, or by bind())
    // the loss gradient is a leaf: exact scale now; the very first backward pass seeds every gradient slot with it
    SSDN_PROF(K_SCALE, 0, dout_bytes, st,
              (launch_pdl(pw::leaf_scale_kernel, dim3(64), dim3(256), 0, st, dout, (long long)N * Cout * H * W, scales, g_out.sid, kSlotG, bwd_runs == 0 ? n_slot_g : 0)));
    ++bwd_runs;
    ScaleRef no_amax = g_out.sc; no_amax.amax = nullptr;
    SSDN_PROF(K_PACK, 0, 2 * dout_bytes, st,
              (launch_pdl(pw::pack_nchw_pixel_kernel, dim3(pw::grid_for((long long)N * H * W)), dim3(pw::kBlock), 0, st, dout, g_out.hi, g_out.lo, N, Cout, H, W, gh,
                                                                                                    g_out.cpitch, 0, 0, no_amax)));
    SSDN_PROF(K_BIAS, 0, dout_bytes, st, (launch_pdl(pw::nchw_colsum_kernel, dim3(Cout, N), dim3(256), 0, st, dout, Cout, H * W, L("output_conv").bias_buf)));
    auto both = [&](const std::string& nm, bool dgrad) -> int {
      Layer& l = L(nm);
      int rr = run_wgrad(l, grads, st);
      if (rr) return rr;
      return dgrad ? run_dgrad(l, st) : 0;
    };
    if ((r = both("output_conv", true))) return r;
    if ((r = both("output_block.2", true))) return r;
    if ((r = both("output_block.0", true))) return r;
    size_t ui = 0, qi = 0;
    auto up_bwd = [&]() {
      const UpBwdOp& u = up_bwds[ui++];
      const Geom& gl = u.dz->g;
      const long long n = (long long)gl.B * gl.H * gl.W * (u.C / 4);
      (void)n;
      // reads the fp32 gradient at the upsampled resolution (4 pixels per output) + one hi-plane sign per output, writes a plane pair
      SSDN_PROF(K_UP_BWD, 0, B0(gl) * u.C * (4 * 4.0 + 2.0 + 4.0), st,
                (launch_pdl(pw::up_bwd_kernel, dim3(u.grid), dim3(pw::kFusedColsumBlock), pw::kFusedColsumBlock * 8 * sizeof(float), st, 
                    u.g->v, u.g->g, u.g->cpitch, 0, u.act_up->hi, u.act_up->cpitch, 0, gl, u.dz->hi, u.dz->lo, u.dz->cpitch, 0, u.dz->sc, u.C, u.colsum)));
    };
    auto pool_bwd = [&]() {
      const PoolBwdOp& q = pool_bwds[qi++];
      // per pooled pixel and channel: 4 activations (plane pairs) in, 1 or 2 fp32 gradients in, 4 dZ values (plane pairs) out
      SSDN_PROF(K_POOL_BWD, 0, B0(q.gp) * 48 * (4 * 4.0 + (q.g2 ? 8.0 : 4.0) + 4 * 4.0), st,
                (launch_pdl(pw::pool_bwd_kernel, dim3(q.grid), dim3(pw::kFusedColsumBlock), pw::kFusedColsumBlock * 8 * sizeof(float), st, 
                    q.act->hi, q.act->lo, q.act->g, q.act->cpitch, 0, q.g1->v, q.g1->cpitch, 0, q.g2 ? q.g2->v : nullptr, q.g2 ? q.g2->cpitch : 0, q.g2_coff, q.gp,
                    q.dz->hi, q.dz->lo, q.dz->cpitch, 0, q.dz->sc, 48, blind ? 1 : 0, q.colsum)));
    };
    for (int i = 1; i <= 5; ++i) {
      if ((r = both("decode_block_" + std::to_string(i) + ".2", true))) return r;
      if ((r = both("decode_block_" + std::to_string(i) + ".0", true))) return r;
      up_bwd();
    }
    if ((r = both("encode_block_6.0", true))) return r;
    for (int i = 5; i >= 2; --i) {
      pool_bwd();
      if ((r = both("encode_block_" + std::to_string(i) + ".0", true))) return r;
    }
    if ((r = reduce_wgrads(grads, (side && use_side && !profiler().on) ? side : st, false))) return r;   // all layers launched so far
    pool_bwd();
    if ((r = both("encode_block_1.2", true))) return r;
    if ((r = both("encode_block_1.0", false))) return r;
    {   // all bias gradients: one batched fixed-order reduction of the per-layer column-sum partials
      pw::BiasJobs jobs{};
      int nj = 0, maxc = 0;
      for (auto& l : layers)
        if (l.bias_fused) { jobs.j[nj++] = {l.bias_partial, grads + l.b_off, l.bias_nblk, l.cout}; maxc = std::max(maxc, l.cout); }
      double pb = 0;
      for (int j = 0; j < nj; ++j) pb += 4.0 * jobs.j[j].nblk * jobs.j[j].C;
      if (nj) SSDN_PROF(K_BIAS, 0, pb, st, (launch_pdl(pw::colsum_stage2_batched_kernel, dim3((maxc + 31) / 32, nj), dim3(32, 32), 0, st, jobs)));
    }
    if ((r = reduce_wgrads(grads, (side && use_side && !profiler().on) ? side : st, true))) return r;
    if ((r = join_side(st))) return r;
    SSDN_PROF(K_SCALE, 0, 0, st, (launch_pdl(pw::scale_finish_kernel, dim3(1), dim3(32), 0, st, scales, kSlotG, n_slot_g, 1, stale_out, 1, 0, 0)));
    SSDN_CUDA(cudaGetLastError());
    return 0;
  }
};

}  // namespace net
