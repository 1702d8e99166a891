"""Developer check (GPU box): compare internal buffers of the network plan with oracle intermediates."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "selfsupervised-denoising_b200")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import cases as C
import ssdn_oracle as O
from ssdn import _engine as E

def rel(a, b): return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()

def oracle_trace(p, x, blind):
    """Forward with every pre-activation kept (retain_grad) -> dict name -> (z, act)."""
    conv = O.shift_conv2d if blind else O.conv2d_same
    T = {}
    def cl(name, t, c=conv):
        z = c(t, p[name + ".weight"], p[name + ".bias"]); z.retain_grad(); a = O.lrelu(z); T[name] = (z, a); return a
    if blind: x = O.rot4_stack(x)
    t = cl("encode_block_1.0", x); t = cl("encode_block_1.2", t)
    pools = [O.maxpool2(t, blind)]
    for i in (2, 3, 4, 5): pools.append(O.maxpool2(cl(f"encode_block_{i}.0", pools[-1]), blind))
    t = O.upsample2(cl("encode_block_6.0", pools[4]))
    for i, skip in ((5, pools[3]), (4, pools[2]), (3, pools[1]), (2, pools[0])):
        t = torch.cat((t, skip), 1); t = cl(f"decode_block_{i}.0", t); t = cl(f"decode_block_{i}.2", t); t = O.upsample2(t)
    t = torch.cat((t, x), 1); t = cl("decode_block_1.0", t); t = cl("decode_block_1.2", t)
    if blind: t = O.shift_unrot_concat(t)
    T["head_in"] = (None, t)
    t = cl("output_block.0", t, O.conv2d_same); t = cl("output_block.2", t, O.conv2d_same)
    return O.conv2d_same(t, p["output_conv.weight"], p["output_conv.bias"]), T

for name in sys.argv[1:] or ["net_plain_mono", "net_blind_rgb"]:
    cin, cout, blind, n, size = C.NETWORK_CASES[name]
    params, x, dout = C.network_inputs(name)
    order = O.param_order(cin, cout, blind)
    flat = torch.cat([params[k].reshape(-1) for k in order]).cuda()
    plan = E.NetPlan(n, cin, cout, size, size, blind, "cuda")
    out = plan.forward(flat, x.cuda(), training=True); grads = plan.backward(flat, dout.cuda()); plan.check()
    po = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    oo, T = oracle_trace(po, x, blind)
    oo.backward(dout)
    print("==", name, "out", rel(out.cpu(), oo.detach()))
    pairs = [("h2", "output_block.2", 96, False), ("h1", "output_block.0", 384 if blind else 96, False),
             ("d_a1", "decode_block_1.0", 96, False), ("e1a", "encode_block_1.0", 48, False), ("e1", "encode_block_1.2", 48, False),
             ("dz_h2", "output_block.2", 96, True), ("dz_h1", "output_block.0", 384 if blind else 96, True),
             ("dz_db1", "decode_block_1.2", 96, True), ("dz_da1", "decode_block_1.0", 96, True),
             ("dz_db2", "decode_block_2.2", 96, True), ("dz_da2", "decode_block_2.0", 96, True), ("dz_e6", "encode_block_6.0", 48, True),
             ("dz_e5", "encode_block_5.0", 48, True), ("dz_e1", "encode_block_1.2", 48, True), ("dz_e1a", "encode_block_1.0", 48, True)]
    for buf, lname, ch, is_grad in pairs:
        z, a = T[lname]
        ref = z.grad if is_grad else a.detach()
        got = plan.debug_read(buf, ch).cpu()
        lo = plan.debug_read(buf, ch, plane=1).cpu()
        hi = (got.view(torch.int32) & -8192).view(torch.float32)
        d = (got - ref).abs()
        nbad = int((d > 1e-4 * ref.abs().max()).sum())
        print(f"  {buf:8s} rel {rel(got, ref):.2e}  bad elems {nbad}/{got.numel()}  lo-consistency {((got - hi) - lo).abs().max().item():.1e}")
        if nbad and is_grad:
            idx = torch.nonzero(d > 1e-4 * ref.abs().max())[:6]
            for i in idx: print("      at", tuple(i.tolist()), "got", got[tuple(i)].item(), "ref", ref[tuple(i)].item(), "act", a[tuple(i)].item())
