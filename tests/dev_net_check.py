"""Developer check (GPU box): whole-network forward/backward vs golden fixtures and the live oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "selfsupervised-denoising_b200")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import cases as C
import ssdn_oracle as O
from ssdn import _engine as E

def rel(a, b): return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()
bad = 0
for name, (cin, cout, blind, n, size) in C.NETWORK_CASES.items():
    params, x, dout = C.network_inputs(name)
    gold = C.load_golden(name)
    order = O.param_order(cin, cout, blind)
    flat = torch.cat([params[k].reshape(-1) for k in order]).cuda()
    plan = E.NetPlan(n, cin, cout, size, size, blind, "cuda")
    out = plan.forward(flat, x.cuda(), training=True)
    plan.check()
    e_out = rel(out.cpu(), gold["out"])
    grads = plan.backward(flat, dout.cuda())
    plan.check()
    # live oracle gradients
    po = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    O.noise_network_forward(po, x, blind).backward(dout)
    off = 0; worst = 0.0; worst_k = ""
    gs = {}
    for k in order:
        nel = params[k].numel()
        g = grads[off:off + nel].cpu().reshape(params[k].shape); off += nel
        gs[k] = g
        e = rel(g, po[k].grad)
        if e > worst: worst, worst_k = e, k
    e_sum = rel(C.grad_summary(gs), gold["grad_summary"])
    ok = e_out < 1e-4 and worst < 1e-4
    bad += (not ok)
    print(f"{name}: out rel {e_out:.2e}  worst grad rel {worst:.2e} ({worst_k})  summary rel {e_sum:.2e}  {'OK' if ok else 'FAIL'}", flush=True)
    if not ok:
        off = 0
        for k in order:
            print("   ", k, f"{rel(gs[k], po[k].grad):.2e}")
print("FAILED" if bad else "ALL OK")
