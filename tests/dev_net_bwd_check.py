"""Developer check (GPU box): backward pass run on the ORACLE's forward activations (identical LeakyReLU masks)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "selfsupervised-denoising_b200")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import cases as C
import ssdn_oracle as O
from ssdn import _engine as E
from oracle_trace import oracle_trace, upload_activations

def rel(a, b): return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()
bad = 0
for name, (cin, cout, blind, n, size) in C.NETWORK_CASES.items():
    params, x, dout = C.network_inputs(name)
    order = O.param_order(cin, cout, blind)
    flat = torch.cat([params[k].reshape(-1) for k in order]).cuda()
    plan = E.NetPlan(n, cin, cout, size, size, blind, "cuda")
    plan.forward(flat, x.cuda(), training=True)
    po = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    oo, T = oracle_trace(po, x, blind)
    oo.backward(dout)
    upload_activations(plan, T)
    grads = plan.backward(flat, dout.cuda()); plan.check()
    off = 0; worst = 0.0; worst_k = ""
    for k in order:
        nel = params[k].numel(); g = grads[off:off + nel].cpu().reshape(params[k].shape); off += nel
        e = rel(g, po[k].grad)
        if e > worst: worst, worst_k = e, k
    ok = worst < 2e-5; bad += (not ok)
    print(f"{name}: worst grad rel {worst:.2e} ({worst_k}) {'OK' if ok else 'FAIL'}", flush=True)
print("FAILED" if bad else "ALL OK")
