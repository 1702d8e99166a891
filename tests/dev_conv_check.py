"""Developer check (run on the GPU box): conv forward / data-gradient vs the CPU oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "selfsupervised-denoising_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import ssdn_oracle as O
from ssdn import _engine as E

def rel(a, b): return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()

torch.manual_seed(0)
cases = [(2, 48, 16, 16, 48, 3, True), (2, 96, 32, 32, 96, 3, True), (1, 3, 64, 64, 48, 3, True), (2, 144, 16, 16, 96, 3, False),
         (2, 99, 32, 32, 96, 3, True), (4, 48, 2, 2, 48, 3, True), (2, 384, 32, 32, 384, 1, False), (2, 96, 32, 32, 9, 1, False),
         (8, 96, 64, 64, 96, 3, True), (3, 1, 32, 32, 48, 3, False)]
bad = 0
for (n, cin, h, w, cout, k, blind) in cases:
    x = torch.randn(n, cin, h, w); wt = torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5; b = torch.randn(cout)
    ref = O.lrelu((O.shift_conv2d if blind else O.conv2d_same)(x.double(), wt.double(), b.double())).float()
    t0 = time.time()
    y = E.conv2d_forward(x.cuda(), wt.cuda(), b.cuda(), blind=blind, lrelu=True).cpu()
    e = rel(y, ref)
    # data gradient: autograd of the oracle
    xg = x.double().requires_grad_(True)
    yo = (O.shift_conv2d if blind else O.conv2d_same)(xg, wt.double(), None)
    dy = torch.randn(n, cout, h, w)
    yo.backward(dy.double())
    dx = E.conv2d_backward_data(dy.cuda(), wt.cuda(), blind=blind).cpu()
    e2 = rel(dx, xg.grad.float())
    wg = wt.double().requires_grad_(True); bg = b.double().requires_grad_(True)
    (O.shift_conv2d if blind else O.conv2d_same)(x.double(), wg, bg).backward(dy.double())
    dw, db = E.conv2d_backward_weight(x.cuda(), dy.cuda(), k, blind=blind)
    e3 = rel(dw.cpu(), wg.grad.float()); e4 = rel(db.cpu(), bg.grad.float())
    ok = e < 2e-5 and e2 < 2e-5 and e3 < 2e-5 and e4 < 2e-5
    bad += (not ok)
    print(f"n{n} cin{cin} {h}x{w} cout{cout} k{k} blind{int(blind)}: fwd rel {e:.2e}  dgrad rel {e2:.2e}  wgrad rel {e3:.2e}  bgrad rel {e4:.2e}  {'OK' if ok else 'FAIL'}  ({time.time()-t0:.2f}s)", flush=True)
print("FAILED" if bad else "ALL OK")
sys.exit(1 if bad else 0)
