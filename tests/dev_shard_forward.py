"""Developer tool (one GPU): forward outputs of a batch against the same samples run as two half batches, pass by pass
after plan creation (the operand scales settle during the first passes).  SSDN_LIB=<path>: another build of the engine."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "selfsupervised-denoising_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
from ssdn import _engine as E
if os.environ.get("SSDN_LIB"):
    E.LIB_PATH = os.environ["SSDN_LIB"]
import ssdn_oracle as O
torch.manual_seed(0)
p = O.init_params(3, 9, True)
flat = torch.cat([p[k].reshape(-1) for k in O.param_order(3, 9, True)]).cuda()
n, size = 8, 32
_, noisy = O.synthetic_batch(n, 3, size, seed=1234)
full = E.NetPlan(n, 3, 9, size, size, True, "cuda")
half = E.NetPlan(n // 2, 3, 9, size, size, True, "cuda")
for it in range(4):
    f = full.forward(flat, noisy.cuda(), training=True).clone()
    lo = half.forward(flat, noisy[:n // 2].cuda(), training=True).clone()
    hi = half.forward(flat, noisy[n // 2:].cuda(), training=True).clone()
    h = torch.cat([lo, hi])
    d = (f - h).abs()
    print(f"pass {it}: differing outputs {float((d > 0).float().mean()):.3f}, max |diff| {float(d.max()):.2e} (range {float(f.abs().max()):.2e}); scales full {full.debug_scales()[0][:24]}\n        half {half.debug_scales()[0][:24]}")
