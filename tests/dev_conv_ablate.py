import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "selfsupervised-denoising_b200"))
import torch
from ssdn import _engine as E
torch.manual_seed(0)
n, cin, h, w, cout, k = 128, 96, 64, 64, 96, 3
x = torch.randn(n, cin, h, w, device="cuda"); wt = torch.randn(cout, cin, k, k, device="cuda") / 30; b = torch.randn(cout, device="cuda")
for _ in range(2): y = E.conv2d_forward(x, wt, b, blind=True, lrelu=True)
E.profile_begin(); y = E.conv2d_forward(x, wt, b, blind=True, lrelu=True); E.profile_end()
print("debug", os.environ.get("SSDN_CONV_DEBUG"), "kernel us", [round(r[1] * 1e3, 1) for r in E.profile_records()], "checksum", float(y.abs().sum()))
