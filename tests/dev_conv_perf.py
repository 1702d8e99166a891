import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "selfsupervised-denoising_b200"))
import torch
from ssdn import _engine as E
torch.manual_seed(0)
for (n, cin, h, w, cout, k, blind) in [(128, 96, 64, 64, 96, 3, True), (128, 48, 64, 64, 48, 3, True), (128, 144, 32, 32, 96, 3, True), (32, 384, 64, 64, 384, 1, False)]:
    x = torch.randn(n, cin, h, w, device="cuda"); wt = torch.randn(cout, cin, k, k, device="cuda") / (cin * k * k) ** 0.5; b = torch.randn(cout, device="cuda")
    for _ in range(2):
        y = E.conv2d_forward(x, wt, b, blind=blind, lrelu=True)
    torch.cuda.synchronize()
    print(n, cin, h, w, cout, k, "GFLOP", 2 * n * h * w * cin * cout * k * k / 1e9)
