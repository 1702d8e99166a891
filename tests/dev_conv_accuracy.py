"""Developer tool (GPU box): accuracy of one split-operand (fp16 hi/lo planes, 3 MMAs per product) convolution against fp64.  The tensor core's fp32 accumulator
TRUNCATES: every accumulate step shrinks the magnitude by about half an ulp, so the error is a bias proportional to the
number of MMA instructions per output, not noise.  Prints the error before and after the first-order compensation
y * (1 + beta * n_mma) for several input distributions (beta is fitted on the first case only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "selfsupervised-denoising_b200"))
import torch
import torch.nn.functional as F
from ssdn import _engine as E
torch.manual_seed(0)
RAW = os.environ.get("SSDN_ACC_COMP") == "0"     # engine built/run without the in-kernel compensation: fit beta here
beta = None
def dist(name, shape):
    if name == "uniform+": return torch.rand(shape)
    if name == "normal": return torch.randn(shape)
    if name == "lrelu(normal)": return F.leaky_relu(torch.randn(shape), 0.1)
    if name == "sparse+": return torch.rand(shape) * (torch.rand(shape) < 0.2)
for cin, cout, k in ((96, 96, 3), (384, 384, 1), (48, 48, 3), (144, 96, 3)):
    n_mma = (cin + 15) // 16 * k * k * 3            # k-steps of 16 channels x taps x 3 products
    for xd, wd in (("uniform+", "normal"), ("normal", "normal"), ("lrelu(normal)", "normal"), ("sparse+", "normal"), ("uniform+", "uniform+")):
        x = dist(xd, (4, cin, 32, 32)); w = dist(wd, (cout, cin, k, k)) / (cin * k * k) ** 0.5
        y64 = F.conv2d(x.double(), w.double(), padding=k // 2)
        ye = E.conv2d_forward(x.cuda(), w.cuda(), None, blind=False, lrelu=False).double().cpu()
        err = ye - y64
        rms = lambda e: (e.pow(2).mean().sqrt() / y64.pow(2).mean().sqrt()).item()
        bias = ((err * torch.sign(y64)).mean() / y64.abs().mean()).item()
        msg = f"{cin:3d}->{cout:3d} k{k} n_mma {n_mma:4d} x~{xd:14s} w~{wd:8s} rms err {rms(err):.2e}  bias {bias:+.2e} ({bias / n_mma:+.2e} per MMA)"
        if RAW:
            if beta is None: beta = -bias / n_mma
            msg += f"  | host-side compensation with beta {beta:.3e}: rms err {rms(ye * (1 + beta * n_mma) - y64):.2e}"
        print(msg)
