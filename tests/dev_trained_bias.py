"""Developer tool (GPU box): signed bias / rms error of every convolution of the TRAINED checkpoint, layer by layer, each fed
with the exact (fp64 oracle) input of that layer - the data the accumulator-truncation compensation (SSDN_ACC_BETA) has to
be right for.  Run with SSDN_ACC_COMP=0 to see the uncompensated bias per MMA instruction."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("selfsupervised-denoising_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, torch
import ssdn_oracle as O
from oracle_trace import oracle_trace
from ssdn import _engine as E
z = np.load(os.path.join(ROOT, "tests", "golden", "wt_ssdn_gauss25_sigma_known.npz"))
params = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p.")}
noisy = torch.from_numpy(z["noisy"])
_, T = oracle_trace({k: v.double().requires_grad_(True) for k, v in params.items()}, noisy.double(), True)
inputs = {"encode_block_1.0": T["x"][1], "encode_block_1.2": T["encode_block_1.0"][1], "encode_block_2.0": T["pools"][1][0],
          "encode_block_3.0": T["pools"][1][1], "encode_block_4.0": T["pools"][1][2], "encode_block_5.0": T["pools"][1][3], "encode_block_6.0": T["pools"][1][4],
          "decode_block_5.0": T["cat5"][1], "decode_block_5.2": T["decode_block_5.0"][1], "decode_block_4.0": T["cat4"][1], "decode_block_4.2": T["decode_block_4.0"][1],
          "decode_block_3.0": T["cat3"][1], "decode_block_3.2": T["decode_block_3.0"][1], "decode_block_2.0": T["cat2"][1], "decode_block_2.2": T["decode_block_2.0"][1],
          "decode_block_1.0": T["cat1"][1], "decode_block_1.2": T["decode_block_1.0"][1], "output_block.0": T["head_in"][1],
          "output_block.2": T["output_block.0"][1], "output_conv": T["output_block.2"][1]}
tot = 0.0
for name, x64 in inputs.items():
    x = x64.detach().float()
    w, b = params[name + ".weight"], params[name + ".bias"]
    k = w.shape[-1]
    blind = k == 3
    ref = (O.shift_conv2d if blind else O.conv2d_same)(x.double(), w.double(), b.double())
    got = E.conv2d_forward(x.cuda(), w.cuda(), b.cuda(), blind=blind, lrelu=False).double().cpu()
    err = got - ref
    n_mma = (w.shape[1] + 15) // 16 * k * k * 3
    bias = ((err * torch.sign(ref)).mean() / ref.abs().mean()).item()
    rms = (err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    tot += bias
    print(f"{name:18s} cin {w.shape[1]:3d} n_mma {n_mma:4d}  bias {bias:+.2e} ({bias / n_mma:+.2e} per MMA)  rms {rms:.2e}")
print(f"sum of the signed biases over the 20 layers: {tot:+.2e}")
