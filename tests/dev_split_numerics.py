"""Dev tool (CPU, not a test): operand-split numerics for the three GEMMs of a layer, on REAL tensors of the path.

    python tests/dev_split_numerics.py > profiles/r01_split_numerics.txt

Question: can the 3xTF32 operand split (hi/lo tf32 planes, products hi*hi + lo*hi + hi*lo on
kind::tf32, K = 8 per instruction) be replaced by a two-term fp16 split on kind::f16 (K = 16 per instruction at the same
issue rate, half the operand bytes) without losing accuracy?  tf32 and fp16 carry the same 11 significant bits; what
differs is the exponent range (8 vs 5 bits), so the answer depends on the dynamic range of the actual activations,
weights and gradients, and on how precisely a per-tensor power-of-two scale has to be chosen.

Tensors: the trained reference checkpoint of tests/golden (weights as shipped) on a seeded synthetic batch, traced
through the oracle in fp64 (activations X, pre-activation gradients dZ of the NLL loss).  Each scheme's products are
accumulated in fp64 here, so the numbers isolate the OPERAND error (accumulation order / accumulator truncation are
separate effects, measured on the GPU in profiles/r01_conv_accuracy.log)."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
sys.path.insert(0, HERE)
import ssdn_oracle as O  # noqa: E402
from oracle_trace import oracle_trace  # noqa: E402

torch.set_num_threads(8)


def rn_tf32(x: torch.Tensor) -> torch.Tensor:
    """Round an fp32-representable fp64 tensor to tf32 (10 explicit mantissa bits), nearest-even."""
    i = x.float().contiguous().view(torch.int32)
    r = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
    return r.view(torch.float32).double()


def split_tf32(x):
    hi = rn_tf32(x)
    return hi, rn_tf32(x - hi)


def split_f16(x, scale):
    xs = x * scale
    hi = xs.float().half()
    lo = (xs - hi.double()).float().half()
    return hi.double() / scale, lo.double() / scale


def pow2_scale(x, target_log2):
    m = float(x.abs().max())
    return 2.0 ** (target_log2 - int(np.ceil(np.log2(m)))) if m > 0 else 1.0


def three_products(op, a_hi, a_lo, b_hi, b_lo):
    return op(a_hi, b_hi) + op(a_lo, b_hi) + op(a_hi, b_lo)


def rel(a, b):
    return float((a - b).norm() / b.norm())


def study(name, x, w, dz, blind3x3):
    """x: layer input [B,Cin,H,W]; w: [Cout,Cin,k,k]; dz: gradient of the loss w.r.t. the layer's pre-activation."""
    k = w.shape[-1]
    if k == 3 and blind3x3:                       # shifted conv: rows {h-2, h-1, h} (noise_network.py:256-260)
        pad = lambda t: F.pad(t, (0, 0, 1, 0))[:, :, :-1]          # noqa: E731
        fwd = lambda a, b: F.conv2d(pad(a), b, padding=1)          # noqa: E731
    else:
        fwd = lambda a, b: F.conv2d(a, b, padding=k // 2)          # noqa: E731
    # the same three contractions the engine runs, written with autograd-free torch ops in fp64
    def dgrad(g, b):
        xx = torch.zeros_like(x).requires_grad_(True)
        return torch.autograd.grad(fwd(xx, b), xx, g)[0]

    def wgrad(g, a):
        ww = torch.zeros_like(w).requires_grad_(True)
        return torch.autograd.grad(fwd(a, ww), ww, g)[0]

    exact = {"fwd": fwd(x, w), "dgrad": dgrad(dz, w), "wgrad": wgrad(dz, x)}
    rows = []
    stats = {n: (float(t.abs().max()), float(t.abs()[t != 0].min()) if (t != 0).any() else 0.0) for n, t in (("X", x), ("W", w), ("dZ", dz))}

    def run(label, sx, sw, sdz):
        xs, ws, gs = sx(x), sw(w), sdz(dz)
        out = {"fwd": three_products(fwd, *xs, *ws), "dgrad": three_products(dgrad, *gs, *ws), "wgrad": three_products(wgrad, *gs, *xs)}
        rows.append((label, [rel(out[n], exact[n]) for n in ("fwd", "dgrad", "wgrad")]))

    run("3xTF32 (today)", split_tf32, split_tf32, split_tf32)
    run("fp16x2, no scaling", lambda t: split_f16(t, 1.0), lambda t: split_f16(t, 1.0), lambda t: split_f16(t, 1.0))
    for target in (15, 14, 8, 6, 4, 0, -6, -10):
        s = lambda t, target=target: split_f16(t, pow2_scale(t, target))          # noqa: E731
        run(f"fp16x2, per-tensor 2^k scale, max -> 2^{target}", s, s, s)
    # one plane only (what a single-pass half-precision GEMM would give), for scale
    run("fp16 single plane (scaled to 2^8)", *([lambda t: (split_f16(t, pow2_scale(t, 8))[0], torch.zeros_like(t))] * 3))
    print(f"\n{name}: X {tuple(x.shape)} W {tuple(w.shape)}   |max| / smallest non-zero |.|:  "
          + "   ".join(f"{n} {a:.2e} / {b:.1e} ({np.log2(a / b) if b > 0 else 0:.0f} binades)" for n, (a, b) in stats.items()))
    print(f"  {'scheme':<52} {'fwd':>10} {'dgrad':>10} {'wgrad':>10}   (relative L2 error of the operand split, fp64 accumulation)")
    for label, errs in rows:
        print(f"  {label:<52} " + " ".join(f"{e:10.2e}" for e in errs))


def main():
    z = np.load(os.path.join(HERE, "golden", "wt_ssdn_gauss25_sigma_known.npz"))
    p = {k[2:]: torch.from_numpy(z[k]).double().requires_grad_(True) for k in z.files if k.startswith("p.")}
    clean, noisy = O.synthetic_batch(2, 3, 64, seed=4242)
    sigma = torch.full((2, 1, 1, 1), 25.0 / 255.0, dtype=torch.float64)
    out, T = oracle_trace(p, noisy.double(), True)
    post = O.ssdn_posterior(out, noisy.double(), sigma, True)
    loss = post["loss"] if isinstance(post, dict) else post[1]
    loss.mean().backward()
    print("operand-split numerics on the trained gauss25 sigma-known checkpoint, batch 2 x 3 x 64 x 64 (8 rotated images), loss",
          float(loss.mean()))
    inputs = {"encode_block_1.2": T["encode_block_1.0"][1], "decode_block_1.0": T["cat1"][1], "decode_block_1.2": T["decode_block_1.0"][1],
              "decode_block_3.0": T["cat3"][1], "encode_block_6.0": T["pools"][1][4], "output_block.0": T["head_in"][1],
              "output_block.2": T["output_block.0"][1]}
    for name, x in inputs.items():
        zt = T[name][0]
        study(name, x.detach(), p[name + ".weight"].detach(), zt.grad.detach(), blind3x3=not name.startswith("output_block"))


if __name__ == "__main__":
    main()
