"""Seeded test cases shared by make_golden.py (reference side, build container) and the parity
tests (engine side, GPU box).  Inputs and weights are regenerated from seeds with the torch CPU
generator, which is deterministic across machines; only reference OUTPUTS are stored as fixtures."""
from __future__ import annotations

import os
import sys
from collections import OrderedDict

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.path.join(_ROOT, "oracle") not in sys.path:
    sys.path.insert(0, os.path.join(_ROOT, "oracle"))
import ssdn_oracle as O  # noqa: E402

GOLDEN_DIR = os.path.dirname(os.path.abspath(__file__))

# name -> (in_channels, out_channels, blindspot, N, size)
NETWORK_CASES = OrderedDict(
    net_blind_rgb=(3, 9, True, 2, 32),
    net_blind_mono=(1, 2, True, 1, 32),
    net_plain_rgb=(3, 3, False, 2, 32),
    net_plain_mono=(1, 1, False, 2, 32),
    net_blind_rgb_64=(3, 9, True, 1, 64),
)

# name -> (algorithm, sigma_mode, channels, N, size)
PIPELINE_CASES = OrderedDict(
    ssdn_known_rgb=("ssdn", "known", 3, 2, 32),
    ssdn_const_rgb=("ssdn", "const", 3, 2, 32),
    ssdn_var_rgb=("ssdn", "var", 3, 2, 32),
    ssdn_var_rgb_perchannel=("ssdn", "var", 3, 2, 32),
    ssdn_known_mono=("ssdn", "known", 1, 2, 32),
    ssdn_var_mono=("ssdn", "var", 1, 2, 32),
    n2c_mono=("n2c", None, 1, 4, 32),
    n2v_rgb=("n2v", None, 3, 2, 32),
    # appended (seeds are 200 + position): Poisson noise, denoiser.py:285-297
    ssdn_known_rgb_poisson=("ssdn", "known", 3, 2, 32),
    ssdn_const_rgb_poisson=("ssdn", "const", 3, 2, 32),
    ssdn_const_mono_poisson=("ssdn", "const", 1, 2, 32),
    # appended: diagonal covariance (cfg DIAGONAL_COVARIANCE, denoiser.py:213, :236-243), 3 + 3 network outputs
    ssdn_known_rgb_diag=("ssdn", "known", 3, 2, 32),
    ssdn_const_rgb_diag=("ssdn", "const", 3, 2, 32),
)


def make_params(in_ch, out_ch, blindspot, seed, zero_output_weights=False):
    """Reference-style init, then small random biases so the bias path is exercised."""
    g = torch.Generator().manual_seed(seed)
    p = O.init_params(in_ch, out_ch, blindspot, zero_output_weights, generator=g)
    for k in p:
        if k.endswith(".bias"):
            p[k] = 0.05 * torch.randn(p[k].shape, generator=g)
    if zero_output_weights:  # give the estimator a non-trivial head so gradients flow in the test
        p["output_conv.weight"] = 0.05 * torch.randn(p["output_conv.weight"].shape, generator=g)
    return p


def network_inputs(name):
    cin, cout, blind, n, size = NETWORK_CASES[name]
    seed = 100 + list(NETWORK_CASES).index(name)
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, cin, size, size, generator=g)
    dout = torch.randn(n, cout, size, size, generator=g)
    return make_params(cin, cout, blind, seed), x, dout


def pipeline_inputs(name):
    algo, mode, c, n, size = PIPELINE_CASES[name]
    seed = 200 + list(PIPELINE_CASES).index(name)
    g = torch.Generator().manual_seed(seed)
    clean, noisy = O.synthetic_batch(n, c, size, seed=seed)
    d = dict(algorithm=algo, sigma_mode=mode, channels=c, clean=clean, noisy=noisy)
    if algo == "ssdn":
        d["diagonal"] = name.endswith("diag")
        d["params"] = make_params(c, 2 * c if d["diagonal"] else c + c * (c + 1) // 2, True, seed)
        if name.endswith("poisson"):
            d["noise_style"] = "poisson30"
            d["noise_values"] = torch.full((n, 1, 1, 1), 30.0) if mode == "known" else torch.full((n, 1, 1, 1), 25.0 / 255.0)
        elif name.endswith("perchannel"):
            d["noise_values"] = (torch.rand(n, c, 1, 1, generator=g) * 45 + 5) / 255.0
        else:
            d["noise_values"] = torch.full((n, 1, 1, 1), 25.0 / 255.0)
        if mode == "var":
            d["est_params"] = make_params(c, 1, False, seed + 1, zero_output_weights=True)
        if mode == "const":
            d["est_sigma"] = torch.full((1, 1, 1, 1), 1.25)
    else:
        d["params"] = make_params(c, c, False, seed)
        if algo == "n2v":
            d["ref"] = (clean + torch.randn(clean.shape, generator=g) * 25 / 255).clamp(0, 1)
            d["coords"] = torch.randint(0, size, (n, 64, 2), generator=g)
        else:
            d["ref"] = clean
    return d


def grad_summary(named_grads):
    """Compact, order-stable description of a gradient set: per tensor (sum, l2, first 4 values)."""
    rows = []
    for k, g in named_grads.items():
        f = g.detach().reshape(-1).double()
        head = torch.zeros(4, dtype=torch.float64)
        head[: min(4, f.numel())] = f[:4]
        rows.append(torch.cat([f.sum()[None], f.norm()[None], head]))
    return torch.stack(rows).float()


def load_golden(name):
    import numpy as np
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    z = np.load(path)
    return {k: torch.from_numpy(z[k]) for k in z.files}
