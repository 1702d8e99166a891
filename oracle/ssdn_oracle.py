"""CPU oracle for the blind-spot denoising hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain, functional restatement (torch CPU ops + numpy) of what the reference
``ssdn`` package computes on the path BASELINE.json names.  It exists so that the CUDA engine can
be checked against it; nothing in the product package imports it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` leg may import it.

Parity status: the reference ships no golden vectors or tests for this path
(``ssdn/tests/test_sampler.py`` covers the sampler only).  The oracle is therefore pinned against
the reference ITSELF, imported on CPU in the build container by ``tests/golden/make_golden.py``;
that script asserts oracle == reference on every case and writes the small fixtures under
``tests/golden/`` that travel to the GPU box.

All tensors are float32, NCHW.  Citations are into /root/reference/ssdn/ssdn/.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1

# --------------------------------------------------------------------------------------------
# Index operations (bit exact)
# --------------------------------------------------------------------------------------------


def rotate(x: torch.Tensor, angle: int) -> torch.Tensor:
    """utils/data.py:42-67.  90: out[i,j] = in[j, W-1-i];  180: out[i,j] = in[H-1-i, W-1-j];
    270: out[i,j] = in[H-1-j, i]  (the last two dims are H, W)."""
    if angle == 0:
        return x
    if angle == 90:
        return torch.flip(x, dims=(-1,)).transpose(-2, -1)
    if angle == 180:
        return torch.flip(x, dims=(-2, -1))
    if angle == 270:
        return torch.flip(x, dims=(-2,)).transpose(-2, -1)
    raise NotImplementedError("rotation must be a multiple of 90 degrees")


def rotate_np(x: np.ndarray, angle: int) -> np.ndarray:
    """Explicit index formulas of :func:`rotate` (independent restatement for the bit-exact tests)."""
    h, w = x.shape[-2:]
    if angle == 0:
        return x.copy()
    if angle == 180:
        return x[..., ::-1, ::-1].copy()
    assert h == w, "90/270 rotation of the stacked batch needs square images"
    out = np.empty_like(x)
    for i in range(h):
        for j in range(w):
            if angle == 90:
                out[..., i, j] = x[..., j, w - 1 - i]
            elif angle == 270:
                out[..., i, j] = x[..., h - 1 - j, i]
            else:
                raise NotImplementedError
    return out


def rot4_stack(x: torch.Tensor) -> torch.Tensor:
    """models/noise_network.py:187-189: the four rotations stacked on the batch axis."""
    return torch.cat([rotate(x, a) for a in (0, 90, 180, 270)], dim=0)


def shift2d(x: torch.Tensor, vert: int, horz: int) -> torch.Tensor:
    """models/utility.py:27-53: out[y,x] = in[y-vert, x-horz], zero where out of range."""
    n, c, h, w = x.shape
    out = torch.zeros_like(x)
    ys0, ys1 = max(0, -vert), min(h, h - vert)
    xs0, xs1 = max(0, -horz), min(w, w - horz)
    if ys1 > ys0 and xs1 > xs0:
        out[:, :, ys0 + vert:ys1 + vert, xs0 + horz:xs1 + horz] = x[:, :, ys0:ys1, xs0:xs1]
    return out


def shift_unrot_concat(x: torch.Tensor) -> torch.Tensor:
    """models/noise_network.py:213-222: shift down one row, split the 4 rotation groups, undo the
    rotations (0, 270, 180, 90) and concatenate on channels."""
    s = shift2d(x, 1, 0)
    parts = torch.chunk(s, 4, dim=0)
    return torch.cat([rotate(p, a) for p, a in zip(parts, (0, 270, 180, 90))], dim=1)


# --------------------------------------------------------------------------------------------
# Layers
# --------------------------------------------------------------------------------------------


def conv2d_same(x, w, b):
    k = w.shape[-1]
    return F.conv2d(x, w, b, stride=1, padding=k // 2)


def shift_conv2d(x, w, b):
    """models/noise_network.py:241-260 (ShiftConv2d): pad k//2 rows on top, 'same' conv, drop the
    bottom k//2 rows  =>  out[h] sees input rows h-2..h for k = 3; identical to a plain conv for k = 1."""
    k = w.shape[-2]
    s = k // 2
    if s == 0:
        return conv2d_same(x, w, b)
    xp = F.pad(x, (0, 0, s, 0))
    y = conv2d_same(xp, w, b)
    return y[:, :, : y.shape[2] - s, :]


def lrelu(x):
    return F.leaky_relu(x, LRELU_SLOPE)


def maxpool2(x, blindspot: bool):
    """models/noise_network.py:64-67: in blind-spot mode the input is shifted down one row (zero row
    on top) before the 2x2 max-pool."""
    if blindspot:
        x = shift2d(x, 1, 0)
    return F.max_pool2d(x, 2)


def upsample2(x):
    return F.interpolate(x, scale_factor=2, mode="nearest")


# Conv layers in construction order: (state-dict prefix, Cin, Cout, k).  models/noise_network.py:70-156.
def layer_table(in_channels: int, out_channels: int, blindspot: bool) -> List[Tuple[str, int, int, int]]:
    t = [("encode_block_1.0", in_channels, 48, 3), ("encode_block_1.2", 48, 48, 3)]
    t += [(f"encode_block_{i}.0", 48, 48, 3) for i in (2, 3, 4, 5, 6)]
    t += [("decode_block_5.0", 96, 96, 3), ("decode_block_5.2", 96, 96, 3)]
    for i in (4, 3, 2):
        t += [(f"decode_block_{i}.0", 144, 96, 3), (f"decode_block_{i}.2", 96, 96, 3)]
    t += [("decode_block_1.0", 96 + in_channels, 96, 3), ("decode_block_1.2", 96, 96, 3)]
    nin = 384 if blindspot else 96
    t += [("output_conv", 96, out_channels, 1), ("output_block.0", nin, nin, 1), ("output_block.2", nin, 96, 1)]
    return t


def param_order(in_channels: int, out_channels: int, blindspot: bool) -> List[str]:
    """Order of ``nn.Module.parameters()`` for the reference network (registration order); this is
    the order Adam sees and the order of the engine's flat parameter buffer."""
    names = []
    for prefix, _, _, _ in layer_table(in_channels, out_channels, blindspot):
        names += [prefix + ".weight", prefix + ".bias"]
    return names


def init_params(in_channels=3, out_channels=3, blindspot=False, zero_output_weights=False,
                generator: Optional[torch.Generator] = None) -> "OrderedDict[str, torch.Tensor]":
    """models/noise_network.py:48-184.  Reproduces the reference's RNG consumption: every nn.Conv2d
    first draws its default init at construction (weight: kaiming_uniform(a=sqrt 5), bias: uniform),
    then init_weights() redraws kaiming_normal(a=0.1) for every conv in modules() order and finally
    redraws (or zeroes) output_conv.  With ``generator=None`` the global torch RNG is used, as in the
    reference, so ``torch.manual_seed(s); init_params(...)`` equals ``torch.manual_seed(s); NoiseNetwork(...)``."""
    table = layer_table(in_channels, out_channels, blindspot)
    p: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for prefix, cin, cout, k in table:  # construction-time draws (discarded later)
        w = torch.empty(cout, cin, k, k)
        torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5), generator=generator)
        bound = 1.0 / math.sqrt(cin * k * k)
        b = torch.empty(cout).uniform_(-bound, bound, generator=generator)
        p[prefix + ".weight"], p[prefix + ".bias"] = w, b
    for prefix, cin, cout, k in table:
        torch.nn.init.kaiming_normal_(p[prefix + ".weight"], a=LRELU_SLOPE, generator=generator)
        p[prefix + ".bias"].zero_()
    if zero_output_weights:
        p["output_conv.weight"].zero_()
    else:
        torch.nn.init.kaiming_normal_(p["output_conv.weight"], nonlinearity="linear", generator=generator)
    return p


def noise_network_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, blindspot: bool) -> torch.Tensor:
    """models/noise_network.py:186-226."""
    conv = shift_conv2d if blindspot else conv2d_same

    def cl(name, t):
        return lrelu(conv(t, p[name + ".weight"], p[name + ".bias"]))

    if blindspot:
        x = rot4_stack(x)
    t = cl("encode_block_1.0", x)
    t = cl("encode_block_1.2", t)
    pools = [maxpool2(t, blindspot)]
    for i in (2, 3, 4, 5):
        pools.append(maxpool2(cl(f"encode_block_{i}.0", pools[-1]), blindspot))
    t = cl("encode_block_6.0", pools[4])
    t = upsample2(t)
    for i, skip in ((5, pools[3]), (4, pools[2]), (3, pools[1]), (2, pools[0])):
        t = torch.cat((t, skip), dim=1)
        t = cl(f"decode_block_{i}.0", t)
        t = cl(f"decode_block_{i}.2", t)
        t = upsample2(t)
    t = torch.cat((t, x), dim=1)
    t = cl("decode_block_1.0", t)
    t = cl("decode_block_1.2", t)
    if blindspot:
        t = shift_unrot_concat(t)
    t = lrelu(conv2d_same(t, p["output_block.0.weight"], p["output_block.0.bias"]))
    t = lrelu(conv2d_same(t, p["output_block.2.weight"], p["output_block.2.bias"]))
    return conv2d_same(t, p["output_conv.weight"], p["output_conv.bias"])


# --------------------------------------------------------------------------------------------
# Pipelines (denoiser.py)
# --------------------------------------------------------------------------------------------


def softplus_sigma(raw: torch.Tensor) -> torch.Tensor:
    """denoiser.py:272-275: softplus(raw - 4) + 1e-3 (beta 1, threshold 20)."""
    return F.softplus(raw - 4.0) + 1e-3


def _inv3(m: torch.Tensor) -> torch.Tensor:
    return torch.inverse(m)


def ssdn_posterior(net_out: torch.Tensor, noisy: torch.Tensor, noise_std: torch.Tensor, sigma_known: bool, diagonal: bool = False):
    """denoiser.py:222-255 and :320-397 for Gaussian noise.

    diagonal (cfg DIAGONAL_COVARIANCE, denoiser.py:213, :236-243): net_out is N x 2C x H x W and Sigma_x = diag(d0^2, d1^2, d2^2).
    The reference's branch stops at :240 with a TypeError (`torch.zeros(c00.shape())`: torch.Size is not callable); this follows
    its evident meaning, and tests/golden/make_golden.py runs the UNMODIFIED reference past that line (with a tensor subclass
    whose .shape is callable) to pin it.

    net_out  N x (C + C(C+1)/2) x H x W   (mean, then the triangular factor A)
    noisy    N x C x H x W
    noise_std N x 1 x 1 x 1 or N x C x 1 x 1  (already max(.,1e-3) / softplus-mapped by the caller)
    Returns dict(mu, pme, loss [N,1], model_std [N,H,W], noise_std [N,1,1]).
    """
    n, c, h, w = noisy.shape
    mu = net_out[:, :c]
    a = net_out[:, c:]
    if c == 1:
        sx = a ** 2
        sn = noise_std ** 2
        sy = sx + sn
        loss = (noisy - mu) ** 2 / sy + torch.log(sy)
        pme = (noisy * sx + mu * sn) / (sx + sn)
        model_std = (sx ** 0.5)[:, 0]
        noise_out = noise_std[:, 0]
        if not sigma_known:
            loss = loss - 0.1 * noise_std
    else:
        assert c == 3
        a = a.permute(0, 2, 3, 1)
        if diagonal:
            zero = torch.zeros_like(a[..., 0])
            a0, a1, a2, a3, a4, a5 = a[..., 0], zero, zero, a[..., 1], zero, a[..., 2]
        else:
            a0, a1, a2, a3, a4, a5 = [a[..., i] for i in range(6)]
        c00 = a0 ** 2 + a1 ** 2 + a2 ** 2
        c01 = a1 * a3 + a2 * a4
        c02 = a2 * a5
        c11 = a3 ** 2 + a4 ** 2
        c12 = a4 * a5
        c22 = a5 ** 2
        sx = torch.stack([torch.stack([c00, c01, c02], -1), torch.stack([c01, c11, c12], -1),
                          torch.stack([c02, c12, c22], -1)], -1)  # N H W 3 3
        eye = torch.eye(3, dtype=noisy.dtype).reshape(1, 1, 1, 3, 3)
        sn = (noise_std ** 2).permute(0, 2, 3, 1)[..., None] * eye
        sy = sx + sn
        sy_inv = _inv3(sy)
        mu2 = mu.permute(0, 2, 3, 1)
        y2 = noisy.permute(0, 2, 3, 1)
        d = y2 - mu2
        quad = torch.sum(d[..., :, None] * d[..., None, :] * sy_inv, dim=(-2, -1))
        dets = torch.clamp_min(torch.det(sy), 0.0)
        loss = 0.5 * torch.log(dets) + 0.5 * quad
        if not sigma_known:
            loss = loss - 0.1 * torch.mean(noise_std, dim=1)
        eps = eye * 1e-6
        sx_inv = _inv3(sx + eps)
        sn_inv = _inv3(sn + eps)
        c1 = _inv3(sx_inv + sn_inv + eps)
        c2 = torch.sum(sx_inv * mu2[..., None, :], -1) + torch.sum(sn_inv * y2[..., None, :], -1)
        pme = torch.sum(c1 * c2[..., None, :], -1).permute(0, 3, 1, 2)
        model_std = torch.clamp_min(torch.det(sx), 0.0) ** (1.0 / 6.0)
        noise_out = torch.clamp_min(torch.det(sn), 0.0) ** (1.0 / 6.0)
    loss = loss.reshape(n, -1).mean(1, keepdim=True)
    return {"mu": mu, "pme": pme, "loss": loss, "model_std": model_std, "noise_std": noise_out}


def ssdn_pipeline(params: Dict[str, torch.Tensor], noisy: torch.Tensor, noise_values: torch.Tensor,
                  sigma_mode: str, est_params: Optional[Dict[str, torch.Tensor]] = None,
                  est_sigma: Optional[torch.Tensor] = None, noise_style: str = "gauss", diagonal: bool = False):
    """denoiser.py:182-397.  sigma_mode in {"known", "const", "var"}; noise_style "gauss..." or "poisson..." (the
    signal-dependent approximation of :285-297: sigma = sqrt(max(mu, 1e-3) / lambda) per pixel and channel when lambda is
    known, sqrt(max(mu, 1e-3) * estimate) otherwise)."""
    c = noisy.shape[1]
    net_out = noise_network_forward(params, noisy, blindspot=True)
    if sigma_mode == "known":
        est = None
    elif sigma_mode == "const":
        est = softplus_sigma(est_sigma)
    elif sigma_mode == "var":
        e = noise_network_forward(est_params, noisy, blindspot=False)
        est = softplus_sigma(torch.mean(e, dim=(2, 3), keepdim=True))
    else:
        raise NotImplementedError(sigma_mode)
    if noise_style.startswith("gauss"):
        noise_std = torch.max(noise_values, torch.tensor(1e-3, dtype=noise_values.dtype)) if est is None else est
    elif noise_style.startswith("poisson"):
        base = torch.max(net_out[:, :c], torch.tensor(1e-3, dtype=net_out.dtype))
        noise_std = (base / noise_values) ** 0.5 if est is None else (base * est) ** 0.5
    else:
        raise NotImplementedError(noise_style)
    out = ssdn_posterior(net_out, noisy, noise_std, sigma_known=(sigma_mode == "known"), diagonal=diagonal and c == 3)
    out["net_out"] = net_out
    return out


def mse_pipeline(params, inp, ref, blindspot=False):
    """denoiser.py:140-157: per-sample mean squared error."""
    cleaned = noise_network_forward(params, inp, blindspot)
    loss = ((cleaned - ref) ** 2).reshape(inp.shape[0], -1).mean(1, keepdim=True)
    return {"out": cleaned, "loss": loss}


def masked_mse(coords: torch.Tensor, cleaned: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    """utils/n2v_loss.py:6-17 + denoiser.py:176-178: the FIRST sample's coordinate list is used for
    the whole batch and indexes [:, :, x, y] (row = first coordinate); squared errors are summed over
    the coordinates, then averaged over channels -> N x 1."""
    lst = coords[0].tolist()
    acc = torch.zeros(cleaned.shape[0], cleaned.shape[1])
    for (x, y) in lst:
        d = ref[:, :, x, y] - cleaned[:, :, x, y]
        acc = acc + d * d
    return acc.mean(1, keepdim=True)


def mask_mse_pipeline(params, inp, ref, coords):
    cleaned = noise_network_forward(params, inp, blindspot=False)
    return {"out": cleaned, "loss": masked_mse(coords, cleaned, ref)}


def psnr(img: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    """utils/data.py:94-105: per-sample -10 log10(mean_{C,H,W} (img-ref)^2) for float images."""
    mse = ((img - ref) ** 2).mean(dim=(1, 2, 3))
    return -10.0 * torch.log10(mse)


# --------------------------------------------------------------------------------------------
# Optimiser side (train.py:100-107,197-202,274-282; utils/utils.py:18-37)
# --------------------------------------------------------------------------------------------


def compute_ramped_lrate(i, iteration_count, ramp_up_fraction, ramp_down_fraction, learning_rate):
    if ramp_up_fraction > 0.0:
        if i <= iteration_count * ramp_up_fraction:
            t = (i / ramp_up_fraction) / iteration_count
            learning_rate = learning_rate * (0.5 - np.cos(t * np.pi) / 2)
    if ramp_down_fraction > 0.0:
        start = iteration_count * (1 - ramp_down_fraction)
        if i >= start:
            t = ((i - start) / ramp_down_fraction) / iteration_count
            learning_rate = learning_rate * (0.5 + np.cos(t * np.pi) / 2) ** 2
    return learning_rate


def effective_lrate(iteration, cfg_iterations, cfg_rampup=0.3, cfg_rampdown=0.1, lr=3e-4):
    """train.py:274-282 passes (RAMPDOWN, RAMPUP) into (ramp_up, ramp_down): with the defaults of
    cfg.py:18-19 the effective schedule ramps UP over the first 10 % and DOWN over the last 30 %."""
    return compute_ramped_lrate(iteration, cfg_iterations, cfg_rampdown, cfg_rampup, lr)


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.99, eps=1e-8):
    """torch.optim.Adam (no weight decay, no amsgrad) single-tensor update, in place; ``step`` is
    the 1-based step count after the increment."""
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


# --------------------------------------------------------------------------------------------
# Synthetic workload and a complete CPU training step (the CPU baseline)
# --------------------------------------------------------------------------------------------


def synthetic_batch(n, c, size, seed=1234, sigma=25.0 / 255.0, clip=True):
    """Smooth random clean images + Gaussian noise as utils/noise.py:57-61 (gauss25, clipped)."""
    g = torch.Generator().manual_seed(seed)
    clean = F.interpolate(torch.rand(n, c, 8, 8, generator=g), size=(size, size), mode="bilinear",
                          align_corners=False)
    noisy = clean + torch.randn(n, c, size, size, generator=g) * sigma
    if clip:
        noisy = noisy.clamp(0.0, 1.0)
    return clean, noisy


class CpuTrainer:
    """zero_grad -> pipeline -> mean(loss).backward() -> Adam.step() exactly as train.py:197-202, on
    the oracle's functional network (autograd supplies the backward, as it does in the reference)."""

    def __init__(self, algorithm="ssdn", sigma_mode="known", channels=3, seed=0, lr=3e-4):
        torch.manual_seed(seed)
        self.algorithm, self.sigma_mode, self.c = algorithm, sigma_mode, channels
        if algorithm == "ssdn":
            out_ch = channels + channels * (channels + 1) // 2
            self.params = init_params(channels, out_ch, blindspot=True)
        else:
            self.params = init_params(channels, channels, blindspot=False)
        self.est_params = None
        self.est_sigma = None
        if algorithm == "ssdn" and sigma_mode == "var":
            self.est_params = init_params(channels, 1, blindspot=False, zero_output_weights=True)
        if algorithm == "ssdn" and sigma_mode == "const":
            self.est_sigma = torch.zeros(1, 1, 1, 1)
        self.leaves = list(self.params.values())
        if self.est_params is not None:
            self.leaves += list(self.est_params.values())
        if self.est_sigma is not None:
            self.leaves.append(self.est_sigma)
        for t in self.leaves:
            t.requires_grad_(True)
        self.m = [torch.zeros_like(t) for t in self.leaves]
        self.v = [torch.zeros_like(t) for t in self.leaves]
        self.step_count = 0
        self.lr = lr

    def loss(self, noisy, noise_values=None, ref=None, coords=None):
        if self.algorithm == "ssdn":
            return ssdn_pipeline(self.params, noisy, noise_values, self.sigma_mode, self.est_params, self.est_sigma)
        if self.algorithm == "n2v":
            return mask_mse_pipeline(self.params, noisy, ref, coords)
        return mse_pipeline(self.params, noisy, ref)

    def step(self, noisy, noise_values=None, ref=None, coords=None, lr=None):
        for t in self.leaves:
            t.grad = None
        out = self.loss(noisy, noise_values, ref, coords)
        out["loss"].mean().backward()
        self.step_count += 1
        with torch.no_grad():
            for t, m, v in zip(self.leaves, self.m, self.v):
                adam_step(t, t.grad, m, v, self.step_count, self.lr if lr is None else lr)
        return out
