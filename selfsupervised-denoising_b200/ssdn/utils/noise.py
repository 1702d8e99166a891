"""Synthetic noise for training inputs (reference: ssdn/ssdn/utils/noise.py).  Runs in the data
pipeline on CPU tensors, outside the engine's hot path; kept for drop-in compatibility.

Style strings: 'gauss25', 'gauss5_50' (std drawn uniformly per leading-axis element), 'poisson30',
'poisson5_50', optional '_nc' suffix = do not clip to the image range.  Integer parameters of the
Gaussian styles are in 8-bit units (divided by 255)."""
import re
from numbers import Number
from typing import Tuple, Union

import torch
from torch import Tensor

from ssdn.utils.data import clip_img


def _range_sample(lo: float, hi: float, like: Tensor) -> Tensor:
    shape = [like.shape[0]] + [1] * (like.dim() - 1)
    return torch.distributions.Uniform(lo, hi).sample(shape)


def add_gaussian(tensor: Tensor, std_dev, mean: Number = 0, inplace: bool = False, clip: bool = True):
    out = tensor if inplace else tensor.clone()
    if isinstance(std_dev, (list, tuple)):
        if len(std_dev) == 1:
            std_dev = std_dev[0]
        else:
            lo, hi = std_dev
            lo = lo / 255 if isinstance(lo, int) else lo
            hi = hi / 255 if isinstance(hi, int) else hi
            std_dev = _range_sample(lo, hi, out)
    if isinstance(std_dev, int):
        std_dev = std_dev / 255
    out.add_(torch.randn(out.size()) * std_dev + mean)
    if clip:
        out = clip_img(out, inplace=True)
    return out, std_dev


def add_poisson(tensor: Tensor, lam, inplace: bool = False, clip: bool = True):
    out = tensor if inplace else tensor.clone()
    if isinstance(lam, (list, tuple)):
        lam = lam[0] if len(lam) == 1 else _range_sample(lam[0], lam[1], out)
    out.mul_(lam)
    out.add_(torch.distributions.Poisson(torch.tensor(1, dtype=float)).sample(out.shape))
    out.div_(lam)
    if clip:
        out = clip_img(out, inplace=True)
    return out, lam


def add_style(images: Tensor, style: str, inplace: bool = False) -> Tuple[Tensor, Union[Number, Tensor]]:
    kind = re.findall(r"[a-zA-Z]+", style)[0]
    tokens = [t for t in style.replace(kind, "").split("_") if t != ""]
    clip = "nc" not in tokens
    tokens = [t for t in tokens if t != "nc"]
    as_float = any("." in t for t in tokens)
    values = [float(t) if as_float else int(t) for t in tokens]
    if kind == "gauss":
        return add_gaussian(images, values, inplace=inplace, clip=clip)
    if kind == "poisson":
        return add_poisson(images, values, inplace=inplace, clip=clip)
    raise NotImplementedError("Noise type not supported")
