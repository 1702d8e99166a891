"""Noise2Void uniform pixel selection (reference: ssdn/ssdn/utils/n2v_ups.py; stratified sampling after
juglab/n2v).  CPU data-pipeline code.  The reference's neighbourhood bounds use ``min(x - r, 0)``, so
replacement pixels may come from negative (wrap-around) indices; that behaviour is kept on purpose
because the training statistics - and therefore PSNR parity - depend on it."""
import math

import numpy as np
import torch
from torch import Tensor


def _stratified_coords(shape, perc_pix: float = 1.5):
    box = int(np.round(np.sqrt(100 / perc_pix)))
    ys, xs = [], []
    for i in range(int(np.ceil(shape[0] / box))):
        for j in range(int(np.ceil(shape[1] / box))):
            y = int(i * box + torch.rand(1) * box)
            x = int(j * box + torch.rand(1) * box)
            if y < shape[0] and x < shape[1]:
                ys.append(y)
                xs.append(x)
    return ys, xs


def _rand_excluding(lo: int, hi: int, exclude) -> int:
    while True:
        r = int(torch.randint(lo, hi, (1,))[0])
        if r not in exclude:
            return r


def manipulate(image: Tensor, subpatch_size: int = 5, inplace: bool = False):
    """Replace ~1.5 % of the pixels of a CHW image by a random neighbour; returns (image, coords [K, 2])."""
    if subpatch_size % 2 == 0:
        raise ValueError("subpatch_size must be odd")
    if not inplace:
        image = image.clone()
    size_x, size_y = image.shape[2], image.shape[1]
    r = math.floor(subpatch_size / 2)
    coords = []
    for x, y in zip(*_stratified_coords((size_x, size_y))):
        coords.append((x, y))
        rx = _rand_excluding(min(x - r, 0), min(x + r, size_x - 1), [x])
        ry = _rand_excluding(min(y - r, 0), min(y + r, size_y - 1), [y])
        image[:, y, x] = image[:, ry, rx]
    return image, torch.tensor(coords)
