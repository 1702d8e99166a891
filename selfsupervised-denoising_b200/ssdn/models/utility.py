"""Crop2d / Shift2d (reference: ssdn/ssdn/models/utility.py).

Pure index operators.  Inside :class:`ssdn.models.NoiseNetwork` they are never executed as separate
passes - the engine folds every shift into the address arithmetic of the neighbouring kernel - but
the classes are kept for API compatibility and for standalone use (any device, exact)."""
from typing import Tuple

import torch
import torch.nn as nn
from torch import Tensor


class Crop2d(nn.Module):
    """Remove ``(left, right, top, bottom)`` pixels from the borders of a BCHW tensor."""

    def __init__(self, crop: Tuple[int, int, int, int]):
        super().__init__()
        assert len(crop) == 4
        self.crop = crop

    def forward(self, x: Tensor) -> Tensor:
        left, right, top, bottom = self.crop
        return x[:, :, top:x.shape[-2] - bottom, left:x.shape[-1] - right]


class Shift2d(nn.Module):
    """``out[y, x] = in[y - vert, x - horz]`` with zeros shifted in (positive = towards bottom / right)."""

    def __init__(self, shift: Tuple[int, int]):
        super().__init__()
        self.shift = shift
        vert, horz = shift
        top, bottom = (abs(vert), 0) if vert >= 0 else (0, abs(vert))
        left, right = (abs(horz), 0) if horz >= 0 else (0, abs(horz))
        self.pad = nn.ZeroPad2d((left, right, top, bottom))
        self.crop = Crop2d((right, left, bottom, top))

    def forward(self, x: Tensor) -> Tensor:
        return self.crop(self.pad(x))
