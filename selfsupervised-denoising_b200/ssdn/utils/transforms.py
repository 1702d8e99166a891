"""Dataset transforms (reference: ssdn/ssdn/utils/transforms.py)."""
from typing import NewType

import ssdn

Transform = NewType("Transform", object)


class NoiseTransform:
    """Callable that applies a noise style string ('gauss25', 'poisson30', ...) to a batch of images."""

    def __init__(self, style: str):
        self.style = style

    def __call__(self, imgs):
        return ssdn.utils.noise.add_style(imgs, self.style)[0]
