"""Callable transforms for data sets (reference: ssdn/ssdn/utils/transforms.py).

``Transform`` only labels "anything a data set may call on a loaded image" in type hints.  ``NoiseTransform("gauss25")(x)``
returns ``x`` with the named synthetic noise applied - the noisy half of ``ssdn.utils.noise.add_style``'s result, the noise
parameters are dropped."""
from typing import NewType

Transform = NewType("Transform", object)


class NoiseTransform:
    def __init__(self, style: str):
        self.style = style

    def __call__(self, imgs):
        from ssdn.utils.noise import add_style
        noisy, _params = add_style(imgs, self.style)
        return noisy
