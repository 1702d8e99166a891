"""Network classes of the hot path (reference: ssdn/ssdn/models/)."""
from .noise_network import NoiseNetwork, ShiftConv2d
from .utility import Crop2d, Shift2d

__all__ = ["NoiseNetwork", "ShiftConv2d", "Crop2d", "Shift2d"]
