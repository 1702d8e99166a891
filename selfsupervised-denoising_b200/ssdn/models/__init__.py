from ssdn.models.utility import Shift2d, Crop2d
from ssdn.models.noise_network import NoiseNetwork, ShiftConv2d

__all__ = ["Shift2d", "Crop2d", "NoiseNetwork", "ShiftConv2d"]
