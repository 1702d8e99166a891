from ssdn.datasets.noise_wrapper import NoisyDataset, NULL_IMAGE
from ssdn.datasets.sampler import FixedLengthSampler, SamplingOrder

__all__ = ["NoisyDataset", "NULL_IMAGE", "FixedLengthSampler", "SamplingOrder"]
