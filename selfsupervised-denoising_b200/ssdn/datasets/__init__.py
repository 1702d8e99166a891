from ssdn.datasets.noise_wrapper import NoisyDataset, NULL_IMAGE
from ssdn.datasets.sampler import FixedLengthSampler, SamplingOrder
from ssdn.datasets.gpu_pipeline import GpuNoisyPatches

__all__ = ["NoisyDataset", "NULL_IMAGE", "FixedLengthSampler", "SamplingOrder", "GpuNoisyPatches"]
