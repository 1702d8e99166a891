"""Batch sources: the reference's noisy-dataset wrapper and resumable sampler, plus the on-GPU patch pipeline."""
from .gpu_pipeline import GpuNoisyPatches
from .noise_wrapper import NULL_IMAGE, NoisyDataset
from .sampler import FixedLengthSampler, SamplingOrder

__all__ = ["GpuNoisyPatches", "NULL_IMAGE", "NoisyDataset", "FixedLengthSampler", "SamplingOrder"]
