"""Batch sources: the reference's noisy-dataset wrapper and resumable sampler, plus the on-GPU patch pipeline."""
from .folder import UnlabelledImageFolderDataset
from .gpu_pipeline import GpuNoisyPatches
from .hdf5 import HDF5Dataset
from .noise_wrapper import NULL_IMAGE, NoisyDataset
from .sampler import FixedLengthSampler, SamplingOrder

__all__ = ["UnlabelledImageFolderDataset", "HDF5Dataset", "GpuNoisyPatches", "NULL_IMAGE", "NoisyDataset", "FixedLengthSampler", "SamplingOrder"]
