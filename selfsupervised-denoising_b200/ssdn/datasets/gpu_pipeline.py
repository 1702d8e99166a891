"""On-GPU input pipeline: random crops of a uint8 image cache resident in HBM, converted to float and
noised by one CUDA kernel per batch (``ssdn_noisy_crops``), in the ``(input, reference, metadata)`` layout that
``Denoiser.run_pipeline`` consumes.

Reference: ``train.py:756-760`` (RandomCrop), ``datasets/noise_wrapper.py:98-163`` (prepare_input),
``utils/noise.py:14-63`` (add_gaussian).  The reference feeds the GPU from 4 PIL / h5py worker processes; at the engine's
step rate (> 4000 patches/s per GPU) that loader is the bottleneck by orders of magnitude.  Randomness is Philox
counter-based: batches are a pure function of ``(seed, step)`` - a resumed run only needs the step counter - and parity
with the CPU generator is statistical, not bit-wise.  Gaussian (``gauss25``, ``gauss5_50``, ``_nc``) and the reference's
Poisson styles (``poisson30``, ``poisson5_50``: ``ssdn_poisson_crops``).
Noise2Void batches are masked on the device too (``ssdn_n2v_mask``, the reference's ``manipulate`` with its index quirks)."""
from __future__ import annotations

import re
from typing import Dict, List

import torch

from ssdn import _engine as E
from ssdn.datasets.noise_wrapper import NULL_IMAGE, NoisyDataset
from ssdn.params import NoiseAlgorithm


def parse_style(style: str):
    """'gauss25' -> ('gauss', 25/255, 25/255, clip); 'gauss5_50_nc' -> ('gauss', 5/255, 50/255, no clip); 'poisson30' ->
    ('poisson', 30, 30, clip).  Integer parameters of the Gaussian styles are 8-bit units, Poisson parameters are taken
    as written (utils/noise.py:57-58, 112-153)."""
    kind = re.findall(r"[a-zA-Z]+", style)[0]
    if kind not in ("gauss", "poisson"):
        raise NotImplementedError("Noise type not supported")
    tokens = [t for t in style.replace(kind, "").split("_") if t != ""]
    clip = "nc" not in tokens
    tokens = [t for t in tokens if t != "nc"]
    as_float = any("." in t for t in tokens)
    unit = 255.0 if kind == "gauss" and not as_float else 1.0
    vals = [float(t) / unit for t in tokens]
    if len(vals) == 1:
        return kind, vals[0], vals[0], clip
    if len(vals) == 2:
        return kind, vals[0], vals[1], clip
    raise ValueError(f"cannot parse noise style '{style}'")


def parse_gaussian_style(style: str):
    """(sigma_lo, sigma_hi, clip) of a Gaussian style string."""
    kind, lo, hi, clip = parse_style(style)
    if kind != "gauss":
        raise NotImplementedError("not a Gaussian noise style: '{}'".format(style))
    return lo, hi, clip


def image_cache_from_dataset(dataset, limit: int = None) -> torch.Tensor:
    """uint8 [n][C][H][W] cache (CPU) from a dataset of clean float images in [0, 1], e.g. ``UnlabelledImageFolderDataset``;
    8-bit sources survive exactly.  All images must have one size (crop or pad the data set first)."""
    count = len(dataset) if limit is None else min(limit, len(dataset))
    imgs = [dataset[i][0] for i in range(count)]
    sizes = sorted({tuple(t.shape) for t in imgs})
    if len(sizes) != 1:
        raise ValueError("the image cache holds equal-size images, found {}".format(sizes[:4]))
    return (torch.stack(imgs) * 255.0).round().clamp(0, 255).to(torch.uint8)


class GpuNoisyPatches:
    """``batch(step)`` -> ``[input, reference, metadata]`` of ``batch_size`` noisy ``patch`` x ``patch`` crops."""

    def __init__(self, images_u8: torch.Tensor, noise_style: str, algorithm: NoiseAlgorithm, patch: int, batch_size: int, seed: int = 0):
        if not images_u8.is_cuda or images_u8.dtype != torch.uint8 or images_u8.dim() != 4:
            raise E.EngineError("image cache must be a CUDA uint8 tensor [n_images][C][H][W]")
        self.images, self.style, self.algorithm = images_u8.contiguous(), noise_style, algorithm
        self.patch, self.batch_size, self.seed = patch, batch_size, seed
        self.kind, self.sigma_lo, self.sigma_hi, self.clip = parse_style(noise_style)
        self._crops = E.noisy_crops if self.kind == "gauss" else E.poisson_crops

    @classmethod
    def from_dataset(cls, dataset, noise_style: str, algorithm: NoiseAlgorithm, patch: int, batch_size: int, seed: int = 0,
                     device: str = "cuda", limit: int = None) -> "GpuNoisyPatches":
        return cls(image_cache_from_dataset(dataset, limit).to(device), noise_style, algorithm, patch, batch_size, seed)

    def batch(self, step: int) -> List:
        M = NoisyDataset.Metadata
        n, c = self.batch_size, self.images.shape[1]
        clean, noisy, sigma = self._crops(self.images, n, self.patch, self.seed, step, self.sigma_lo, self.sigma_hi, self.clip)
        ranged = self.sigma_hi > self.sigma_lo
        coeff = sigma.reshape(n, c, 1, 1) if ranged else sigma[:, :1].reshape(n, 1, 1, 1)
        md: Dict = {M.CLEAN: clean, M.INPUT_NOISE_VALUES: coeff, M.IMAGE_SHAPE: torch.tensor([[c, self.patch, self.patch]] * n),
                    M.INDEXES: torch.arange(n) + step * n}
        if self.algorithm == NoiseAlgorithm.NOISE_TO_CLEAN:
            ref = clean
            md[M.REFERENCE_NOISE_VALUES] = torch.zeros(n, 1, 1, 1)
        elif self.algorithm in (NoiseAlgorithm.NOISE_TO_NOISE, NoiseAlgorithm.NOISE_TO_VOID):
            _, ref, rs = self._crops(self.images, n, self.patch, self.seed, step, self.sigma_lo, self.sigma_hi, self.clip, stream_id=1,
                                     want_clean=False)
            md[M.REFERENCE_NOISE_VALUES] = rs.reshape(n, c, 1, 1) if ranged else rs[:, :1].reshape(n, 1, 1, 1)
        elif self.algorithm == NoiseAlgorithm.SELFSUPERVISED_DENOISING_MEAN_ONLY:
            ref = noisy
            md[M.REFERENCE_NOISE_VALUES] = coeff
        else:
            ref = NULL_IMAGE
            md[M.REFERENCE_NOISE_VALUES] = torch.zeros(n, 1, 1, 1)
        if self.algorithm == NoiseAlgorithm.NOISE_TO_VOID:      # noise_wrapper.py:107-111: mask the input, keep the coordinates
            noisy, md[M.MASK_COORDS] = E.n2v_mask(noisy, self.seed, step, 5)
        return [noisy, ref, md]

    def __iter__(self):
        step = 0
        while True:
            yield self.batch(step)
            step += 1
