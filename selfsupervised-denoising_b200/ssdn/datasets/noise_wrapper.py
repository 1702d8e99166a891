"""NoisyDataset: turns a dataset of clean CHW images into (input, reference, metadata) training
triples (reference: ssdn/ssdn/datasets/noise_wrapper.py).  ``Denoiser.run_pipeline`` consumes exactly
this tuple/metadata layout, which is why the wrapper is part of the drop-in surface.  CPU code."""
from enum import Enum
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch
from torch import Tensor
from torch.utils.data import Dataset

import ssdn
from ssdn.params import NoiseAlgorithm

NULL_IMAGE = torch.zeros(0)


class NoisyDataset(Dataset):
    INPUT = 0
    REFERENCE = 1
    METADATA = 2

    class Metadata(Enum):
        CLEAN = 1
        IMAGE_SHAPE = 2
        INDEXES = 3
        INPUT_NOISE_VALUES = 4
        REFERENCE_NOISE_VALUES = 5
        MASK_COORDS = 6

    def __init__(self, child: Dataset, noise_style: str, algorithm: NoiseAlgorithm, pad_uniform: bool = False,
                 pad_multiple: int = None, square: bool = False, enable_metadata: bool = True, data_format: str = "CHW",
                 training_mode: bool = False):
        self.child, self.noise_style, self.algorithm = child, noise_style, algorithm
        self.pad_uniform, self.pad_multiple, self.square = pad_uniform, pad_multiple, square
        self.enable_metadata, self.data_format, self.training_mode = enable_metadata, data_format, training_mode
        self._max_image_size = None
        if pad_uniform:
            _ = self.max_image_size

    def __len__(self) -> int:
        return len(self.child)

    def __getitem__(self, index: int):
        clean = self.child[index][0]
        md = {NoisyDataset.Metadata.INDEXES: index} if self.enable_metadata else None
        inp, ref, md = self.prepare_input(clean, md)
        return (inp, ref, md) if self.enable_metadata else (inp, ref)

    def prepare_input(self, clean: Tensor, metadata: Optional[Dict] = None) -> Tuple[Tensor, Tensor, Dict]:
        M = NoisyDataset.Metadata
        scratch = metadata if metadata is not None else {}
        noisy, coeff = ssdn.utils.noise.add_style(clean, self.noise_style)
        if self.algorithm == NoiseAlgorithm.NOISE_TO_VOID and self.training_mode:
            noisy, coords = ssdn.utils.n2v_ups.manipulate(noisy, 5)
            scratch[M.MASK_COORDS] = coords
        if self.algorithm == NoiseAlgorithm.NOISE_TO_CLEAN:
            ref, ref_coeff = clean, 0
        elif self.algorithm in (NoiseAlgorithm.NOISE_TO_NOISE, NoiseAlgorithm.NOISE_TO_VOID):
            ref, ref_coeff = ssdn.utils.noise.add_style(clean, self.noise_style)
        elif self.algorithm == NoiseAlgorithm.SELFSUPERVISED_DENOISING:
            ref, ref_coeff = NULL_IMAGE, 0
        elif self.algorithm == NoiseAlgorithm.SELFSUPERVISED_DENOISING_MEAN_ONLY:
            ref, ref_coeff = noisy, coeff
        else:
            raise NotImplementedError("Denoising algorithm not supported")
        inp = self.pad_to_output_size(noisy)
        if ref is not NULL_IMAGE:
            ref = self.pad_to_output_size(ref)
        if metadata is not None:
            metadata[M.CLEAN] = self.pad_to_output_size(clean)
            metadata[M.IMAGE_SHAPE] = torch.tensor(clean.shape)
            metadata[M.INPUT_NOISE_VALUES] = torch.zeros((1, 1, 1)) + coeff      # [1,1,1] or [C,1,1]; batching adds N
            metadata[M.REFERENCE_NOISE_VALUES] = torch.zeros((1, 1, 1)) + ref_coeff
        return inp, ref, metadata

    @property
    def max_image_size(self) -> Tensor:
        if self._max_image_size is None:
            try:
                sizes = [self.child.image_size(i) for i in range(len(self.child))]
            except AttributeError:
                sizes = [torch.tensor(d[0].shape) for d in self.child]
            self._max_image_size = torch.stack(sizes).max(dim=0).values
        return self._max_image_size

    def get_output_size(self, image: Tensor) -> Tensor:
        fmt = self.data_format.upper()
        hd, wd = fmt.index("H"), fmt.index("W")
        size = list(self.max_image_size if self.pad_uniform else image.shape)
        size = [int(s) for s in size]
        if self.pad_multiple:
            m = self.pad_multiple
            size[hd] = (size[hd] + m - 1) // m * m
            size[wd] = (size[wd] + m - 1) // m * m
        if self.square:
            size[hd] = size[wd] = max(size[hd], size[wd])
        return torch.tensor(size)

    def pad_to_output_size(self, image: Tensor) -> Tensor:
        """Reflect-pad on the right/bottom up to the configured output size."""
        fmt = self.data_format.upper()
        if fmt not in ("CHW", "CWH", "BCHW", "BCWH"):
            raise NotImplementedError("Padding not supported by data format")
        target = self.get_output_size(image)
        if all(int(t) == s for t, s in zip(target, image.shape)):
            return image
        pads = [[0, int(t) - s] if ch in "HW" else [0, 0] for t, s, ch in zip(target, image.shape, fmt)]
        return torch.tensor(np.pad(image, pads, mode="reflect"), device=image.device, requires_grad=image.requires_grad)

    @staticmethod
    def _unpad_single(image: Tensor, shape: Tensor) -> Tensor:
        return image[tuple(slice(0, int(s)) for s in shape)]

    @staticmethod
    def unpad(image: Tensor, metadata: Dict, batch_index: int = None) -> Union[Tensor, List[Tensor]]:
        """Crop back to the original (top-left) image region recorded in the metadata."""
        shape = metadata[NoisyDataset.Metadata.IMAGE_SHAPE]
        if batch_index is not None:
            image, shape = image[batch_index], shape[batch_index]
        if image.dim() <= shape.shape[-1]:
            return NoisyDataset._unpad_single(image, shape)
        return [NoisyDataset._unpad_single(i, s) for i, s in zip(image, shape)]
