"""Clean images from an HDF5 file written by the reference's ``dataset_tool_h5.py`` (reference: ssdn/ssdn/datasets/hdf5.py):
dataset ``images`` holds one flattened uint8 array per image, ``shapes`` its shape in ``h5_format`` axis order.  Needs the
optional ``h5py`` package; the constructor raises ImportError without it."""
from typing import Tuple

import numpy as np
import torch
from PIL import Image
from torch import Tensor
from torch.utils.data import Dataset

import ssdn
from ssdn.utils.data_format import DataFormat, PIL_FORMAT, permute_tuple


class HDF5Dataset(Dataset):
    def __init__(self, file_path: str, transform=None, h5_format: str = PIL_FORMAT, output_format: str = DataFormat.CHW, channels: int = 3):
        try:
            import h5py
        except ImportError as e:
            raise ImportError("HDF5Dataset needs the 'h5py' package, which is not installed") from e
        self._h5py = h5py
        self.file_path, self.transform, self.h5_format, self.output_format, self.channels = file_path, transform, h5_format, output_format, channels
        with h5py.File(file_path, "r") as f:
            self.img_count = f["images"].shape[0]

    def __len__(self) -> int:
        return self.img_count

    def __getitem__(self, index: int) -> Tuple[Tensor, int]:
        import torchvision.transforms.functional as F
        with self._h5py.File(self.file_path, "r") as f:
            flat, shape = f["images"][index], f["shapes"][index]
        arr = np.reshape(flat, shape).transpose(*permute_tuple(self.h5_format, "WHC"))
        img = ssdn.utils.set_color_channels(Image.fromarray(arr), self.channels)
        if self.transform:
            img = self.transform(img)
        if not isinstance(img, Tensor):
            img = F.to_tensor(img)
        if self.output_format is not None:
            img = img.permute(permute_tuple(PIL_FORMAT, self.output_format))
        return img, index

    def image_size(self, index: int, ignore_transform: bool = False) -> Tensor:
        if self.transform is not None and not ignore_transform:
            return torch.tensor(self[index][0].shape)
        with self._h5py.File(self.file_path, "r") as f:
            shape = f["shapes"][index]
        return torch.tensor(shape[list(permute_tuple(self.h5_format, self.output_format))])
