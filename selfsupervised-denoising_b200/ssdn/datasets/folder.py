"""Clean images from a directory tree (reference: ssdn/ssdn/datasets/folder.py).  A data format next to the hot path: it
feeds ``NoisyDataset`` (evaluation at image size, CPU training loader) and fills the uint8 image cache of the on-GPU
input pipeline (``GpuNoisyPatches.from_dataset``).  Image sizes are read from the file header through PIL's lazy open."""
import os
from typing import List, Tuple

import torch
from PIL import Image
from torch import Tensor
from torch.utils.data import Dataset

import ssdn
from ssdn.utils.data_format import DataFormat, PIL_FORMAT, permute_tuple

IMG_EXTENSIONS = (".jpg", ".jpeg", ".png", ".ppm", ".bmp", ".pgm", ".tif", ".tiff", ".webp")


def find_files(dir_path: str, extensions=IMG_EXTENSIONS, recursive: bool = False) -> List[str]:
    """Files below ``dir_path`` whose extension matches (case-insensitively); sorted."""
    wanted = tuple(e.lower() for e in extensions)
    found = []
    for root, dirs, files in os.walk(dir_path):
        found += [os.path.join(root, f) for f in files if f.lower().endswith(wanted)]
        if not recursive:
            break
    return sorted(found)


class UnlabelledImageFolderDataset(Dataset):
    """``dataset[i] -> (image tensor in [0, 1], i)``; ``channels`` 1 or 3 (grey is replicated / RGB is weighted to grey);
    ``transform`` receives the PIL image (e.g. torchvision ``RandomCrop``) and may return a PIL image or a tensor.
    As in the reference, ``to_tensor``'s C x H x W result is labelled PIL_FORMAT ("CWH") before it is permuted to
    ``output_format``: with the default "CHW" items therefore come out as C x W x H (transposed images - harmless for
    denoising, kept so that padding sides and crops match the reference); ``output_format=None`` keeps C x H x W."""

    def __init__(self, dir_path: str, extensions=IMG_EXTENSIONS, transform=None, recursive: bool = False,
                 output_format: str = DataFormat.CHW, channels: int = 3):
        assert channels in [1, 3]
        self.dir_path, self.transform, self.output_format, self.channels = dir_path, transform, output_format, channels
        self.files = find_files(dir_path, extensions, recursive)
        if not self.files:
            raise RuntimeError("Found 0 files in directory: {}\nSupported extensions are: {}".format(dir_path, ",".join(extensions)))

    def __len__(self) -> int:
        return len(self.files)

    def __getitem__(self, index: int) -> Tuple[Tensor, int]:
        import torchvision.transforms.functional as F
        with open(self.files[index], "rb") as f:
            img = ssdn.utils.set_color_channels(Image.open(f).convert("RGB"), self.channels)     # torchvision's default loader
        if self.transform:
            img = self.transform(img)
        if not isinstance(img, Tensor):
            img = F.to_tensor(img)
        if self.output_format is not None:
            img = img.permute(permute_tuple(PIL_FORMAT, self.output_format))
        return img, index

    def image_size(self, index: int, ignore_transform: bool = False) -> Tensor:
        """Shape of item ``index`` in the output format, from the file header when no transform can change it."""
        if self.transform is not None and not ignore_transform:
            return torch.tensor(self[index][0].shape)
        with Image.open(self.files[index]) as img:
            width, height = img.size
        cwh = torch.tensor((self.channels, width, height))
        return cwh[list(permute_tuple(DataFormat.CWH, self.output_format))]
