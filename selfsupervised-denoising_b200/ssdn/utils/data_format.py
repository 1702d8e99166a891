"""Axis-order strings ("CHW", "BCWH", ...) and the permutation between two of them (reference: ssdn/ssdn/utils/data_format.py).
A format is a string with one letter per axis: B batch, C channel, H height, W width."""
from typing import Tuple


class DataFormat:
    BHWC = "BHWC"
    BWHC = "BWHC"
    BCHW = "BCHW"
    BCWH = "BCWH"
    HWC = "HWC"
    WHC = "WHC"
    CHW = "CHW"
    CWH = "CWH"


# what the reference calls the Pillow formats; a tensor from torchvision's to_tensor is treated as being in PIL_FORMAT
PIL_FORMAT = DataFormat.CWH
PIL_BATCH_FORMAT = DataFormat.BCWH


def batch(data_format: str) -> str:
    """The same format with a leading batch axis (unchanged if it already has one)."""
    return data_format if "B" in data_format else "B" + data_format


def unbatch(data_format: str) -> str:
    return data_format.replace("B", "")


def permute_tuple(cur: str, target: str) -> Tuple[int, ...]:
    """Argument for ``Tensor.permute`` / ``ndarray.transpose`` that turns axis order ``cur`` into ``target``."""
    assert sorted(cur) == sorted(target)
    return tuple(cur.index(axis) for axis in target)
