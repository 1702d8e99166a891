"""Image tensor helpers (reference: ssdn/ssdn/utils/data.py): rotations, clipping, PSNR, tensor <-> PIL image."""
import numpy as np
import torch
from torch import Tensor

from ssdn.utils.data_format import batch, permute_tuple, unbatch

__all__ = ["clip_img", "rotate", "mse2psnr", "calculate_psnr", "tensor2image", "show_tensor_image", "save_tensor_image",
           "set_color_channels"]


def _hw_dims(data_format: str):
    fmt = data_format.upper()
    if "H" not in fmt or "W" not in fmt:
        raise ValueError(f"data format '{data_format}' has no H/W axes")
    return fmt.index("H"), fmt.index("W")


def clip_img(img: Tensor, inplace: bool = False) -> Tensor:
    """Clamp to the valid image range: [0, 1] for floating point, [0, 255] for integer tensors."""
    hi = 1 if img.is_floating_point() else 255
    return img.clamp_(0, hi) if inplace else img.clamp(0, hi)


def rotate(x: Tensor, angle: int, data_format: str = "BCHW") -> Tensor:
    """Rotation by a multiple of 90 degrees as flips/transposes (views, exact).
    90: out[i, j] = in[j, W-1-i]; 180: out[i, j] = in[H-1-i, W-1-j]; 270: out[i, j] = in[H-1-j, i]."""
    h, w = _hw_dims(data_format)
    if angle == 0:
        return x
    if angle == 90:
        return x.flip(w).transpose(h, w)
    if angle == 180:
        return x.flip(w).flip(h)
    if angle == 270:
        return x.flip(h).transpose(h, w)
    raise NotImplementedError("Must be rotation divisible by 90 degrees")


def mse2psnr(mse: Tensor, float_imgs: bool = True) -> Tensor:
    peak = torch.tensor(1.0 if float_imgs else 255.0)
    return 20 * torch.log10(peak) - 10 * torch.log10(mse)


def calculate_psnr(img: Tensor, ref: Tensor, data_format: str = "BCHW") -> Tensor:
    """PSNR per batch element (mean squared error over every non-batch axis)."""
    fmt = data_format.upper()
    dims = tuple(i for i, ch in enumerate(fmt) if ch != "B")
    if img.is_cuda and fmt == "BCHW" and img.dtype == torch.float32:
        from ssdn import _engine as E
        mse = E.mse_forward(img.contiguous(), ref.to(img.device).contiguous()).view(-1)
    else:
        # host-side metric helper of the reference's API (uint8 / CPU images, other axis orders: data-set tools, the differential
        # battery); the hot loop's PSNR (CUDA float BCHW, train.py:243-259) is the engine call above
        mse = ((img - ref) ** 2).mean(dim=dims)
    return mse2psnr(mse, img.is_floating_point())


def tensor2image(img: Tensor, data_format: str = "CHW"):
    """Float image tensor in [0, 1] -> 8-bit PIL image (RGB or L).  A batch becomes one grid image."""
    from PIL import Image
    img = img.detach().cpu()
    if img.dim() == 4:
        import torchvision
        grid = torchvision.utils.make_grid(img.permute(permute_tuple(batch(data_format), "BCHW")))
        data_format = unbatch(data_format)
        img = grid.permute(permute_tuple("CHW", data_format))
    arr = np.clip(img.numpy(), 0, 1).transpose(*permute_tuple(data_format, "WHC"))
    channels = arr.shape[-1]
    if channels not in (1, 3):
        raise NotImplementedError("Cannot convert image with {} channels to PIL image.".format(channels))
    arr = np.uint8(arr * 255)
    return Image.fromarray(arr, mode="RGB") if channels == 3 else Image.fromarray(np.squeeze(arr), mode="L")


def show_tensor_image(img: Tensor, data_format: str = "CHW"):
    tensor2image(img, data_format=data_format).show()


def save_tensor_image(img: Tensor, path: str, data_format: str = "CHW"):
    tensor2image(img, data_format=data_format).save(path)


def set_color_channels(img, channels: int):
    """PIL image with the requested number of channels: grey -> RGB replicates, RGB -> grey is PIL's weighted 'L'."""
    if len(img.getbands()) != channels and channels in (1, 3):
        return img.convert("L" if channels == 1 else "RGB")
    return img
