"""NoiseNetwork: the blind-spot / plain U-Net (reference: ssdn/ssdn/models/noise_network.py).

Same constructor, parameter names, initialisation and RNG consumption as the reference, so state
dicts are interchangeable.  The sub-modules only HOLD the parameters: ``forward`` hands the whole
network to the B200 engine (tcgen05 implicit-GEMM convolutions, fused rotate/shift/pool/upsample/
concat), which reads all parameters from one flat fp32 buffer that the parameters are views of."""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn
from torch import Tensor

from ssdn import _engine as E
from ssdn._autograd import NetFunction
from ssdn.models.utility import Shift2d

_MAX_PLANS = 2


class ShiftConv2d(nn.Conv2d):
    """Half-plane convolution of Laine et al.: an h x w kernel only sees rows at or above the output row
    (pad h//2 rows on top, convolve, crop h//2 rows at the bottom).  Standalone calls run the engine's
    single-operator path; inside NoiseNetwork the layer is part of the fused plan."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.shift_size = (self.kernel_size[0] // 2, 0)
        shift = Shift2d(self.shift_size)
        self.pad, self.crop = shift.pad, shift.crop

    def forward(self, x: Tensor) -> Tensor:
        return _ConvOp.apply(x, self.weight, self.bias, True)


class _ConvOp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, blind):
        ctx.save_for_backward(x, w)
        ctx.blind = blind
        return E.conv2d_forward(x, w, b, blind=blind, lrelu=False)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx = E.conv2d_backward_data(dy, w, blind=ctx.blind) if ctx.needs_input_grad[0] else None
        dw, db = E.conv2d_backward_weight(x, dy, w.shape[-1], blind=ctx.blind)
        return dx, dw, db, None


class NoiseNetwork(nn.Module):
    """U-Net for N2C/N2N/N2V (``blindspot=False``) and SSDN (``blindspot=True``: four 90-degree rotations stacked on
    the batch, upward-only receptive field, one-pixel shift, un-rotate, 4x channel concat, 1x1 head).

    Args match the reference: in_channels, out_channels, blindspot, zero_output_weights."""

    def __init__(self, in_channels: int = 3, out_channels: int = 3, blindspot: bool = False, zero_output_weights: bool = False):
        super().__init__()
        self._blindspot = blindspot
        self._zero_output_weights = zero_output_weights
        self.in_channels, self.out_channels = in_channels, out_channels
        self.Conv2d = ShiftConv2d if blindspot else nn.Conv2d
        act = lambda: nn.LeakyReLU(negative_slope=0.1, inplace=True)  # noqa: E731
        conv = lambda ci, co: self.Conv2d(ci, co, 3, stride=1, padding=1)  # noqa: E731
        pool = lambda: nn.Sequential(Shift2d((1, 0)), nn.MaxPool2d(2)) if blindspot else nn.MaxPool2d(2)  # noqa: E731
        up = lambda: nn.Upsample(scale_factor=2, mode="nearest")  # noqa: E731

        self.encode_block_1 = nn.Sequential(conv(in_channels, 48), act(), conv(48, 48), act(), pool())
        for i in (2, 3, 4, 5):
            setattr(self, f"encode_block_{i}", nn.Sequential(conv(48, 48), act(), pool()))
        self.encode_block_6 = nn.Sequential(conv(48, 48), act())
        self.decode_block_6 = nn.Sequential(up())
        self.decode_block_5 = nn.Sequential(conv(96, 96), act(), conv(96, 96), act(), up())
        for i in (4, 3, 2):
            setattr(self, f"decode_block_{i}", nn.Sequential(conv(144, 96), act(), conv(96, 96), act(), up()))
        self.decode_block_1 = nn.Sequential(conv(96 + in_channels, 96), act(), conv(96, 96), act())
        if blindspot:
            self.shift = Shift2d((1, 0))
        width = 384 if blindspot else 96
        self.output_conv = self.Conv2d(96, out_channels, 1)
        self.output_block = nn.Sequential(self.Conv2d(width, width, 1), act(), self.Conv2d(width, 96, 1), act(), self.output_conv)
        self.init_weights()
        self._plans: "OrderedDict[tuple, E.NetPlan]" = OrderedDict()
        self._flat = None
        self._grad_buffer = None
        self._stale_slot = None

    @property
    def blindspot(self) -> bool:
        return self._blindspot

    def init_weights(self):
        """He-normal (a = 0.1) for every convolution, zero biases; the last 1x1 is He-normal 'linear' or zeros."""
        with torch.no_grad():
            for m in self.modules():
                if isinstance(m, nn.Conv2d):
                    nn.init.kaiming_normal_(m.weight.data, a=0.1)
                    m.bias.data.zero_()
            if self._zero_output_weights:
                self.output_conv.weight.zero_()
            else:
                nn.init.kaiming_normal_(self.output_conv.weight.data, nonlinearity="linear")

    @staticmethod
    def input_wh_mul() -> int:
        """Input height/width must be a multiple of 2 ** (number of pooling layers)."""
        return 2 ** 5

    # ------------------------------------------------------------------ flat parameter storage
    def flat_parameters(self) -> Tensor:
        """The flat fp32 buffer all parameters are views of (registration order); rebuilt when a
        ``.to()`` / ``load_state_dict(assign=True)`` replaced the parameter storages."""
        params = list(self.parameters())
        flat = self._flat
        ok = flat is not None and flat.device == params[0].device
        if ok:
            off = 0
            for p in params:
                if p.data_ptr() != flat.data_ptr() + 4 * off or p.dtype != torch.float32 or not p.is_contiguous():
                    ok = False
                    break
                off += p.numel()
        if not ok:
            flat = flatten_parameters(params)
            self._flat = flat
        return self._flat

    def adopt_flat(self, flat: Tensor):
        """Called by an owner (Denoiser) that re-homed this network's parameters into a larger flat buffer."""
        self._flat = flat

    def grad_buffer(self):
        return self._grad_buffer

    def set_grad_buffer(self, buf, stale_slot=None):
        """Optional flat buffer the engine writes parameter gradients into (p.grad become views of it) and an optional
        one-float CUDA tensor that receives the step's stale-scale flag (see _engine.NetPlan)."""
        self._grad_buffer = buf
        self._stale_slot = stale_slot

    def stale_slot(self):
        return self._stale_slot

    def _plan(self, x: Tensor) -> E.NetPlan:
        n, c, h, w = x.shape
        key = (n, h, w, x.device.index)
        plan = self._plans.get(key)
        if plan is None:
            plan = E.NetPlan(n, self.in_channels, self.out_channels, h, w, self._blindspot, x.device)
            self._plans[key] = plan
            while len(self._plans) > _MAX_PLANS:
                self._plans.popitem(last=False)
        else:
            self._plans.move_to_end(key)
        return plan

    def forward(self, x: Tensor) -> Tensor:
        if not x.is_cuda:
            raise E.EngineError("NoiseNetwork runs on the B200 engine only: move the module and its input to a CUDA "
                                "device (there is no CPU or PyTorch fallback)")
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise ValueError(f"expected an N x {self.in_channels} x H x W input, got {tuple(x.shape)}")
        x = x.contiguous().float()
        params = list(self.parameters())
        training = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return NetFunction.apply(self, self._plan(x), training, x, *params)


def flatten_parameters(params) -> Tensor:
    """Re-home the given parameters into one contiguous fp32 buffer (values preserved) and return it."""
    total = sum(p.numel() for p in params)
    flat = torch.empty(total, dtype=torch.float32, device=params[0].device)
    off = 0
    with torch.no_grad():
        for p in params:
            n = p.numel()
            flat[off:off + n].copy_(p.data.reshape(-1))
            p.data = flat[off:off + n].view(p.shape)
            off += n
    return flat
