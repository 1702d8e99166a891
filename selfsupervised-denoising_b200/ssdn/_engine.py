"""ctypes binding of libssdn_b200.so (the C-ABI of the CUDA engine, include/ssdn_b200.h).

PyTorch is used for device memory, streams and autograd bookkeeping only.  There is no CPU
fallback: if the shared library is missing or a tensor is not on a CUDA device, every operator raises."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_int, c_longlong, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# SSDN_B200_LIB: another build of the same engine (developer A/B runs); there is no other implementation to fall back to
LIB_PATH = os.environ.get("SSDN_B200_LIB") or os.path.join(os.path.dirname(_HERE), "libssdn_b200.so")
_lib = None

_P, _I, _Z, _LL, _D = c_void_p, c_int, c_size_t, c_longlong, c_double
# name -> (restype, argtypes); mirrors include/ssdn_b200.h
_SIGNATURES = {
    "ssdn_b200_last_error": (ctypes.c_char_p, []),
    "ssdn_b200_version": (_I, []),
    "ssdn_conv2d_workspace_bytes": (_Z, [_I] * 6),
    "ssdn_conv2d_forward": (_I, [_P, _Z, _P, _P, _P, _P] + [_I] * 8 + [_P]),
    "ssdn_conv2d_backward_data": (_I, [_P, _Z, _P, _P, _P] + [_I] * 7 + [_P]),
    "ssdn_conv2d_backward_weight_workspace_bytes": (_Z, [_I] * 6),
    "ssdn_conv2d_backward_weight": (_I, [_P, _Z, _P, _P, _P, _P] + [_I] * 7 + [_P]),
    "ssdn_maxpool2_workspace_bytes": (_Z, [_I] * 4),
    "ssdn_maxpool2": (_I, [_P, _Z, _P, _P, _P, _P] + [_I] * 5 + [_P]),
    "ssdn_net_create": (_I, [_I] * 6 + [ctypes.POINTER(c_void_p)]),
    "ssdn_net_destroy": (None, [_P]),
    "ssdn_net_workspace_bytes": (_Z, [_P]),
    "ssdn_net_param_count": (_Z, [_P]),
    "ssdn_net_bind": (_I, [_P, _P, _Z, _P]),
    "ssdn_net_forward": (_I, [_P, _P, _P, _P, _I, _P]),
    "ssdn_net_backward": (_I, [_P, _P, _P, _P, _P, _P]),
    "ssdn_net_scale_status": (_I, [_P, ctypes.POINTER(c_int), _P]),
    "ssdn_net_debug_scales": (_I, [_P, ctypes.POINTER(c_int), ctypes.POINTER(ctypes.c_uint), _P]),
    "ssdn_net_check": (_I, [_P, _P]),
    "ssdn_net_kernel_launches": (_I, [_P, _I]),
    "ssdn_profile_begin": (_I, []),
    "ssdn_profile_kinds": (_I, []),
    "ssdn_profile_end": (_I, [ctypes.POINTER(c_double)]),
    "ssdn_profile_records": (_I, [ctypes.POINTER(c_double), _I]),
    "ssdn_net_debug_write": (_I, [_P, ctypes.c_char_p, _I, _P, _P]),
    "ssdn_net_debug_read": (_I, [_P, ctypes.c_char_p, _I, _I, _P, ctypes.POINTER(c_int), _P]),
    "ssdn_rot4_stack": (_I, [_P, _P] + [_I] * 4 + [_P]),
    "ssdn_shift_unrot_concat": (_I, [_P, _P] + [_I] * 4 + [_P]),
    "ssdn_loss_workspace_bytes": (_Z, [_I, _I]),
    "ssdn_posterior_forward": (_I, [_P, _P, _P, _P] + [_I] * 7 + [_P, _P, _P, _P, _P]),
    "ssdn_posterior_backward": (_I, [_P, _P, _P, _P, _P] + [_I] * 7 + [_P, _P, _P]),
    "ssdn_spatial_mean_forward": (_I, [_P, _I, _I, _P, _P]),
    "ssdn_spatial_mean_backward": (_I, [_P, _I, _I, _P, _P]),
    "ssdn_mse_forward": (_I, [_P, _P, _P, _I, _I, _P, _P]),
    "ssdn_mse_backward": (_I, [_P, _P, _P, _I, _I, _P, _P]),
    "ssdn_masked_mse_forward": (_I, [_P, _P, _P, _P] + [_I] * 5 + [_P, _P]),
    "ssdn_masked_mse_backward": (_I, [_P, _P, _P, _I, _P] + [_I] * 4 + [_P, _P]),
    "ssdn_adam_step": (_I, [_P, _P, _P, _P, _LL, _D, _D, _D, _D, _LL, _D, _P, _I, _P]),
    "ssdn_adam_step_dev": (_I, [_P, _P, _P, _P, _LL, _P, _P, _I, _P]),
    "ssdn_tensor_peak": (_I, [_I, _D, ctypes.POINTER(c_double), _P]),
    "ssdn_n2v_mask": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, ctypes.c_ulonglong, ctypes.c_ulonglong, _P]),
    "ssdn_noisy_crops": (_I, [_P, _I, _I, _I, _I, _P, _I, _I, ctypes.c_ulonglong, ctypes.c_ulonglong, _I, ctypes.c_float, ctypes.c_float, _I,
                              _P, _P, _P, _P]),
    "ssdn_poisson_crops": (_I, [_P, _I, _I, _I, _I, _P, _I, _I, ctypes.c_ulonglong, ctypes.c_ulonglong, _I, ctypes.c_float, ctypes.c_float, _I,
                                _P, _P, _P, _P]),
}
EXPORTS = tuple(_SIGNATURES)


class EngineError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(the engine has no CPU or PyTorch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        msg = lib().ssdn_b200_last_error().decode()
        if rc == -1:
            raise ValueError(f"ssdn_b200: {msg}")
        raise EngineError(f"ssdn_b200 error {rc}: {msg}")


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise EngineError("ssdn_b200 operators need CUDA tensors (no CPU fallback exists)")
    assert t.is_contiguous(), "engine tensors must be contiguous"
    return t.data_ptr()


def _f32(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise EngineError("ssdn_b200 operators need CUDA tensors (no CPU fallback exists)")
    return t.detach().contiguous().float()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _on_tensor_device(fn):
    """Run an operator with the device of its first tensor argument current: the engine launches on the CURRENT device's
    current stream, so device pointers, stream and kernels then agree whatever device the caller had selected."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        t = next((a for a in args if torch.is_tensor(a)), None)
        if t is None or not t.is_cuda:
            return fn(*args, **kwargs)
        with torch.cuda.device(t.device):
            return fn(*args, **kwargs)
    return wrapped


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------ conv operators
@_on_tensor_device
def conv2d_forward(x, w, bias=None, blind=True, lrelu=True):
    """ShiftConv2d / Conv2d (+LeakyReLU 0.1).  x [N,Cin,H,W], w [Cout,Cin,k,k]."""
    x, w, bias = _f32(x), _f32(w), _f32(bias)
    n, cin, h, wd = x.shape
    cout, _, k, _ = w.shape
    y = torch.empty(n, cout, h, wd, device=x.device, dtype=torch.float32)
    ws = _workspace(lib().ssdn_conv2d_workspace_bytes(n, cin, h, wd, cout, k), x.device)
    check(lib().ssdn_conv2d_forward(_ptr(ws), ws.numel(), _ptr(x), _ptr(w), _ptr(bias), _ptr(y), n, cin, h, wd, cout, k,
                                    int(blind), int(lrelu), _stream()))
    return y


@_on_tensor_device
def conv2d_backward_data(dy, w, blind=True):
    dy, w = _f32(dy), _f32(w)
    n, cout, h, wd = dy.shape
    _, cin, k, _ = w.shape
    dx = torch.empty(n, cin, h, wd, device=dy.device, dtype=torch.float32)
    ws = _workspace(lib().ssdn_conv2d_workspace_bytes(n, cin, h, wd, cout, k), dy.device)
    check(lib().ssdn_conv2d_backward_data(_ptr(ws), ws.numel(), _ptr(dy), _ptr(w), _ptr(dx), n, cin, h, wd, cout, k,
                                          int(blind), _stream()))
    return dx


@_on_tensor_device
def conv2d_backward_weight(x, dy, ksize, blind=True):
    x, dy = _f32(x), _f32(dy)
    n, cin, h, wd = x.shape
    cout = dy.shape[1]
    dw = torch.empty(cout, cin, ksize, ksize, device=x.device, dtype=torch.float32)
    db = torch.empty(cout, device=x.device, dtype=torch.float32)
    ws = _workspace(lib().ssdn_conv2d_backward_weight_workspace_bytes(n, cin, h, wd, cout, ksize), x.device)
    check(lib().ssdn_conv2d_backward_weight(_ptr(ws), ws.numel(), _ptr(x), _ptr(dy), _ptr(dw), _ptr(db), n, cin, h, wd, cout,
                                            ksize, int(blind), _stream()))
    return dw, db


@_on_tensor_device
def maxpool2(x, blind=True, dy=None):
    """Shift2d((1,0)) + MaxPool2d(2) (blind) or MaxPool2d(2) of x [N,C,H,W] (C % 8 == 0) through the network's pool kernels.
    With dy: also returns dz for x = LeakyReLU(z) (the network's fused backward).  -> y or (y, dz)."""
    x = _f32(x)
    n, c, h, w = x.shape
    y = torch.empty(n, c, h // 2, w // 2, device=x.device, dtype=torch.float32)
    dz = torch.empty_like(x) if dy is not None else None
    dy = _f32(dy)
    ws = _workspace(lib().ssdn_maxpool2_workspace_bytes(n, c, h, w), x.device)
    check(lib().ssdn_maxpool2(_ptr(ws), ws.numel(), _ptr(x), _ptr(y), _ptr(dy), _ptr(dz), n, c, h, w, int(blind), _stream()))
    return y if dz is None else (y, dz)


# ------------------------------------------------------------------------------------ index operators
@_on_tensor_device
def rot4_stack(x):
    x = _f32(x)
    n, c, h, w = x.shape
    y = torch.empty(4 * n, c, h, w, device=x.device, dtype=torch.float32)
    check(lib().ssdn_rot4_stack(_ptr(x), _ptr(y), n, c, h, w, _stream()))
    return y


@_on_tensor_device
def shift_unrot_concat(x):
    x = _f32(x)
    b, c, h, w = x.shape
    assert b % 4 == 0
    y = torch.empty(b // 4, 4 * c, h, w, device=x.device, dtype=torch.float32)
    check(lib().ssdn_shift_unrot_concat(_ptr(x), _ptr(y), b // 4, c, h, w, _stream()))
    return y


# ------------------------------------------------------------------------------------ whole network
class NetPlan:
    """One U-Net execution plan for a fixed (N, Cin, Cout, H, W, blindspot): owns the workspace."""

    def __init__(self, n, cin, cout, h, w, blindspot, device):
        self.key = (n, cin, cout, h, w, bool(blindspot))
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise EngineError("the ssdn_b200 network runs on CUDA devices only (no CPU fallback exists)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        handle = c_void_p()
        check(lib().ssdn_net_create(n, cin, cout, h, w, int(blindspot), ctypes.byref(handle)))
        self.handle = handle
        self.n_params = lib().ssdn_net_param_count(handle)
        with torch.cuda.device(self.device):
            self.ws = _workspace(lib().ssdn_net_workspace_bytes(handle), self.device)
            check(lib().ssdn_net_bind(handle, _ptr(self.ws), self.ws.numel(), _stream()))
        self.out_shape = (n, cout, h, w)
        self.fwd_settled = False
        self.bwd_settled = False

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                lib().ssdn_net_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # Operand scales (include/ssdn_b200.h, csrc/common.cuh): the engine reads activations and gradients as fp16 plane pairs
    # scaled by per-tensor powers of two taken from the previous pass.  A pass whose maxima left the accurate band is stale
    # and is simply run again with the scales it left behind.  The first passes of a plan are checked (one stream
    # synchronisation each) until one is clean; inference passes are always checked, their result is read back anyway.
    # Unchecked training passes are covered by the stale flag that backward() leaves for the optimiser step.
    MAX_SCALE_PASSES = 8

    def scale_status(self):
        """(forward stale, backward stale, stale passes since bind); synchronises the current stream."""
        buf = (c_int * 3)()
        with torch.cuda.device(self.device):
            check(lib().ssdn_net_scale_status(self.handle, buf, _stream()))
        return int(buf[0]), int(buf[1]), int(buf[2])

    def debug_scales(self):
        k, a = (c_int * 96)(), (ctypes.c_uint * 96)()
        check(lib().ssdn_net_debug_scales(self.handle, k, a, _stream()))
        return list(k), list(a)

    def forward(self, flat_params, x, training=True, verify=None):
        assert flat_params.numel() == self.n_params and flat_params.dtype == torch.float32
        out = torch.empty(self.out_shape, device=self.device, dtype=torch.float32)
        if verify is None:
            verify = (not training) or not self.fwd_settled
        if x.device != self.device or flat_params.device != self.device:
            raise EngineError(f"plan lives on {self.device}, got tensors on {x.device} / {flat_params.device}")
        with torch.cuda.device(self.device):                       # kernels, streams and events belong to the plan's device
            for _ in range(self.MAX_SCALE_PASSES):
                check(lib().ssdn_net_forward(self.handle, _ptr(flat_params), _ptr(x), _ptr(out), int(training), _stream()))
                if not verify:
                    return out
                if not self.scale_status()[0]:
                    self.fwd_settled = True
                    return out
        raise EngineError("operand scales of the forward pass did not settle (non-finite activations?)")

    def backward(self, flat_params, dout, grads=None, stale_out=None, verify=None):
        if grads is None:
            grads = torch.empty(self.n_params, device=self.device, dtype=torch.float32)
        if verify is None:
            verify = not self.bwd_settled
        with torch.cuda.device(self.device):
            for _ in range(self.MAX_SCALE_PASSES):
                check(lib().ssdn_net_backward(self.handle, _ptr(flat_params), _ptr(dout), _ptr(grads), _ptr(stale_out), _stream()))
                if not verify:
                    return grads
                if not self.scale_status()[1]:
                    self.bwd_settled = True
                    return grads
        raise EngineError("operand scales of the backward pass did not settle (non-finite gradients?)")

    def check(self):
        """Synchronises; raises EngineError if a kernel's bounded pipeline wait expired since the last check (the flag is
        cleared by the check, so the next one reports only new trouble)."""
        with torch.cuda.device(self.device):
            check(lib().ssdn_net_check(self.handle, _stream()))

    def debug_read(self, name, channels, plane=0):
        """Internal activation / gradient buffer `name` (see net.cuh) as a dense [B, channels, H, W] tensor."""
        dims = (c_int * 4)()
        check(lib().ssdn_net_debug_read(self.handle, name.encode(), plane, channels, None, dims, _stream()))
        out = torch.empty(tuple(dims), device=self.device, dtype=torch.float32)
        check(lib().ssdn_net_debug_read(self.handle, name.encode(), plane, channels, _ptr(out), dims, _stream()))
        return out

    def debug_write(self, name, tensor):
        """Overwrite the first tensor.shape[1] channels of internal buffer `name` (test hook)."""
        t = _f32(tensor)
        check(lib().ssdn_net_debug_write(self.handle, name.encode(), t.shape[1], _ptr(t), _stream()))

    def kernel_launches(self, training=True):
        return lib().ssdn_net_kernel_launches(self.handle, int(training))


# ------------------------------------------------------------------------------------ losses
def _loss_ws(n, c, device):
    return _workspace(lib().ssdn_loss_workspace_bytes(n, c), device)


def _posterior_channels(net_out, c, diagonal):
    want = 2 * c if diagonal else c + c * (c + 1) // 2
    if net_out.shape[1] != want:
        raise ValueError("network output has {} channels, the posterior needs {} ({} covariance)".format(
            net_out.shape[1], want, "diagonal" if diagonal else "full"))


@_on_tensor_device
def posterior_forward(net_out, noisy, sigma_raw, sigma_known, poisson=False, diagonal=False):
    """poisson: sigma_raw is the known lambda (sigma_known) or the raw estimate of the per-unit-signal variance; the noise
    level is then per pixel and noise_std comes back as [n][h][w] instead of [n][1][1] (denoiser.py:285-297, :375-380).
    diagonal: net_out carries c diagonal factors of Sigma_x instead of the triangular one (denoiser.py:213, :236-243)."""
    n, c, h, w = noisy.shape
    _posterior_channels(net_out, c, diagonal)
    cs = sigma_raw.numel() // n
    dev = noisy.device
    pme = torch.empty_like(noisy)
    loss = torch.empty(n, 1, device=dev)
    model_std = torch.empty(n, h, w, device=dev)
    noise_std = torch.empty(n, h, w, device=dev) if poisson else torch.empty(n, 1, 1, device=dev)
    ws = _loss_ws(n, c, dev)
    check(lib().ssdn_posterior_forward(_ptr(ws), _ptr(net_out), _ptr(noisy), _ptr(sigma_raw), n, c, h, w, cs, int(sigma_known), int(poisson) | (2 if diagonal else 0),
                                       _ptr(pme), _ptr(loss), _ptr(model_std), _ptr(noise_std), _stream()))
    return pme, loss, model_std, noise_std


@_on_tensor_device
def posterior_backward(net_out, noisy, sigma_raw, gloss, sigma_known, poisson=False, diagonal=False):
    n, c, h, w = noisy.shape
    _posterior_channels(net_out, c, diagonal)
    cs = sigma_raw.numel() // n
    dnet = torch.empty_like(net_out)
    dsig = None if sigma_known else torch.empty_like(sigma_raw)
    ws = _loss_ws(n, c, noisy.device)
    check(lib().ssdn_posterior_backward(_ptr(ws), _ptr(net_out), _ptr(noisy), _ptr(sigma_raw), _ptr(gloss), n, c, h, w, cs,
                                        int(sigma_known), int(poisson) | (2 if diagonal else 0), _ptr(dnet), _ptr(dsig), _stream()))
    return dnet, dsig


@_on_tensor_device
def spatial_mean_forward(x):
    n, c, h, w = x.shape
    out = torch.empty(n, c, 1, 1, device=x.device)
    check(lib().ssdn_spatial_mean_forward(_ptr(x), n * c, h * w, _ptr(out), _stream()))
    return out


@_on_tensor_device
def spatial_mean_backward(g, shape):
    n, c, h, w = shape
    dx = torch.empty(shape, device=g.device)
    check(lib().ssdn_spatial_mean_backward(_ptr(g), n * c, h * w, _ptr(dx), _stream()))
    return dx


@_on_tensor_device
def mse_forward(a, b):
    n = a.shape[0]
    loss = torch.empty(n, 1, device=a.device)
    ws = _loss_ws(n, 1, a.device)
    check(lib().ssdn_mse_forward(_ptr(ws), _ptr(a), _ptr(b), n, a.numel() // n, _ptr(loss), _stream()))
    return loss


@_on_tensor_device
def mse_backward(a, b, gloss):
    n = a.shape[0]
    da = torch.empty_like(a)
    check(lib().ssdn_mse_backward(_ptr(a), _ptr(b), _ptr(gloss), n, a.numel() // n, _ptr(da), _stream()))
    return da


@_on_tensor_device
def masked_mse_forward(out, ref, coords):
    n, c, h, w = out.shape
    loss = torch.empty(n, 1, device=out.device)
    ws = _loss_ws(n, c, out.device)
    check(lib().ssdn_masked_mse_forward(_ptr(ws), _ptr(out), _ptr(ref), _ptr(coords), coords.shape[0], n, c, h, w, _ptr(loss),
                                        _stream()))
    return loss


@_on_tensor_device
def masked_mse_backward(out, ref, coords, gloss):
    n, c, h, w = out.shape
    dout = torch.empty_like(out)
    check(lib().ssdn_masked_mse_backward(_ptr(out), _ptr(ref), _ptr(coords), coords.shape[0], _ptr(gloss), n, c, h, w, _ptr(dout),
                                         _stream()))
    return dout


@_on_tensor_device
def adam_step(p, g, m, v, lr, step, beta1=0.9, beta2=0.99, eps=1e-8, grad_scale=1.0, skip=None):
    """In-place torch.optim.Adam update of the flat fp32 buffer p (train.py:100-107 hyper-parameters).
    skip: optional CUDA float tensor (<= 8 values); the update is a no-op when any of them is non-zero (stale-gradient flags)."""
    check(lib().ssdn_adam_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), float(lr), beta1, beta2, eps, int(step),
                               float(grad_scale), _ptr(skip), 0 if skip is None else skip.numel(), _stream()))


@_on_tensor_device
def adam_step_dev(p, g, m, v, hyper, skip=None):
    """adam_step with {lr / bias_correction1, beta1, beta2, eps, sqrt(bias_correction2), grad_scale} in the CUDA tensor `hyper`."""
    check(lib().ssdn_adam_step_dev(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), _ptr(hyper), _ptr(skip), 0 if skip is None else skip.numel(),
                                   _stream()))


def tensor_peak(f16=True, seconds=1.0):
    """Measured full-chip tcgen05.mma rate of the current device: {tflops, flop_per_clk_sm, sm_mhz, seconds}."""
    buf = (c_double * 4)()
    check(lib().ssdn_tensor_peak(int(f16), float(seconds), buf, _stream()))
    return {"tflops": buf[0], "flop_per_clk_sm": buf[1], "sm_mhz": buf[2], "seconds": buf[3]}


def profile_begin():
    check(lib().ssdn_profile_begin())


PROFILE_KINDS = ("conv_fwd", "conv_dgrad", "wgrad", "wgrad_reduce", "pool_fwd", "pool_bwd", "up_bwd", "pack", "weight_prep", "bias", "scale",
                 "posterior_fwd", "posterior_bwd", "adam", "other")


def profile_end():
    """-> {kind: (launches, device ms, algorithmic FLOPs, algorithmic HBM bytes)} for every kind that launched."""
    nk = lib().ssdn_profile_kinds()
    assert nk == len(PROFILE_KINDS)
    buf = (c_double * (4 * nk))()
    check(lib().ssdn_profile_end(buf))
    return {k: (int(buf[4 * i]), buf[4 * i + 1], buf[4 * i + 2], buf[4 * i + 3]) for i, k in enumerate(PROFILE_KINDS) if buf[4 * i] > 0}


def profile_records(max_records=1024):
    """[(kind, ms, algorithmic FLOPs, algorithmic bytes)] per launch of the last profiled region, in launch order."""
    buf = (c_double * (4 * max_records))()
    n = min(lib().ssdn_profile_records(buf, max_records), max_records)
    return [(PROFILE_KINDS[int(buf[4 * i])], buf[4 * i + 1], buf[4 * i + 2], buf[4 * i + 3]) for i in range(n)]


@_on_tensor_device
def noisy_crops(images_u8, n, patch, seed, step, sigma_lo, sigma_hi=None, clip=True, order=None, stream_id=0, want_clean=True):
    """Random crops of a uint8 image cache [n_images][C][H][W] on the device with synthetic Gaussian noise
    (include/ssdn_b200.h: ssdn_noisy_crops).  Returns (clean or None, noisy, sigma [n][C])."""
    if not images_u8.is_cuda or images_u8.dtype != torch.uint8 or images_u8.dim() != 4:
        raise EngineError("the image cache must be a CUDA uint8 tensor [n_images][C][H][W] (no CPU fallback exists)")
    images_u8 = images_u8.contiguous()
    ni, c, h, w = images_u8.shape
    dev = images_u8.device
    clean = torch.empty(n, c, patch, patch, device=dev) if want_clean else None
    noisy = torch.empty(n, c, patch, patch, device=dev)
    sigma = torch.empty(n, c, device=dev)
    if order is not None:
        order = order.to(device=dev, dtype=torch.int32).contiguous()
    hi = sigma_lo if sigma_hi is None else sigma_hi
    check(lib().ssdn_noisy_crops(_ptr(images_u8), ni, c, h, w, _ptr(order), n, patch, int(seed), int(step), int(stream_id), float(sigma_lo),
                                 float(hi), 1 if clip else 0, _ptr(clean), _ptr(noisy), _ptr(sigma), _stream()))
    return clean, noisy, sigma


@_on_tensor_device
def poisson_crops(images_u8, n, patch, seed, step, lam_lo, lam_hi=None, clip=True, order=None, stream_id=0, want_clean=True):
    """The same crops with the reference's Poisson styles (include/ssdn_b200.h: ssdn_poisson_crops; utils/noise.py:66-109).
    Returns (clean or None, noisy, lam [n][C])."""
    if not images_u8.is_cuda or images_u8.dtype != torch.uint8 or images_u8.dim() != 4:
        raise EngineError("the image cache must be a CUDA uint8 tensor [n_images][C][H][W] (no CPU fallback exists)")
    images_u8 = images_u8.contiguous()
    ni, c, h, w = images_u8.shape
    dev = images_u8.device
    clean = torch.empty(n, c, patch, patch, device=dev) if want_clean else None
    noisy = torch.empty(n, c, patch, patch, device=dev)
    lam = torch.empty(n, c, device=dev)
    if order is not None:
        order = order.to(device=dev, dtype=torch.int32).contiguous()
    hi = lam_lo if lam_hi is None else lam_hi
    check(lib().ssdn_poisson_crops(_ptr(images_u8), ni, c, h, w, _ptr(order), n, patch, int(seed), int(step), int(stream_id), float(lam_lo),
                                   float(hi), 1 if clip else 0, _ptr(clean), _ptr(noisy), _ptr(lam), _stream()))
    return clean, noisy, lam


@_on_tensor_device
def n2v_mask(noisy, seed, step, subpatch_size=5):
    """Noise2Void uniform pixel selection on the device (utils/n2v_ups.py:7-49).  Returns (masked copy, coords int64 [n][k][2])."""
    noisy = _f32(noisy)
    n, c, h, w = noisy.shape
    masked = torch.empty_like(noisy)
    coords = torch.empty(n, (h // 8) * (w // 8), 2, dtype=torch.int64, device=noisy.device)
    check(lib().ssdn_n2v_mask(_ptr(noisy), _ptr(masked), _ptr(coords), n, c, h, w, int(subpatch_size), int(seed), int(step), _stream()))
    return masked, coords
