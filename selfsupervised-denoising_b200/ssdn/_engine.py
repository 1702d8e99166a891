"""ctypes binding of libssdn_b200.so (the C-ABI of the CUDA engine, include/ssdn_b200.h).

PyTorch is used for device memory and streams only.  There is no CPU fallback: if the shared
library is missing or a CUDA device is absent, every operator raises."""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libssdn_b200.so")
_lib = None


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
        L.ssdn_b200_last_error.restype = ctypes.c_char_p
        L.ssdn_conv2d_workspace_bytes.restype = ctypes.c_size_t
        L.ssdn_conv2d_backward_weight_workspace_bytes.restype = ctypes.c_size_t
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise EngineError(f"ssdn_b200 error {rc}: {lib().ssdn_b200_last_error().decode()}")


def _ptr(t):
    if t is None:
        return ctypes.c_void_p(0)
    assert t.is_cuda and t.is_contiguous() and t.dtype in (torch.float32, torch.int32, torch.int64, torch.uint8), \
        "engine tensors must be contiguous CUDA tensors"
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise EngineError("ssdn_b200 operators need CUDA tensors (no CPU fallback exists)")


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------ operators
def conv2d_forward(x, w, bias=None, blind=True, lrelu=True):
    """ShiftConv2d / Conv2d (+LeakyReLU 0.1).  x [N,Cin,H,W], w [Cout,Cin,k,k]."""
    _require_cuda(x, w, bias)
    n, cin, h, wd = x.shape
    cout, _, k, _ = w.shape
    y = torch.empty(n, cout, h, wd, device=x.device, dtype=torch.float32)
    nb = lib().ssdn_conv2d_workspace_bytes(n, cin, h, wd, cout, k)
    ws = _workspace(nb, x.device)
    check(lib().ssdn_conv2d_forward(_ptr(ws), ctypes.c_size_t(ws.numel()), _ptr(x.contiguous()), _ptr(w.contiguous()),
                                    _ptr(bias), _ptr(y), n, cin, h, wd, cout, k, int(blind), int(lrelu), _stream()))
    return y


def conv2d_backward_data(dy, w, blind=True):
    _require_cuda(dy, w)
    n, cout, h, wd = dy.shape
    _, cin, k, _ = w.shape
    dx = torch.empty(n, cin, h, wd, device=dy.device, dtype=torch.float32)
    nb = lib().ssdn_conv2d_workspace_bytes(n, cin, h, wd, cout, k)
    ws = _workspace(nb, dy.device)
    check(lib().ssdn_conv2d_backward_data(_ptr(ws), ctypes.c_size_t(ws.numel()), _ptr(dy.contiguous()), _ptr(w.contiguous()),
                                          _ptr(dx), n, cin, h, wd, cout, k, int(blind), _stream()))
    return dx


def conv2d_backward_weight(x, dy, ksize, blind=True):
    _require_cuda(x, dy)
    n, cin, h, wd = x.shape
    cout = dy.shape[1]
    dw = torch.empty(cout, cin, ksize, ksize, device=x.device, dtype=torch.float32)
    db = torch.empty(cout, device=x.device, dtype=torch.float32)
    nb = lib().ssdn_conv2d_backward_weight_workspace_bytes(n, cin, h, wd, cout, ksize)
    ws = _workspace(nb, x.device)
    check(lib().ssdn_conv2d_backward_weight(_ptr(ws), ctypes.c_size_t(ws.numel()), _ptr(x.contiguous()), _ptr(dy.contiguous()),
                                            _ptr(dw), _ptr(db), n, cin, h, wd, cout, ksize, int(blind), _stream()))
    return dw, db
