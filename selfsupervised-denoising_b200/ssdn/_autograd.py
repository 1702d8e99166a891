"""torch.autograd glue around the C-ABI engine: each Function's forward/backward is one engine call."""
from __future__ import annotations

import torch

from ssdn import _engine as E


class NetFunction(torch.autograd.Function):
    """Whole U-Net.  Inputs: (owner module, plan, training flag, x, *parameters in registration order).

    The activations needed by the backward pass live in the plan's workspace, so the backward of a
    forward call must run before the same plan is used for another forward (the usual training loop)."""

    @staticmethod
    def forward(ctx, owner, plan, training, x, *params):
        out = plan.forward(owner.flat_parameters(), x, training=training)
        ctx.owner, ctx.plan, ctx.shapes = owner, plan, [p.shape for p in params]
        plan.forward_token = ctx
        return out

    @staticmethod
    def backward(ctx, dout):
        owner, plan = ctx.owner, ctx.plan
        if getattr(plan, "forward_token", None) is not ctx:
            raise RuntimeError("NoiseNetwork: backward() called after another forward() reused the same plan; the "
                               "engine keeps one set of activations per (batch, size) plan")
        target = owner.grad_buffer()
        if target is not None:
            first = next(iter(owner.parameters())).grad
            if first is not None and first.data_ptr() == target.data_ptr():
                # p.grad already ARE views of the buffer the engine is about to overwrite: autograd would then add the new
                # gradient to itself (2 x new instead of old + new)
                raise RuntimeError("NoiseNetwork: gradients of an earlier backward() are still attached; the engine writes into one flat "
                                   "gradient buffer and cannot accumulate - call zero_grad(set_to_none=True) before every backward pass")
        grads = plan.backward(owner.flat_parameters(), dout.contiguous().float(), target, owner.stale_slot())
        outs, off = [], 0
        for shp in ctx.shapes:
            n = 1
            for s in shp:
                n *= s
            outs.append(grads[off:off + n].view(shp))
            off += n
        return (None, None, None, None, *outs)


class PosteriorFunction(torch.autograd.Function):
    """SSDN posterior mean + NLL for Gaussian or Poisson noise (denoiser.py:222-397).  Only `loss` is differentiable."""

    @staticmethod
    def forward(ctx, net_out, noisy, sigma_raw, sigma_known, poisson=False, diagonal=False):
        net_out, noisy, sigma_raw = net_out.contiguous(), noisy.contiguous(), sigma_raw.contiguous().float()
        pme, loss, model_std, noise_std = E.posterior_forward(net_out, noisy, sigma_raw, sigma_known, poisson, diagonal)
        ctx.save_for_backward(net_out, noisy, sigma_raw)
        ctx.known, ctx.poisson, ctx.diagonal = sigma_known, poisson, diagonal
        ctx.mark_non_differentiable(pme, model_std, noise_std)
        return pme, loss, model_std, noise_std

    @staticmethod
    def backward(ctx, _gpme, gloss, _gms, _gns):
        net_out, noisy, sigma_raw = ctx.saved_tensors
        dnet, dsig = E.posterior_backward(net_out, noisy, sigma_raw, gloss.contiguous().float().view(-1), ctx.known, ctx.poisson, ctx.diagonal)
        return dnet, None, dsig, None, None, None


class SpatialMeanFunction(torch.autograd.Function):
    """mean over (H, W) keeping dims (denoiser.py:264)."""

    @staticmethod
    def forward(ctx, x):
        ctx.shape = x.shape
        return E.spatial_mean_forward(x.contiguous())

    @staticmethod
    def backward(ctx, g):
        return E.spatial_mean_backward(g.contiguous(), ctx.shape)


class MSEFunction(torch.autograd.Function):
    """Per-sample mean squared error -> [N, 1] (denoiser.py:153-154)."""

    @staticmethod
    def forward(ctx, out, ref):
        out, ref = out.contiguous(), ref.contiguous()
        ctx.save_for_backward(out, ref)
        return E.mse_forward(out, ref)

    @staticmethod
    def backward(ctx, gloss):
        out, ref = ctx.saved_tensors
        return E.mse_backward(out, ref, gloss.contiguous().view(-1)), None


class MaskedMSEFunction(torch.autograd.Function):
    """N2V masked loss of the reference (utils/n2v_loss.py + denoiser.py:176-178) -> [N, 1]."""

    @staticmethod
    def forward(ctx, out, ref, coords):
        out, ref = out.contiguous(), ref.contiguous()
        ctx.save_for_backward(out, ref, coords)
        return E.masked_mse_forward(out, ref, coords)

    @staticmethod
    def backward(ctx, gloss):
        out, ref, coords = ctx.saved_tensors
        return E.masked_mse_backward(out, ref, coords, gloss.contiguous().view(-1)), None, None
