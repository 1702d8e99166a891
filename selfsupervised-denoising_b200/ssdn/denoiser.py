"""Denoiser: owns the networks and runs the training / inference pipelines
(reference: ssdn/ssdn/denoiser.py).  Same constructor, attributes, ``run_pipeline`` input/output
dictionaries and state-dict key schema as the reference; the arithmetic runs in the B200 engine.

Differences that are deliberate and documented in DESIGN.md:
  * networks are NOT wrapped in ``nn.DataParallel`` (one process per GPU with a single gradient
    all-reduce replaces it, see ssdn.train); a transparent wrapper keeps the ``models.<id>.module.*`` keys;
  * ``forward`` works (in the reference it raises IndexError for every pipeline);
  * the diagonal-covariance option works (the reference's branch raises TypeError at denoiser.py:240, `c00.shape()`; the
    engine computes Sigma_x = diag(d^2) from the c diagonal factors, which is what that branch evidently means)."""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn as nn
from torch import Tensor

import ssdn
from ssdn import _engine as E
from ssdn._autograd import MSEFunction, PosteriorFunction, SpatialMeanFunction
from ssdn.datasets import NoisyDataset
from ssdn.models import NoiseNetwork
from ssdn.models.noise_network import flatten_parameters
from ssdn.params import ConfigValue, NoiseValue, Pipeline, PipelineOutput


class _Replica(nn.Module):
    """Stands where the reference has ``nn.DataParallel``: exposes the wrapped network as ``.module`` so that
    state-dict keys read ``models.<id>.module.<param>`` exactly like reference checkpoints."""

    def __init__(self, module: nn.Module):
        super().__init__()
        self.module = module

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


class Denoiser(nn.Module):
    MODEL = "denoiser_model"
    SIGMA_ESTIMATOR = "sigma_estimation_model"
    ESTIMATED_SIGMA = "estimated_sigma"

    def __init__(self, cfg: Dict, device: str = None):
        super().__init__()
        self.device = torch.device(device) if device else torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.cfg = cfg
        self.models = nn.ModuleDict()    # what the pipelines call
        self._models = nn.ModuleDict()   # the bare networks
        self.init_networks()
        self.l_params = nn.ParameterDict()
        self.init_l_params()
        self._flat = None
        self._flat_grad = None
        self.dp_world_size = 1           # set by ssdn.train.train_step: ranks that each hold a shard of one global batch

    # ------------------------------------------------------------------ construction
    def init_networks(self):
        c = self.cfg[ConfigValue.IMAGE_CHANNELS]
        ssdn_pipe = self.cfg[ConfigValue.PIPELINE] == Pipeline.SSDN
        if ssdn_pipe:
            out_c = 2 * c if self.cfg[ConfigValue.DIAGONAL_COVARIANCE] else c + c * (c + 1) // 2   # mean + factor of Sigma_x
        else:
            out_c = c
        self.add_model(Denoiser.MODEL, NoiseNetwork(in_channels=c, out_channels=out_c, blindspot=self.cfg[ConfigValue.BLINDSPOT]))
        if ssdn_pipe and self.cfg[ConfigValue.NOISE_VALUE] == NoiseValue.UNKNOWN_VARIABLE:
            self.add_model(Denoiser.SIGMA_ESTIMATOR,
                           NoiseNetwork(in_channels=c, out_channels=1, blindspot=False, zero_output_weights=True))

    def init_l_params(self):
        if self.cfg[ConfigValue.PIPELINE] == Pipeline.SSDN and self.cfg[ConfigValue.NOISE_VALUE] == NoiseValue.UNKNOWN_CONSTANT:
            self.l_params[Denoiser.ESTIMATED_SIGMA] = nn.Parameter(torch.zeros((1, 1, 1, 1), device=self.device))

    def get_model(self, model_id: str, parallelised: bool = True) -> nn.Module:
        return (self.models if parallelised else self._models)[model_id]

    def add_model(self, model_id: str, model: nn.Module, parallelise: bool = True):
        self._models[model_id] = model
        wrapped = _Replica(model) if parallelise else model
        wrapped.to(self.device)
        self.models[model_id] = wrapped

    # ------------------------------------------------------------------ flat parameter / gradient storage
    def flat_parameters(self) -> Tensor:
        """All parameters (nn.Module.parameters() order) as views of ONE flat fp32 buffer: the unit of the
        optimiser update and of the single gradient all-reduce."""
        params = list(self.parameters())
        flat, ok = self._flat, self._flat is not None and self._flat.device == params[0].device
        if ok:
            off = 0
            for p in params:
                if p.data_ptr() != flat.data_ptr() + 4 * off:
                    ok = False
                    break
                off += p.numel()
        if not ok:
            flat = flatten_parameters(params)
            # gradient buffer = the parameters' layout + N_FLAGS trailing floats: every network's backward pass leaves its
            # stale-scale flag there, so the flags ride along in the one gradient all-reduce and gate the Adam kernel
            self._flat_grad_all = torch.zeros(flat.numel() + self.N_FLAGS, dtype=torch.float32, device=flat.device)
            self._flat, self._flat_grad = flat, self._flat_grad_all[:flat.numel()]
            off = 0
            for i, net in enumerate(self._models.values()):
                n = sum(p.numel() for p in net.parameters())
                net.adopt_flat(flat[off:off + n])
                net.set_grad_buffer(self._flat_grad[off:off + n], self._flat_grad_all[flat.numel() + i:flat.numel() + i + 1])
                off += n
        return self._flat

    N_FLAGS = 8

    def flat_gradients_with_flags(self) -> Tensor:
        """flat_gradients() followed by the stale-scale flags: the buffer the data-parallel all-reduce sums."""
        self.flat_gradients()
        return self._flat_grad_all

    def stale_flags(self) -> Tensor:
        """The N_FLAGS trailing floats of the gradient buffer (non-zero: this step's gradients must not be applied)."""
        self.flat_parameters()
        return self._flat_grad_all[self._flat.numel():]

    def flat_gradients(self) -> Tensor:
        """Flat gradient buffer matching flat_parameters().  The engine writes each network's gradients straight into
        its slice (p.grad are views of it); anything that is not already in place - the scalar parameters, or gradients
        produced before the flat storage existed - is copied in."""
        self.flat_parameters()
        off = 0
        for net in self._models.values():
            params = list(net.parameters())
            first = params[0].grad
            if first is None or first.data_ptr() != self._flat_grad.data_ptr() + 4 * off:
                o = off
                for p in params:
                    g = self._flat_grad[o:o + p.numel()]
                    if p.grad is None:
                        g.zero_()
                    elif p.grad.data_ptr() != g.data_ptr():
                        g.copy_(p.grad.reshape(-1))
                    o += p.numel()
            off += sum(p.numel() for p in params)
        for p in self.l_params.values():
            g = self._flat_grad[off:off + p.numel()]
            if p.grad is None:
                g.zero_()
            elif p.grad.data_ptr() != g.data_ptr():
                g.copy_(p.grad.reshape(-1))
            off += p.numel()
        return self._flat_grad

    # ------------------------------------------------------------------ pipelines
    def forward(self, data: Tensor, noise_std: Tensor = None) -> Tensor:
        """Inference on a batch of (noisy) images; SSDN with known sigma needs ``noise_std`` ([N,1,1,1] or [N,C,1,1])."""
        md = {NoisyDataset.Metadata.IMAGE_SHAPE: torch.tensor([list(data.shape[1:])] * data.shape[0])}
        if self.cfg[ConfigValue.PIPELINE] == Pipeline.SSDN and self.cfg[ConfigValue.NOISE_VALUE] == NoiseValue.KNOWN:
            if noise_std is None:
                raise ValueError("sigma-known SSDN inference needs the noise standard deviation")
        md[NoisyDataset.Metadata.INPUT_NOISE_VALUES] = noise_std
        with torch.no_grad():
            return self.run_pipeline([data, None, md])[PipelineOutput.IMG_DENOISED]

    def run_pipeline(self, data: List, **kwargs) -> Dict:
        if self.device.type == "cuda":
            self.flat_parameters()       # parameters / gradient slots must be in their flat home BEFORE autograd records them
        pipeline = self.cfg[ConfigValue.PIPELINE]
        if pipeline == Pipeline.MSE:
            return self._mse_pipeline(data, **kwargs)
        if pipeline == Pipeline.SSDN:
            return self._ssdn_pipeline(data, **kwargs)
        if pipeline == Pipeline.MASK_MSE:
            return self._mask_mse_pipeline(data, **kwargs)
        raise NotImplementedError("Unsupported processing pipeline")

    def _input(self, data: List) -> Tensor:
        return data[NoisyDataset.INPUT].to(self.device, non_blocking=True).float()

    @staticmethod
    def _has_reference(data: List) -> bool:
        return len(data) > NoisyDataset.REFERENCE and data[NoisyDataset.REFERENCE] is not None \
            and data[NoisyDataset.REFERENCE].numel() > 0

    def _mse_pipeline(self, data: List, **kwargs) -> Dict:
        """N2C / N2N / SSDN-mean-only: per-sample mean squared error against the reference image."""
        cleaned = self.models[Denoiser.MODEL](self._input(data))
        out = {PipelineOutput.INPUTS: data, PipelineOutput.IMG_DENOISED: cleaned}
        if self._has_reference(data):
            ref = data[NoisyDataset.REFERENCE].to(self.device, non_blocking=True).float()
            out[PipelineOutput.LOSS] = MSEFunction.apply(cleaned, ref)
        return out

    def _mask_mse_pipeline(self, data: List, **kwargs) -> Dict:
        """Noise2Void: squared error at the masked coordinates only (reference semantics, see utils/n2v_loss)."""
        cleaned = self.models[Denoiser.MODEL](self._input(data))
        out = {PipelineOutput.INPUTS: data, PipelineOutput.IMG_DENOISED: cleaned}
        md = data[NoisyDataset.METADATA] if len(data) > NoisyDataset.METADATA else None
        if self._has_reference(data) and md and NoisyDataset.Metadata.MASK_COORDS in md:
            ref = data[NoisyDataset.REFERENCE].to(self.device, non_blocking=True).float()
            coords = md[NoisyDataset.Metadata.MASK_COORDS]
            if cleaned.is_cuda and getattr(self, "dp_world_size", 1) > 1 and torch.is_grad_enabled():
                # The reference applies the coordinate list of the batch's FIRST sample to every sample (utils/n2v_loss.py:12).
                # One process per GPU holds a shard of that batch: rank 0's first sample is the global batch's first sample, so
                # its list is broadcast (64 x 2 int64) and the sharded step stays equal to the reference's global-batch step.
                first = coords[0:1].to(self.device).long().contiguous().clone()
                torch.distributed.broadcast(first, 0)
                coords = first
            loss = ssdn.utils.n2v_loss.loss_mask_mse(coords, cleaned, ref)
            out[PipelineOutput.LOSS] = loss.reshape(loss.shape[0], -1).mean(1, keepdim=True)
        return out

    def _ssdn_pipeline(self, data: List, **kwargs) -> Dict:
        """Blind-spot network -> per-pixel Gaussian N(mu, Sigma_x); combined with the noise model N(0, sigma^2 I) it gives
        the NLL training loss and the posterior-mean estimate (Laine et al. 2019)."""
        noisy = self._input(data)
        md = data[NoisyDataset.METADATA]
        style = self.cfg[ConfigValue.NOISE_STYLE]
        mode = self.cfg[ConfigValue.NOISE_VALUE]
        c = self.cfg[ConfigValue.IMAGE_CHANNELS]
        assert c in [1, 3]
        # diagonal covariance (cfg DIAGONAL_COVARIANCE): c diagonal factors instead of the triangular one.  The reference's own
        # branch stops with a TypeError (denoiser.py:240, `c00.shape()`); the engine computes what it evidently means.  With
        # one channel the two parameterisations coincide.
        diagonal = bool(self.cfg[ConfigValue.DIAGONAL_COVARIANCE]) and c == 3
        if not (style.startswith("gauss") or style.startswith("poisson")):
            raise NotImplementedError("Noise type not supported")
        poisson = style.startswith("poisson")
        n = noisy.shape[0]
        est_stream = None
        if mode == NoiseValue.UNKNOWN_VARIABLE and noisy.is_cuda:
            # The sigma estimator (a plain U-Net on N images, a quarter of the blind-spot net's pixels) and the main network
            # are independent until the loss: run the estimator on a second stream so that it fills the SMs the main
            # network's small layers leave idle.  autograd replays each backward node on the stream of its forward, so the
            # two backward passes overlap the same way.
            if getattr(self, "_est_stream", None) is None:
                self._est_stream = torch.cuda.Stream(device=noisy.device)
            est_stream = self._est_stream
            est_stream.wait_stream(torch.cuda.current_stream(noisy.device))
            with torch.cuda.stream(est_stream):
                est = self.models[Denoiser.SIGMA_ESTIMATOR](noisy)
                sigma_est = SpatialMeanFunction.apply(est).reshape(n, 1)
            if not torch.cuda.is_current_stream_capturing():
                noisy.record_stream(est_stream)
        net_out = self.models[Denoiser.MODEL](noisy)
        if mode == NoiseValue.KNOWN:
            sigma_raw = md[NoisyDataset.Metadata.INPUT_NOISE_VALUES].to(self.device, non_blocking=True).float().reshape(n, -1)
            stat_shape = (n, 1, 1)
        elif mode == NoiseValue.UNKNOWN_CONSTANT:
            sigma_raw = self.l_params[Denoiser.ESTIMATED_SIGMA].reshape(1, 1).expand(n, 1)
            stat_shape = (1, 1, 1)
        elif mode == NoiseValue.UNKNOWN_VARIABLE:
            if est_stream is not None:
                torch.cuda.current_stream(noisy.device).wait_stream(est_stream)
                sigma_raw = sigma_est
                if not torch.cuda.is_current_stream_capturing():
                    sigma_raw.record_stream(torch.cuda.current_stream(noisy.device))
            else:
                est = self.models[Denoiser.SIGMA_ESTIMATOR](noisy)
                sigma_raw = SpatialMeanFunction.apply(est).reshape(n, 1)
            stat_shape = (n, 1, 1)
        else:
            raise NotImplementedError("Unsupported noise value mode")
        pme, loss, model_std, noise_std = PosteriorFunction.apply(net_out, noisy, sigma_raw, mode == NoiseValue.KNOWN, poisson, diagonal)
        if poisson:             # signal-dependent noise: the level is per pixel (denoiser.py:378-380 -> N x H x W)
            stat_shape = tuple(noise_std.shape)
        return {
            PipelineOutput.INPUTS: data,
            PipelineOutput.IMG_MU: net_out[:, 0:c, ...],
            PipelineOutput.IMG_DENOISED: pme,
            PipelineOutput.LOSS: loss,
            PipelineOutput.NOISE_STD_DEV: noise_std if poisson else noise_std[: stat_shape[0]].reshape(stat_shape),
            PipelineOutput.MODEL_STD_DEV: model_std,
        }

    # ------------------------------------------------------------------ persistence
    def state_dict(self, params_only: bool = False, **kwargs) -> Dict:
        state = super().state_dict(**kwargs)
        if not params_only:
            state["cfg"] = self.cfg
        return state

    @staticmethod
    def from_state_dict(state_dict: Dict, device: str = None) -> "Denoiser":
        denoiser = Denoiser(state_dict["cfg"], device=device)
        denoiser.load_state_dict({k: v for k, v in state_dict.items() if k != "cfg"}, strict=False)
        return denoiser

    def config_name(self) -> str:
        return ssdn.cfg.config_name(self.cfg)
