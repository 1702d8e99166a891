"""Noise2Void masked loss (reference: ssdn/ssdn/utils/n2v_loss.py).

Semantics kept bug-for-bug: the coordinate list of the FIRST sample is used for every sample of the
batch and indexes ``[:, :, c0, c1]``; squared errors are summed over the coordinates -> N x C."""
from torch import Tensor


def loss_mask_mse(masked_coords: Tensor, input: Tensor, target: Tensor) -> Tensor:
    if input.is_cuda:
        from ssdn._autograd import MaskedMSEFunction
        coords = masked_coords[0].to(device=input.device).long().contiguous()
        # engine result is already averaged over channels (N x 1); expand so the caller's .view(N,-1).mean(1) is a no-op
        return MaskedMSEFunction.apply(input, target.to(input.device), coords).expand(-1, input.shape[1])
    # CPU tensors: this is a public helper of the reference's API (ssdn.utils.n2v_loss.loss_mask_mse) that host-side code and the
    # differential battery (tests/host_battery.py) call on CPU tensors, where it must answer like the reference.  It is NOT a
    # fallback of the training path: Denoiser pipelines refuse CPU inputs (EngineError) and CUDA tensors take the branch above.
    acc = 0
    for x, y in masked_coords.tolist()[0]:
        acc = acc + (target[:, :, x, y] - input[:, :, x, y]) ** 2
    return acc
