"""Golden fixture from a TRAINED reference checkpoint (build container only).

    python tests/golden/make_wt_golden.py

Loads /root/reference/models/final-ssdn-gauss25-sigma_known.wt with the UNMODIFIED reference Denoiser
(Denoiser.from_state_dict, eval.py:36-43), runs Denoiser.run_pipeline on a seeded synthetic batch
(smooth clean images + clipped Gaussian noise sigma = 25/255) and stores the trained parameters together
with the reference outputs (posterior mean, network mean, per-sample loss, PSNRs).  The parity tests load
the parameters into the oracle (CPU) and into the CUDA engine (GPU box) and must reproduce the outputs
within 1e-4 relative and the PSNRs within 1e-3 dB - SURVEY.md 8(d) "PSNR parity (i)"."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_shim import import_reference  # noqa: E402

ssdn = import_reference()
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import ssdn_oracle as O  # noqa: E402
from ssdn.datasets import NoisyDataset  # noqa: E402
from ssdn.denoiser import Denoiser  # noqa: E402
from ssdn.params import PipelineOutput  # noqa: E402

WT = "/root/reference/models/final-ssdn-gauss25-sigma_known.wt"
torch.set_num_threads(8)


def psnr(a, b):
    return -10.0 * torch.log10(((a - b) ** 2).mean(dim=(1, 2, 3)))


def main():
    state = torch.load(WT, map_location="cpu", weights_only=False)
    den = Denoiser.from_state_dict(state)
    den.eval()
    clean, noisy = O.synthetic_batch(2, 3, 64, seed=4242)
    sigma = torch.full((2, 1, 1, 1), 25.0 / 255.0)
    md = {NoisyDataset.Metadata.INPUT_NOISE_VALUES: sigma, NoisyDataset.Metadata.CLEAN: clean,
          NoisyDataset.Metadata.IMAGE_SHAPE: torch.tensor([[3, 64, 64]] * 2)}
    with torch.no_grad():
        out = den.run_pipeline([noisy, torch.zeros(0), md])
    pme, mu, loss = out[PipelineOutput.IMG_DENOISED], out[PipelineOutput.IMG_MU], out[PipelineOutput.LOSS]
    params = {k[len("_models.denoiser_model."):]: v for k, v in state.items() if k.startswith("_models.denoiser_model.") and torch.is_tensor(v)}
    params = {k: v for k, v in params.items() if not k.startswith("output_block.4")}
    assert len(params) == 40, sorted(params)
    # the oracle must reproduce the reference at the trained weights too
    ref = O.ssdn_pipeline(params, noisy, sigma, "known")
    for name, a, b in (("pme", ref["pme"], pme), ("mu", ref["mu"], mu), ("loss", ref["loss"].view(-1), loss.view(-1))):
        err = (a - b).abs().max().item() / b.abs().max().item()
        assert err < 1e-5, (name, err)
    arrs = {"p." + k: v.numpy() for k, v in params.items()}
    arrs.update(clean=clean.numpy(), noisy=noisy.numpy(), sigma=sigma.numpy(), pme=pme.numpy(), mu=mu.numpy(), loss=loss.view(-1).numpy(),
                psnr_in=psnr(noisy, clean).numpy(), psnr_pme=psnr(pme, clean).numpy(), psnr_mu=psnr(mu, clean).numpy(),
                model_std=out[PipelineOutput.MODEL_STD_DEV].numpy())
    np.savez_compressed(os.path.join(HERE, "wt_ssdn_gauss25_sigma_known.npz"), **arrs)
    print("PSNR in", psnr(noisy, clean).tolist(), "pme", psnr(pme, clean).tolist(), "mu", psnr(mu, clean).tolist(), "loss", loss.view(-1).tolist())


if __name__ == "__main__":
    main()
