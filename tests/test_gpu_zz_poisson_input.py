"""GPU: the reference's Poisson noise styles in the on-GPU input pipeline (ssdn_poisson_crops; utils/noise.py:66-109).
The reference computes (clean * lam + K) / lam with K ~ Poisson(1) per element; statistical parity with that."""
import math

import pytest
import torch

import ssdn
from ssdn.datasets import GpuNoisyPatches, NoisyDataset
from ssdn.datasets.gpu_pipeline import parse_style
from ssdn.params import NoiseAlgorithm, PipelineOutput
from util import make_cfg

pytestmark = pytest.mark.gpu


def test_poisson_counts_follow_poisson_one(engine):
    imgs = torch.full((2, 3, 96, 96), 128, dtype=torch.uint8)
    lam = 30.0
    cl, no, lm = engine.poisson_crops(imgs.cuda(), 64, 64, seed=1, step=0, lam_lo=lam, clip=False)
    assert torch.equal(lm, torch.full_like(lm, lam))
    k = ((no.double() - cl.double()) * lam).cpu()
    assert float((k - k.round()).abs().max()) < 1e-3                          # integer counts added on the lam scale
    k = k.round()
    n = k.numel()
    assert float(k.min()) == 0.0 and 6 <= float(k.max()) <= 12
    for count in range(6):
        p = math.exp(-1.0) / math.factorial(count)
        assert abs(float((k == count).double().mean()) - p) < 5 * math.sqrt(p * (1 - p) / n) + 1e-5, count
    assert abs(float(k.mean()) - 1.0) < 5 / math.sqrt(n) and abs(float(k.var()) - 1.0) < 0.01
    z = k - 1.0                                                               # independent across channels, pixels, samples
    assert abs(float((z[:, 0] * z[:, 1]).mean())) < 0.012 and abs(float((z[:, :, :, 1:] * z[:, :, :, :-1]).mean())) < 0.012
    assert abs(float((z[0] * z[1]).mean())) < 0.05
    cl2, no2, _ = engine.poisson_crops(imgs.cuda(), 64, 64, seed=1, step=0, lam_lo=lam, clip=False)
    assert torch.equal(no, no2) and torch.equal(cl, cl2)                      # pure function of (seed, step)
    _, no3, _ = engine.poisson_crops(imgs.cuda(), 64, 64, seed=1, step=1, lam_lo=lam, clip=False)
    assert not torch.equal(no, no3)


def test_poisson_crops_share_the_gaussian_kernels_crops_and_clip(engine):
    imgs = torch.randint(0, 256, (6, 3, 40, 48), dtype=torch.uint8, generator=torch.Generator().manual_seed(0)).cuda()
    cg, _, _ = engine.noisy_crops(imgs, 8, 32, seed=7, step=3, sigma_lo=25 / 255, clip=True)
    cp, no, _ = engine.poisson_crops(imgs, 8, 32, seed=7, step=3, lam_lo=30.0, clip=True)
    assert torch.equal(cg, cp)
    assert float(no.min()) >= 0.0 and float(no.max()) <= 1.0 and float((no == 1.0).float().mean()) > 0.001       # clipped at 1
    _, raw, _ = engine.poisson_crops(imgs, 8, 32, seed=7, step=3, lam_lo=30.0, clip=False)
    assert float(raw.max()) > 1.0 and torch.equal(raw.clamp(0, 1), no)
    with pytest.raises(ValueError):
        engine.poisson_crops(imgs, 8, 32, seed=7, step=3, lam_lo=0.0)


def test_poisson_range_style_draws_lam_per_sample_and_channel(engine):
    assert parse_style("poisson5_50_nc") == ("poisson", 5.0, 50.0, False) and parse_style("poisson30") == ("poisson", 30.0, 30.0, True)
    imgs = torch.full((1, 3, 64, 64), 100, dtype=torch.uint8)
    cl, no, lm = engine.poisson_crops(imgs.cuda(), 32, 64, seed=5, step=9, lam_lo=5.0, lam_hi=50.0, clip=False)
    assert float(lm.min()) >= 5.0 and float(lm.max()) <= 50.0 and lm.unique().numel() > 80
    assert 20.0 < float(lm.mean()) < 35.0                                     # U(5, 50) has mean 27.5
    k = (no - cl).reshape(32, 3, -1) * lm[:, :, None]
    assert ((k.mean(dim=2) - 1.0).abs() < 0.1).all()                          # the reported lam is the one applied (4096 px each)


def test_poisson_batches_feed_the_denoiser(engine):
    imgs = torch.randint(0, 256, (4, 3, 64, 80), dtype=torch.uint8, generator=torch.Generator().manual_seed(3)).cuda()
    gen = GpuNoisyPatches(imgs, "poisson30", NoiseAlgorithm.SELFSUPERVISED_DENOISING, patch=32, batch_size=4, seed=11)
    data = gen.batch(0)
    M = NoisyDataset.Metadata
    assert data[0].shape == (4, 3, 32, 32) and data[1].numel() == 0
    assert torch.equal(data[2][M.INPUT_NOISE_VALUES].cpu(), torch.full((4, 1, 1, 1), 30.0))
    den = ssdn.Denoiser(make_cfg("ssdn", "known", 3, style="poisson30"), device="cuda")
    out = den.run_pipeline(data)
    out[PipelineOutput.LOSS].mean().backward()
    assert torch.isfinite(out[PipelineOutput.LOSS]).all() and out[PipelineOutput.NOISE_STD_DEV].shape == (4, 32, 32)
    n2n = GpuNoisyPatches(imgs, "poisson30", NoiseAlgorithm.NOISE_TO_NOISE, patch=32, batch_size=4, seed=11).batch(0)
    assert torch.equal(n2n[2][M.CLEAN], data[2][M.CLEAN]) and not torch.equal(n2n[0], n2n[1])
