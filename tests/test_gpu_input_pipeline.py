"""GPU: on-GPU input pipeline (ssdn_noisy_crops, ssdn_poisson_crops, ssdn_n2v_mask) - crops are bit-exact copies of the image
cache, noise has the reference's statistics (utils/noise.py:14-109), batches are a pure function of (seed, step)."""
import math

import pytest
import torch

import ssdn
from ssdn.datasets import GpuNoisyPatches, NoisyDataset
from ssdn.datasets.gpu_pipeline import parse_gaussian_style, parse_style
from ssdn.params import NoiseAlgorithm, PipelineOutput
from util import make_cfg

pytestmark = pytest.mark.gpu


def _cache(n=6, c=3, h=40, w=48, seed=0):
    return torch.randint(0, 256, (n, c, h, w), dtype=torch.uint8, generator=torch.Generator().manual_seed(seed))


def test_crops_are_exact_and_reproducible(engine):
    imgs = _cache()
    cl, no, sg = engine.noisy_crops(imgs.cuda(), 8, 32, seed=7, step=3, sigma_lo=25 / 255, clip=True)
    cl2, no2, _ = engine.noisy_crops(imgs.cuda(), 8, 32, seed=7, step=3, sigma_lo=25 / 255, clip=True)
    assert torch.equal(cl, cl2) and torch.equal(no, no2)                      # pure function of (seed, step)
    _, no3, _ = engine.noisy_crops(imgs.cuda(), 8, 32, seed=7, step=4, sigma_lo=25 / 255, clip=True)
    assert not torch.equal(no, no3)
    ref = imgs.float() / 255.0                                                # torchvision ToTensor
    offsets = set()
    for i in range(8):
        img = ref[(3 * 8 + i) % 6]
        found = [(oy, ox) for oy in range(40 - 32 + 1) for ox in range(48 - 32 + 1) if torch.equal(img[:, oy:oy + 32, ox:ox + 32], cl[i].cpu())]
        assert len(found) >= 1, i                                             # bit-exact crop of the right image
        offsets.add(found[0])
    assert len(offsets) > 3                                                   # crop origins vary between samples
    assert float(no.min()) >= 0.0 and float(no.max()) <= 1.0                  # clipped
    assert torch.allclose(sg, torch.full_like(sg, 25 / 255))


def test_noise_statistics_match_add_gaussian(engine):
    imgs = torch.full((2, 3, 96, 96), 128, dtype=torch.uint8)                 # mid-grey: clipping never triggers at sigma 25
    cl, no, _ = engine.noisy_crops(imgs.cuda(), 64, 64, seed=1, step=0, sigma_lo=25 / 255, clip=False)
    d = (no - cl).double().cpu()
    n = d.numel()
    sigma = 25 / 255
    assert abs(d.mean().item()) < 5 * sigma / math.sqrt(n)
    assert abs(d.std().item() / sigma - 1) < 5e-3
    z = d / sigma
    assert abs((z ** 4).mean().item() - 3.0) < 0.05 and abs((z ** 3).mean().item()) < 0.02      # Gaussian moments
    # independent across pixels, channels and samples
    assert abs((z[:, 0] * z[:, 1]).mean().item()) < 0.01 and abs((z[:, :, :, 1:] * z[:, :, :, :-1]).mean().item()) < 0.01
    assert abs((z[0] * z[1]).mean().item()) < 0.02


def test_range_style_draws_sigma_per_sample_and_channel(engine):
    lo, hi, clip = parse_gaussian_style("gauss5_50_nc")
    assert (lo, hi, clip) == (5 / 255, 50 / 255, False) and parse_gaussian_style("gauss25") == (25 / 255, 25 / 255, True)
    imgs = torch.full((1, 3, 64, 64), 100, dtype=torch.uint8)
    cl, no, sg = engine.noisy_crops(imgs.cuda(), 32, 64, seed=5, step=9, sigma_lo=lo, sigma_hi=hi, clip=clip)
    assert float(sg.min()) >= lo and float(sg.max()) <= hi and sg.unique().numel() > 80          # [32][3] distinct levels
    emp = (no - cl).reshape(32, 3, -1).std(dim=2)
    assert ((emp / sg - 1).abs() < 0.05).all()                                                  # reported sigma is the one applied
    assert 20 / 255 < float(sg.mean()) < 35 / 255                                               # U(5, 50) / 255 has mean 27.5 / 255


def test_pipeline_batches_feed_the_denoiser(engine):
    imgs = _cache(n=4, h=64, w=80, seed=3).cuda()
    gen = GpuNoisyPatches(imgs, "gauss25", NoiseAlgorithm.SELFSUPERVISED_DENOISING, patch=32, batch_size=4, seed=11)
    data = gen.batch(0)
    M = NoisyDataset.Metadata
    assert data[0].shape == (4, 3, 32, 32) and data[2][M.INPUT_NOISE_VALUES].shape == (4, 1, 1, 1) and data[1].numel() == 0
    den = ssdn.Denoiser(make_cfg("ssdn", "known", 3), device="cuda")
    out = den.run_pipeline(data)
    assert torch.isfinite(out[PipelineOutput.LOSS]).all() and out[PipelineOutput.IMG_DENOISED].shape == (4, 3, 32, 32)
    n2n = GpuNoisyPatches(imgs, "gauss25", NoiseAlgorithm.NOISE_TO_NOISE, patch=32, batch_size=4, seed=11).batch(0)
    assert torch.equal(n2n[2][M.CLEAN], data[2][M.CLEAN])                     # same crops ...
    a, b = (n2n[0] - n2n[2][M.CLEAN]), (n2n[1] - n2n[2][M.CLEAN])
    assert abs(float((a * b).mean()) / float((a * a).mean())) < 0.05          # ... independent noise realisations
    with pytest.raises(NotImplementedError):
        GpuNoisyPatches(imgs, "speckle3", NoiseAlgorithm.SELFSUPERVISED_DENOISING, 32, 4)
    with pytest.raises(ValueError):
        engine.noisy_crops(imgs, 4, 128, 0, 0, 0.1)                           # patch larger than the images


def test_n2v_mask_follows_the_reference_selection_rule(engine):
    """utils/n2v_ups.py:7-49: one coordinate per 8 x 8 box in the reference's list order; the masked pixel equals (in every
    channel) a pixel of the same image whose column / row lie in the reference's (quirky) ranges
    [min(x - 2, 0), min(x + 2, W - 1)) excluding x, with Python-style negative wrap; everything else is untouched."""
    g = torch.Generator().manual_seed(9)
    noisy = torch.rand(4, 3, 64, 64, generator=g)
    masked, coords = engine.n2v_mask(noisy.cuda(), seed=3, step=17)
    masked2, coords2 = engine.n2v_mask(noisy.cuda(), seed=3, step=17)
    assert torch.equal(masked, masked2) and torch.equal(coords, coords2)
    masked, coords = masked.cpu(), coords.cpu()
    assert coords.shape == (4, 64, 2) and coords.dtype == torch.int64
    changed = (masked != noisy).any(dim=1)
    offs = []
    for n in range(4):
        keep = torch.ones(64, 64, dtype=torch.bool)
        for b, (x, y) in enumerate(coords[n].tolist()):
            i, j = b // 8, b % 8
            assert 8 * i <= x < 8 * i + 8 and 8 * j <= y < 8 * j + 8          # stratified, x outer / y inner as in the reference
            keep[y, x] = False
            cols = [c % 64 for c in range(min(x - 2, 0), min(x + 2, 63)) if c != x]
            rows = [r % 64 for r in range(min(y - 2, 0), min(y + 2, 63)) if r != y]
            cand = [(r, c) for r in rows for c in cols if torch.equal(noisy[n, :, r, c], masked[n, :, y, x])]
            assert cand, (n, x, y)                                                # copied from an allowed source, all channels alike
            offs.append(cand[0][1] - x)
        assert not changed[n][keep].any()                                         # nothing else was touched
    assert len(set(offs)) > 20                                                    # sources spread over the whole allowed range


def test_n2v_batches_feed_the_denoiser(engine):
    imgs = _cache(n=4, h=80, w=80, seed=5).cuda()
    gen = GpuNoisyPatches(imgs, "gauss25", NoiseAlgorithm.NOISE_TO_VOID, patch=64, batch_size=4, seed=2)
    data = gen.batch(1)
    M = NoisyDataset.Metadata
    assert data[2][M.MASK_COORDS].shape == (4, 64, 2) and data[1].shape == data[0].shape
    den = ssdn.Denoiser(make_cfg("n2v", None, 3), device="cuda")
    out = den.run_pipeline(data)
    out[PipelineOutput.LOSS].mean().backward()
    assert torch.isfinite(out[PipelineOutput.LOSS]).all()


# ---------------------------------------------------------------- the reference's Poisson styles (utils/noise.py:66-109):
# (clean * lam + K) / lam with K ~ Poisson(1) per element
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
