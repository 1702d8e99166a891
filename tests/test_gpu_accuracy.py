"""GPU: parity at BASELINE.json's full sizes, the accuracy properties of the fp16 two-term operand split, the shifted
max-pool at operator level, evaluation at 512 / 768 pixels and a 200-step convergence run against the CPU oracle."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import ssdn
import ssdn_oracle as O
from oracle_trace import oracle_trace
from ssdn.datasets import NoisyDataset
from ssdn.params import PipelineOutput
from ssdn.train import FlatAdam, GraphedTrainStep, train_step
from util import TOL, make_cfg, rel

pytestmark = pytest.mark.gpu
M = NoisyDataset.Metadata
HERE = os.path.dirname(os.path.abspath(__file__))


def _trained_params():
    z = np.load(os.path.join(HERE, "golden", "wt_ssdn_gauss25_sigma_known.npz"))
    return {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p.")}, z


# ------------------------------------------------------------------------------------------ shifted max-pool, operator level
@pytest.mark.parametrize("blind", [True, False])
def test_shifted_maxpool_ties_and_negative_first_row(engine, blind):
    """models/noise_network.py:64-67.  The activations are small dyadic numbers (exact in the fp16 planes), so the pooled
    output must be BIT-exact and the gradient must be routed to exactly the elements ATen's max_pool2d picks: many exact
    ties inside the 2x2 windows (first maximum in row-major order wins), an all-negative first image row (blind: the zero
    row shifted in from above wins and swallows the gradient), zeros.  The routed gradient is multiplied by LeakyReLU'
    (1 or 0.1f) of the winner: exact for 1, to the planes' 2^-22 for 0.1f."""
    g = torch.Generator().manual_seed(3)
    vals = torch.tensor([-2.0, -1.0, -0.5, -0.25, 0.0, 0.25, 0.5, 1.0, 2.0])
    a = vals[torch.randint(0, len(vals), (3, 16, 8, 12), generator=g)]
    a[:, :, 0, :] = -vals[torch.randint(5, len(vals), (3, 16, 12), generator=g)]          # first row strictly negative
    a[:, :, 2:4, 2:6] = 0.5                                                                # whole windows tied
    a.requires_grad_(True)
    y_ref = O.maxpool2(a, blind)
    dy = vals[torch.randint(0, len(vals), tuple(y_ref.shape), generator=g)]
    y_ref.backward(dy)
    dz_ref = a.grad * torch.where(a.detach() > 0, torch.tensor(1.0), torch.tensor(0.1))     # through LeakyReLU(z) = a
    y, dz = engine.maxpool2(a.detach().cuda(), blind=blind, dy=dy.cuda())
    y, dz = y.cpu(), dz.cpu()
    assert torch.equal(y, y_ref.detach())
    assert torch.equal(dz != 0, dz_ref != 0)                                               # identical routing
    pos = a.detach() > 0
    assert torch.equal(dz[pos], dz_ref[pos])
    assert torch.allclose(dz, dz_ref, rtol=1e-6, atol=0)
    if blind:       # the padding row won every window of the first output row whose real elements are negative
        assert (y_ref[:, :, 0, :] == 0).all() and (a.grad[:, :, 0, :] == 0).all() and (dz[:, :, 0, :] == 0).all()


# ------------------------------------------------------------------------------------------ accumulator-truncation compensation
def test_accumulator_compensation_on_trained_trace(engine):
    """csrc/common.cuh SSDN_ACC_BETA: the tensor core's truncating fp32 accumulator biases every result by a factor
    proportional to the number of MMA instructions; the epilogues multiply the midpoint of its measured range back.  Pinned
    here on the TRAINED checkpoint: every convolution of the network, fed with the exact (fp64 oracle) input of that layer,
    must keep its signed bias below 0.7e-8 per MMA instruction (1.7e-6 for the deepest accumulation) and its rms error below
    5e-6; the signed biases of the 20 layers together stay below 2e-5 - a fifth of the 1e-4 parity bar."""
    params, z = _trained_params()
    noisy = torch.from_numpy(z["noisy"])
    with torch.enable_grad():
        _, T = oracle_trace({k: v.double().requires_grad_(True) for k, v in params.items()}, noisy.double(), True)
    pools = T["pools"][1]
    inputs = {"encode_block_1.0": T["x"][1], "encode_block_1.2": T["encode_block_1.0"][1], "encode_block_2.0": pools[0], "encode_block_3.0": pools[1],
              "encode_block_4.0": pools[2], "encode_block_5.0": pools[3], "encode_block_6.0": pools[4], "decode_block_5.0": T["cat5"][1],
              "decode_block_5.2": T["decode_block_5.0"][1], "decode_block_4.0": T["cat4"][1], "decode_block_4.2": T["decode_block_4.0"][1],
              "decode_block_3.0": T["cat3"][1], "decode_block_3.2": T["decode_block_3.0"][1], "decode_block_2.0": T["cat2"][1],
              "decode_block_2.2": T["decode_block_2.0"][1], "decode_block_1.0": T["cat1"][1], "decode_block_1.2": T["decode_block_1.0"][1],
              "output_block.0": T["head_in"][1], "output_block.2": T["output_block.0"][1], "output_conv": T["output_block.2"][1]}
    total = 0.0
    for name, x64 in inputs.items():
        x = x64.detach().float()
        w, b = params[name + ".weight"], params[name + ".bias"]
        k = w.shape[-1]
        ref = (O.shift_conv2d if k == 3 else O.conv2d_same)(x.double(), w.double(), b.double())
        got = engine.conv2d_forward(x.cuda(), w.cuda(), b.cuda(), blind=(k == 3), lrelu=False).double().cpu()
        err = got - ref
        n_mma = (w.shape[1] + 15) // 16 * k * k * 3
        bias = ((err * torch.sign(ref)).mean() / ref.abs().mean()).item()
        rms = (err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
        total += bias
        assert abs(bias) < 0.7e-8 * n_mma + 1e-7, (name, bias, rms)
        assert rms < 5e-6, (name, bias, rms)
    assert abs(total) < 2e-5, total


# ------------------------------------------------------------------------------------------ BASELINE.json shapes, full size
CONFIGS = {  # name: (algorithm, sigma mode, batch, patch, noise style)      BASELINE.json configs[1..4]
    "cfg2_ssdn_known_64_bs32": ("ssdn", "known", 32, 64, "gauss25"),
    "cfg3_ssdn_var_64_bs32": ("ssdn", "var", 32, 64, "gauss25"),
    "cfg4_n2v_64_bs32": ("n2v", None, 32, 64, "gauss25"),
    "cfg5_ssdn_var_gauss5_50_128_bs16": ("ssdn", "var", 16, 128, "gauss5_50"),
}


@pytest.mark.parametrize("name", list(CONFIGS))
def test_baseline_configurations_full_size(engine, name):
    """One training step of every BASELINE.json configuration at its FULL batch and patch size: per-sample loss and the
    network mean of the first two samples against the oracle (samples are independent units, so two of them pin the batch),
    1e-4 relative - or the fp32 reference path's own distance from exact arithmetic where that is larger; device-side error
    flags clean; finite; the step must then lower the loss on the same batch."""
    algo, mode, n, size, style = CONFIGS[name]
    torch.manual_seed(0)
    den = ssdn.Denoiser(make_cfg(algo, mode or "known", 3, style), device="cuda")
    opt = FlatAdam(den)
    opt.param_groups[0]["lr"] = 3e-4
    clean, noisy = O.synthetic_batch(n, 3, size, seed=5)
    g = torch.Generator().manual_seed(1)
    md = {M.CLEAN: clean}
    md[M.INPUT_NOISE_VALUES] = (torch.rand(n, 3, 1, 1, generator=g) * 45 + 5) / 255 if style == "gauss5_50" else torch.full((n, 1, 1, 1), 25 / 255)
    ref = torch.zeros(0)
    if algo == "n2v":
        ref = (clean + torch.randn(clean.shape, generator=g) * 25 / 255).clamp(0, 1)
        md[M.MASK_COORDS] = torch.randint(0, size, (n, 64, 2), generator=g)
    data = [noisy, ref, md]
    main = den.get_model(ssdn.Denoiser.MODEL, False)
    params = {a: b.detach().cpu().clone() for a, b in main.state_dict().items() if not a.startswith("output_block.4")}
    est = {a: b.detach().cpu().clone() for a, b in den.get_model(ssdn.Denoiser.SIGMA_ESTIMATOR, False).state_dict().items()
           if not a.startswith("output_block.4")} if mode == "var" else None
    out = train_step(den, opt, data)
    torch.cuda.synchronize()
    for net in den._models.values():
        for plan in net._plans.values():
            plan.check()
            assert plan.scale_status()[:2] == (0, 0)
    loss = out[PipelineOutput.LOSS].detach().cpu().view(-1)
    k = 2

    def oracle(dt):
        cast = lambda t: t.to(dt) if torch.is_tensor(t) and t.is_floating_point() else t     # noqa: E731
        pp = {a: cast(b) for a, b in params.items()}
        with torch.no_grad():
            if algo == "ssdn":
                return O.ssdn_pipeline(pp, cast(noisy[:k]), cast(md[M.INPUT_NOISE_VALUES][:k]), mode, {a: cast(b) for a, b in est.items()} if est else None)
            return O.mask_mse_pipeline(pp, cast(noisy[:k]), cast(ref[:k]), md[M.MASK_COORDS][:k])
    o32, o64 = oracle(torch.float32), oracle(torch.float64)
    err = ((loss[:k].double() - o64["loss"].view(-1)).abs().max() / o64["loss"].abs().max()).item()
    err32 = ((o32["loss"].view(-1).double() - o64["loss"].view(-1)).abs().max() / o64["loss"].abs().max()).item()
    assert torch.isfinite(loss).all()
    assert err < max(TOL, 2 * err32), (err, err32)
    mu_key, okey = (PipelineOutput.IMG_MU, "mu") if algo == "ssdn" else (PipelineOutput.IMG_DENOISED, "out")
    assert rel(out[mu_key][:k], o32[okey]) < TOL
    first = float(loss.mean())
    for _ in range(3):
        out = train_step(den, opt, data)
    assert float(out[PipelineOutput.LOSS].mean()) < first


# ------------------------------------------------------------------------------------------ evaluation at full image size
@pytest.mark.parametrize("size", [512, 768])
def test_eval_large_images(engine, size):
    """SURVEY.md 8f rank 1 at the sizes of BSD (512) and Kodak (768): the trained checkpoint on one synthetic image,
    sigma = 25/255.  Output within 1e-4 of the oracle, PSNR within 1e-3 dB."""
    params, _ = _trained_params()
    den = ssdn.Denoiser(make_cfg("ssdn", "known", 3), device="cuda")
    den.get_model(ssdn.Denoiser.MODEL, False).load_state_dict(params, strict=False)
    g = torch.Generator().manual_seed(size)
    clean = F.interpolate(torch.rand(1, 3, size // 16, size // 16, generator=g), size=(size, size), mode="bicubic", align_corners=False).clamp(0, 1)
    noisy = clean + torch.randn(clean.shape, generator=g) * 25 / 255
    sigma = torch.full((1, 1, 1, 1), 25 / 255)
    den.eval()
    with torch.no_grad():
        out = den.run_pipeline([noisy, torch.zeros(0), {M.INPUT_NOISE_VALUES: sigma}])
        ref = O.ssdn_pipeline(params, noisy, sigma, "known")
    for net in den._models.values():
        for plan in net._plans.values():
            plan.check()
    pme = out[PipelineOutput.IMG_DENOISED].cpu()
    assert rel(out[PipelineOutput.IMG_MU], ref["mu"]) < TOL
    psnr = lambda a: float(-10.0 * torch.log10(((a - clean) ** 2).mean()))      # noqa: E731
    assert abs(psnr(pme) - psnr(ref["pme"])) < 1e-3, (psnr(pme), psnr(ref["pme"]))
    assert psnr(pme) > psnr(noisy) + 8.0


# ------------------------------------------------------------------------------------------ convergence
def _train_both(den, tr, steps, n, size, seed0, lr):
    """`steps` optimiser steps of the engine (eager for 3 steps, then one CUDA graph per step) and of the CPU oracle on the same
    seeded stream of batches (a fresh batch every step).  Returns the two loss curves."""
    opt = FlatAdam(den)
    opt.param_groups[0]["lr"] = lr
    tr.lr = lr
    sigma = torch.full((n, 1, 1, 1), 25 / 255)
    graphed, out = None, None
    le, lo = [], []
    for s in range(steps):
        clean, noisy = O.synthetic_batch(n, 3, size, seed=seed0 + s)
        data = [noisy.cuda(), torch.zeros(0), {M.INPUT_NOISE_VALUES: sigma.cuda()}]
        if s < 3:
            out = train_step(den, opt, data)
        elif graphed is None:
            out = None                                                  # (no live autograd graph of an eager step during capture)
            graphed = GraphedTrainStep(den, opt, data, warmup=1)        # runs this batch once itself ...
            out = graphed.outputs
            tr.step(noisy, sigma)                                       # ... after one eager warm-up step on it: the oracle takes it too
        else:
            out = graphed(data)
        le.append(float(out[PipelineOutput.LOSS].detach().mean()))
        lo.append(float(tr.step(noisy, sigma)["loss"].detach().mean()))
    for net in den._models.values():
        for plan in net._plans.values():
            plan.check()
            assert plan.scale_status()[2] <= 2                          # at most the calibration passes of a new plan were stale
    return torch.tensor(le), torch.tensor(lo)


def _heldout_psnr(den, tr, size):
    clean, noisy = O.synthetic_batch(32, 3, size, seed=99)
    sig = torch.full((32, 1, 1, 1), 25 / 255)
    den.eval()
    with torch.no_grad():
        pe = den.run_pipeline([noisy, torch.zeros(0), {M.INPUT_NOISE_VALUES: sig}])[PipelineOutput.IMG_DENOISED].cpu()
        po = O.ssdn_pipeline({k: v.detach() for k, v in tr.params.items()}, noisy, sig, "known")["pme"]
    psnr = lambda a: float(-10.0 * torch.log10(((a - clean) ** 2).mean()))      # noqa: E731
    return psnr(pe), psnr(po), psnr(noisy)


def _oracle_with(den):
    main = den.get_model(ssdn.Denoiser.MODEL, False)
    tr = O.CpuTrainer("ssdn", "known", 3, seed=0, lr=3e-4)
    with torch.no_grad():
        for kname, v in main.state_dict().items():
            if not kname.startswith("output_block.4"):
                tr.params[kname].copy_(v.cpu())
    return tr


def test_convergence_from_scratch_matches_oracle(engine):
    """SURVEY.md 8(d) "PSNR parity (ii)": 200 optimiser steps from the same fresh initialisation on the same stream of batches
    (8 patches of 32 x 32, lr 3e-4), engine against the CPU oracle.  Two floating-point implementations of a training run
    separate slowly (every LeakyReLU input within rounding of zero is a fork, Adam's first steps are sign-like), so what is
    compared is what training is for: the loss curve, window by window, to 1 % of its span, and the PSNR on a held-out batch -
    early in training, where the PSNR still moves by ~0.05 dB per step, to 0.3 dB (measured: 0.15 dB, profiles/r02_convergence.txt)."""
    torch.manual_seed(11)
    den = ssdn.Denoiser(make_cfg("ssdn", "known", 3), device="cuda")
    tr = _oracle_with(den)
    le, lo = _train_both(den, tr, 200, 8, 32, 5000, 3e-4)
    we, wo = le.view(-1, 20).mean(1), lo.view(-1, 20).mean(1)
    span = float(wo.max() - wo.min())
    pe, po, pin = _heldout_psnr(den, tr, 32)
    print("from scratch: loss windows engine", [round(v, 3) for v in we.tolist()], "oracle", [round(v, 3) for v in wo.tolist()],
          "held-out PSNR engine %.3f oracle %.3f input %.3f" % (pe, po, pin))
    assert (we - wo).abs().max().item() < 0.01 * span, (we.tolist(), wo.tolist())
    assert (le[:60] - lo[:60]).abs().max().item() < 0.01 * span          # step by step while the trajectories still coincide
    assert wo[-1] < wo[0] - 0.5 and we[-1] < we[0] - 0.5                  # both actually trained
    assert abs(pe - po) < 0.3, (pe, po)
    assert pe > pin + 3.0


def test_fine_tuning_the_trained_checkpoint_matches_oracle(engine):
    """The same comparison where a run is stable: 60 further steps from the reference's TRAINED checkpoint (lr 1e-4).  Near the
    optimum the two trajectories stay together: every per-step loss within 0.5 %, held-out PSNR within 0.05 dB and not worse
    than before the fine-tuning by more than 0.5 dB."""
    params, _ = _trained_params()
    den = ssdn.Denoiser(make_cfg("ssdn", "known", 3), device="cuda")
    den.get_model(ssdn.Denoiser.MODEL, False).load_state_dict(params, strict=False)
    tr = _oracle_with(den)
    p0 = _heldout_psnr(den, tr, 32)
    den.train()
    le, lo = _train_both(den, tr, 60, 8, 32, 7000, 1e-4)
    pe, po, pin = _heldout_psnr(den, tr, 32)
    print("fine-tuning: max per-step loss difference %.2e of |loss| %.3f; held-out PSNR engine %.3f oracle %.3f (before: %.3f) input %.3f"
          % ((le - lo).abs().max().item(), float(lo.abs().mean()), pe, po, p0[0], pin))
    assert ((le - lo).abs() / lo.abs()).max().item() < 5e-3
    assert abs(pe - po) < 0.05, (pe, po)
    assert pe > p0[0] - 0.5 and pe > pin + 8.0
