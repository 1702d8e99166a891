"""GPU: Denoiser pipelines through the public Python API against the reference fixtures and the oracle."""
import pytest
import torch

import cases as C
import ssdn
import ssdn_oracle as O
from ssdn.datasets import NoisyDataset
from ssdn.params import PipelineOutput
from ssdn.train import FlatAdam, train_step
from util import TOL, as_accurate_as_reference, data_for_case, denoiser_for_case, make_cfg, oracle_case, rel, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(C.PIPELINE_CASES))
def test_pipeline_outputs_and_gradients(engine, name):
    d = C.pipeline_inputs(name)
    gold = C.load_golden(name)
    den = denoiser_for_case(d)
    out = den.run_pipeline(data_for_case(d))
    out[PipelineOutput.LOSS].mean().backward()
    torch.cuda.synchronize()
    o32, g32, ge32, gs32 = oracle_case(d, torch.float32)
    o64, g64, ge64, gs64 = oracle_case(d, torch.float64)
    # well-conditioned outputs: straight against the reference fixture
    assert rel(out[PipelineOutput.LOSS], gold["loss"]) < TOL
    assert out[PipelineOutput.LOSS].shape == gold["loss"].shape
    if d["algorithm"] == "ssdn":
        assert rel(out[PipelineOutput.IMG_MU], gold["mu"]) < TOL
        if d.get("noise_style", "gauss").startswith("poisson"):
            # sigma = sqrt(max(mu, 1e-3) * k) per pixel: the 4e-5 forward error of mu is amplified by 1 / (2 mu) where the
            # (untrained) mean is small - held to the exact result within the fp32 reference's own band, floor 5e-4
            ok, errs = as_accurate_as_reference(out[PipelineOutput.NOISE_STD_DEV], o32["noise_std"], o64["noise_std"], slack=1.0, floor=5e-4)
            assert ok, errs
        else:
            assert rel(out[PipelineOutput.NOISE_STD_DEV].reshape(-1), gold["noise_std"].reshape(-1)) < TOL
        assert out[PipelineOutput.NOISE_STD_DEV].shape == gold["noise_std"].shape
        assert out[PipelineOutput.MODEL_STD_DEV].shape == gold["model_std"].shape
        # posterior mean / model std invert or take the determinant of Sigma_x, near-singular for an untrained net:
        # the fp32 reference is itself only accurate to e_ref there (LAPACK LU in fp32), the engine works in fp64 registers
        for key, okey in ((PipelineOutput.IMG_DENOISED, "pme"), (PipelineOutput.MODEL_STD_DEV, "model_std")):
            # floor 5e-4: with an estimated sigma the posterior mean amplifies the ~4e-5 forward error of both networks
            # (Poisson: sigma itself depends on the small means of the untrained network - one more amplification: slack 3)
            slack = 3.0 if d.get("noise_style", "gauss").startswith("poisson") else 1.0
            ok, errs = as_accurate_as_reference(out[key], o32[okey], o64[okey], slack=slack, floor=5e-4)
            assert ok, (key, errs)
    else:
        assert rel(out[PipelineOutput.IMG_DENOISED], gold["out"]) < TOL
    psnr = ssdn.utils.calculate_psnr(out[PipelineOutput.IMG_DENOISED].detach(), d["clean"].cuda())
    assert rel(psnr, O.psnr(o64["pme"] if "pme" in o64 else o64["out"], d["clean"].double())) < 1e-3
    # gradients end to end: the kernels are exact to 1e-4 on identical activations (test_gpu_network); with the engine's own
    # activations a handful of LeakyReLU derivatives of ~zero pre-activations differ (see test_end_to_end_gradients_and_mask_flips),
    # which bounds the agreement at a few 1e-3 in relative L2.  The last layers (no mask below them) must match to 1e-3.
    main = den.get_model(ssdn.Denoiser.MODEL, False)
    for k, p in main.named_parameters():
        assert rel_l2(p.grad, g32[k]) < 2e-2, k
    # Poisson: d(sigma^2)/d(mu) = k * [mu > 1e-3] and the 1 / (2 sigma) regulariser term amplify the forward error of the
    # small means of an untrained network (the kernels themselves match autograd to 1e-4 on identical inputs:
    # test_posterior_poisson_forward_backward)
    tol_w, tol_b = (1e-2, 1e-2) if d.get("noise_style", "gauss").startswith("poisson") else (2e-3, 1e-3)
    assert rel(dict(main.named_parameters())["output_conv.weight"].grad, gold["g_out_w"]) < tol_w
    assert rel(dict(main.named_parameters())["output_conv.bias"].grad, gold["g_out_b"]) < tol_b
    if ge64 is not None:
        est = den.get_model(ssdn.Denoiser.SIGMA_ESTIMATOR, False)
        for k, p in est.named_parameters():
            assert rel_l2(p.grad, ge32[k]) < 2e-2, ("estimator", k)
    if gs64 is not None:
        assert rel(den.l_params[ssdn.Denoiser.ESTIMATED_SIGMA].grad, gold["g_est_sigma"]) < (1e-3 if d.get("noise_style", "gauss").startswith("poisson") else TOL)


def test_training_trajectory_matches_reference(engine):
    """Three optimiser steps (train.py:197-202) from the fixture weights: per-sample losses of every step."""
    d = C.pipeline_inputs("ssdn_known_rgb")
    gold = C.load_golden("trajectory_ssdn_known_rgb")
    den = denoiser_for_case(d)
    opt = FlatAdam(den)
    opt.param_groups[0]["lr"] = 3e-4
    for k in range(3):
        out = train_step(den, opt, data_for_case(d))
        # step 0 is a pure forward; later steps see Adam's first updates, which are lr * g / (|g| + eps) ~ lr * sign(g):
        # parameters whose gradient is ~0 move by +-lr depending on rounding noise, in any implementation
        assert rel(out[PipelineOutput.LOSS], gold["losses"][k]) < (TOL if k == 0 else 1e-2), k
    main = den.get_model(ssdn.Denoiser.MODEL, False)
    final = {k: p.data.cpu() for k, p in main.named_parameters()}
    assert rel_l2(C.grad_summary(final)[:, 1], gold["final_summary"][:, 1]) < 1e-3     # parameter norms after 3 steps


def test_train_step_reduces_loss_at_baseline_size(engine):
    """BASELINE config 2 shape (32 x 3 x 64 x 64, sigma known): loss goes down over a few steps and stays finite."""
    torch.manual_seed(0)
    den = ssdn.Denoiser(make_cfg("ssdn", "known"), device="cuda")
    opt = FlatAdam(den)
    opt.param_groups[0]["lr"] = 3e-4
    clean, noisy = O.synthetic_batch(32, 3, 64, seed=1234)
    M = NoisyDataset.Metadata
    data = [noisy, torch.zeros(0), {M.INPUT_NOISE_VALUES: torch.full((32, 1, 1, 1), 25 / 255), M.CLEAN: clean}]
    losses = [float(train_step(den, opt, data)[PipelineOutput.LOSS].detach().mean()) for _ in range(6)]
    assert all(map(lambda v: v == v and abs(v) < 1e6, losses)) and losses[-1] < losses[0]
    # first step equals the oracle on the same initial weights
    torch.manual_seed(0)
    ref = O.CpuTrainer("ssdn", "known", 3, seed=0)
    assert abs(losses[0] - float(ref.loss(noisy, data[2][M.INPUT_NOISE_VALUES])["loss"].mean())) < 1e-4 * abs(losses[0]) + 1e-5


def test_inference_forward_and_eval_mode(engine):
    d = C.pipeline_inputs("n2c_mono")
    den = denoiser_for_case(d)
    den.eval()
    out = den(d["noisy"])
    assert rel(out, C.load_golden("n2c_mono")["out"]) < TOL
    s = C.pipeline_inputs("ssdn_known_rgb")
    den = denoiser_for_case(s)
    with pytest.raises(ValueError):
        den(s["noisy"])
    pme = den(s["noisy"], s["noise_values"])
    assert pme.shape == s["noisy"].shape and torch.isfinite(pme).all()


def test_gradients_are_written_into_the_flat_buffer(engine):
    d = C.pipeline_inputs("ssdn_const_rgb")
    den = denoiser_for_case(d)
    den.run_pipeline(data_for_case(d))[PipelineOutput.LOSS].mean().backward()
    flat_g = den.flat_gradients()
    off = 0
    for p in den.parameters():
        assert torch.equal(flat_g[off:off + p.numel()], p.grad.reshape(-1))
        off += p.numel()
    assert off == flat_g.numel() == 1269130


def test_trained_checkpoint_psnr_parity(engine):
    """PSNR parity at TRAINED weights: the engine, loaded with the reference checkpoint
    final-ssdn-gauss25-sigma_known.wt, reproduces the reference's outputs to 1e-4 and its PSNRs to 1e-3 dB."""
    import os
    import numpy as np
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wt_ssdn_gauss25_sigma_known.npz"))
    params = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p.")}
    g = {k: torch.from_numpy(z[k]) for k in z.files if not k.startswith("p.")}
    den = ssdn.Denoiser(make_cfg("ssdn", "known", 3), device="cuda")
    den.get_model(ssdn.Denoiser.MODEL, False).load_state_dict(params, strict=False)
    den.eval()
    M = NoisyDataset.Metadata
    with torch.no_grad():
        out = den.run_pipeline([g["noisy"], torch.zeros(0), {M.INPUT_NOISE_VALUES: g["sigma"], M.CLEAN: g["clean"]}])
    psnr = lambda a, b: -10.0 * torch.log10(((a - b) ** 2).mean(dim=(1, 2, 3)))  # noqa: E731
    pme, mu = out[PipelineOutput.IMG_DENOISED].cpu(), out[PipelineOutput.IMG_MU].cpu()
    assert rel(mu, g["mu"]) < TOL
    # the posterior mean inverts Sigma_x + 1e-6 I: the fp32 reference itself is only accurate to its own rounding there,
    # so the engine (closed form in fp64 registers) is held to the exact (fp64 oracle) result within the reference's band
    ref64 = O.ssdn_pipeline({k: v.double() for k, v in params.items()}, g["noisy"].double(), g["sigma"].double(), "known")
    ok, errs = as_accurate_as_reference(pme, g["pme"], ref64["pme"])
    assert ok, errs
    assert rel(out[PipelineOutput.LOSS].view(-1), g["loss"]) < TOL
    assert (psnr(pme, g["clean"]) - g["psnr_pme"]).abs().max().item() < 1e-3
    assert (psnr(mu, g["clean"]) - g["psnr_mu"]).abs().max().item() < 1e-3


def test_evaluator_on_padded_full_images(engine):
    """Evaluation path (eval.py:19-127, train.py:243-259, 553-584): non-square images are
    reflect-padded to a square multiple of 32, denoised with the trained checkpoint, cropped back and scored.  PSNRs on the
    unpadded region must equal the oracle's to 1e-3 dB."""
    import os
    import numpy as np
    from ssdn.eval import DenoiserEvaluator
    from ssdn.params import NoiseAlgorithm
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wt_ssdn_gauss25_sigma_known.npz"))
    params = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p.")}
    state = {"cfg": make_cfg("ssdn", "known", 3)}
    state.update({"models.denoiser_model.module." + k: v for k, v in params.items()})
    ev = DenoiserEvaluator(state, device="cuda")
    torch.manual_seed(77)
    images = [O.synthetic_batch(1, 3, 96, seed=s)[0][0][:, :h, :w].contiguous() for s, (h, w) in ((1, (40, 56)), (2, (96, 70)), (3, (64, 64)))]
    ds = NoisyDataset([(im,) for im in images], "gauss25", NoiseAlgorithm.SELFSUPERVISED_DENOISING, pad_multiple=32, square=True)
    batches, expect = [], []
    for i in range(len(ds)):
        inp, ref, md = ds[i]
        md = {k: (v[None] if torch.is_tensor(v) else v) for k, v in md.items()}
        assert inp.shape[-1] == inp.shape[-2] and inp.shape[-1] % 32 == 0
        batches.append([inp[None], ref, md])
        sigma = md[NoisyDataset.Metadata.INPUT_NOISE_VALUES].reshape(1, -1, 1, 1)
        out = O.ssdn_pipeline(params, inp[None], sigma, "known")
        h, w = images[i].shape[1:]
        expect.append(float(-10.0 * torch.log10(((out["pme"][0, :, :h, :w] - images[i]) ** 2).mean())))
    got = ev.evaluate(batches)
    assert abs(got["psnr_out"] - sum(expect) / len(expect)) < 1e-3, (got, expect)
    assert got["psnr_out"] > 28.0


def test_forward_at_256_uses_row_windows(engine):
    """Full-size evaluation geometry: at 256 x 256 the 3x3 halo window no longer fits in shared memory as one box set and
    the plan falls back to one window per stencil row; outputs must still match the oracle."""
    p = O.init_params(3, 9, True, generator=torch.Generator().manual_seed(5))
    x = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(6))
    plan = engine.NetPlan(1, 3, 9, 256, 256, True, "cuda")
    flat = torch.cat([p[k].reshape(-1) for k in O.param_order(3, 9, True)]).cuda()
    out = plan.forward(flat, x.cuda(), training=False)
    plan.check()
    assert rel(out, O.noise_network_forward(p, x, True)) < TOL


def test_trainer_snapshot_and_resume_continue_the_run(engine, tmp_path):
    """SURVEY.md 8f rank 3 on the device: two steps, snapshot (reference ``.training`` layout, tests/test_training_wire.py),
    ``resume_run``, one more step == three uninterrupted steps.  The Adam moments travel through the per-parameter
    torch.optim.Adam layout and back into the flat buffers; the learning rate follows the restored image counter."""
    from ssdn.params import ConfigValue, HistoryValue, StateValue
    from ssdn.train import DenoiserTrainer, resume_run
    M = NoisyDataset.Metadata
    batches = []
    for s in range(3):
        clean, noisy = O.synthetic_batch(8, 3, 32, seed=900 + s)
        batches.append([noisy, torch.zeros(0), {M.CLEAN: clean, M.INPUT_NOISE_VALUES: torch.full((8, 1, 1, 1), 25 / 255)}])

    def fresh(run):
        cfg = make_cfg("ssdn", "const")
        cfg[ConfigValue.TRAIN_ITERATIONS] = 80          # LR: 0 at image 0, 3e-4 from image 8 on (ramp-up over the first 10 %)
        torch.manual_seed(3)
        t = DenoiserTrainer(cfg, runs_dir=str(tmp_path), run_dir=run)
        t.new_target(device="cuda")
        return t

    losses = {}
    straight = fresh("straight")
    straight.train(batches, on_step=lambda it, out: losses.setdefault(("straight", it), out[PipelineOutput.LOSS].detach().mean().item()))
    first = fresh("interrupted")
    first.train(batches[:2])
    assert first.state[StateValue.ITERATION] == 16 and first.learning_rate == pytest.approx(3e-4)
    path = first.snapshot()
    assert path.endswith("model_00000016.training")
    del first
    resumed = resume_run(str(tmp_path / "interrupted"), device="cuda")
    assert resumed.state[StateValue.ITERATION] == 16 and resumed._optimizer.step_count == 2
    assert resumed.learning_rate == pytest.approx(3e-4)
    resumed.train(batches[2:], on_step=lambda it, out: losses.setdefault(("resumed", it), out[PipelineOutput.LOSS].detach().mean().item()))
    torch.cuda.synchronize()
    assert resumed.state[StateValue.ITERATION] == straight.state[StateValue.ITERATION] == 24
    assert resumed._optimizer.step_count == straight._optimizer.step_count == 3
    assert resumed.state[StateValue.HISTORY][HistoryValue.TRAIN]["n"] == 24
    assert abs(losses[("resumed", 24)] - losses[("straight", 24)]) < 1e-5 * abs(losses[("straight", 24)]) + 1e-6
    # the resumed process starts with freshly calibrated operand scales (the straight run's third step uses the scales its
    # second step left behind): same arithmetic, different roundings of the fp16 planes, and a handful of LeakyReLU inputs
    # within rounding of zero flip - the moments agree to a few 1e-4 (measured 1.5e-4), the losses and weights to 1e-5 / 1e-4
    assert rel_l2(resumed._optimizer.exp_avg, straight._optimizer.exp_avg) < 5e-4
    assert rel_l2(resumed._optimizer.exp_avg_sq, straight._optimizer.exp_avg_sq) < 5e-4
    for (k, a), b in zip(resumed.denoiser.named_parameters(), straight.denoiser.parameters()):
        if a.dim() > 1 and a.numel() > 1:
            assert rel_l2(a, b) < 1e-4, k
    a = float(resumed.state[StateValue.HISTORY][HistoryValue.TRAIN]["loss"].accumulated())
    b = float(straight.state[StateValue.HISTORY][HistoryValue.TRAIN]["loss"].accumulated())
    assert abs(a - b) < 1e-5 * abs(b) + 1e-6


def test_graphed_step_with_pinned_host_batches_equals_the_eager_step(engine):
    """GraphedTrainStep: a different PINNED HOST batch every call (staged on the copy stream into alternating input slots
    while the previous step computes) gives the losses and the weights of the eager train_step on the same sequence, and
    loss_host() - written by the graph's own last node - is the step's per-sample loss."""
    from ssdn.train import GraphedTrainStep
    M = NoisyDataset.Metadata
    n, size, steps = 4, 32, 7
    sigma = torch.full((n, 1, 1, 1), 25 / 255)
    batches = [O.synthetic_batch(n, 3, size, seed=300 + s)[1] for s in range(steps)]
    cfg = make_cfg("ssdn", "known", 3)

    def fresh():
        torch.manual_seed(0)
        den = ssdn.Denoiser(cfg, device="cuda")
        opt = FlatAdam(den)
        opt.param_groups[0]["lr"] = 3e-4
        return den, opt

    den_e, opt_e = fresh()
    eager = []
    for b in batches:
        out = train_step(den_e, opt_e, [b.cuda(), torch.zeros(0), {M.INPUT_NOISE_VALUES: sigma.cuda()}])
        eager.append(out[PipelineOutput.LOSS].detach().cpu().clone())
        del out
    den_g, opt_g = fresh()
    host = lambda b: [b.pin_memory(), torch.zeros(0), {M.INPUT_NOISE_VALUES: sigma.pin_memory()}]     # noqa: E731
    graphed, got = None, []
    for s, b in enumerate(batches):
        if s == 0:
            out = train_step(den_g, opt_g, [b.cuda(), torch.zeros(0), {M.INPUT_NOISE_VALUES: sigma.cuda()}])     # step 1 (eager)
            got.append(out[PipelineOutput.LOSS].detach().cpu().clone())
            del out
        elif graphed is None:
            graphed = GraphedTrainStep(den_g, opt_g, host(b), warmup=1)     # runs this batch as steps 2 and 3 (warm-up + first replay)
            got.append(None)
        else:
            out = graphed(host(b))
            torch.cuda.synchronize()
            assert torch.equal(graphed.loss_host(), out[PipelineOutput.LOSS].detach().cpu())
            got.append(graphed.loss_host().clone())
    # the eager reference took batch 1 once where the graphed run took it twice: replay that on the eager side for the weights
    den_r, opt_r = fresh()
    seq = [batches[0], batches[1], batches[1]] + batches[2:]
    ref = []
    for b in seq:
        out = train_step(den_r, opt_r, [b.cuda(), torch.zeros(0), {M.INPUT_NOISE_VALUES: sigma.cuda()}])
        ref.append(out[PipelineOutput.LOSS].detach().cpu().clone())
        del out
    torch.cuda.synchronize()
    assert opt_g.step_count == opt_r.step_count == steps + 1
    for s in range(2, steps):
        assert rel(got[s], ref[s + 1]) < 1e-5, s                   # same data, same weights history: rounding only (operand scales)
    assert rel_l2(den_g.flat_parameters(), den_r.flat_parameters()) < 1e-5
    assert rel(eager[0], got[0]) < 1e-6
