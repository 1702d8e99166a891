"""CPU: the oracle reproduces every golden fixture (= outputs of the unmodified reference, tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

import cases as C
import ssdn_oracle as O
from util import oracle_case, rel

torch.set_num_threads(max(1, min(8, torch.get_num_threads())))


def test_index_ops_fixture():
    g = C.load_golden("index_ops")
    assert torch.equal(O.rot4_stack(g["x"]), g["rot4"])
    assert torch.equal(O.shift_unrot_concat(g["y"]), g["unrot"])
    for a in (0, 90, 180, 270):
        assert np.array_equal(O.rotate(g["x"], a).numpy(), O.rotate_np(g["x"].numpy(), a))
    for a, b in ((0, 0), (90, 270), (180, 180), (270, 90)):
        assert torch.equal(O.rotate(O.rotate(g["x"], a), b), g["x"])


def test_shift2d_is_zero_filled_roll():
    x = torch.arange(2 * 1 * 4 * 5, dtype=torch.float32).reshape(2, 1, 4, 5)
    s = O.shift2d(x, 1, 0)
    assert torch.equal(s[:, :, 1:], x[:, :, :-1]) and s[:, :, 0].abs().sum() == 0
    s = O.shift2d(x, 0, -2)
    assert torch.equal(s[..., :-2], x[..., 2:]) and s[..., -2:].abs().sum() == 0


def test_shiftconv_sees_rows_h_minus_2_to_h():
    """Delta weights: tap (kh, kw) of a ShiftConv2d copies input pixel (h + kh - 2, w + kw - 1)."""
    x = torch.rand(1, 1, 6, 6)
    for kh in range(3):
        for kw in range(3):
            w = torch.zeros(1, 1, 3, 3)
            w[0, 0, kh, kw] = 1
            assert torch.equal(O.shift_conv2d(x, w, None), O.shift2d(x, 2 - kh, 1 - kw))


@pytest.mark.parametrize("name", list(C.NETWORK_CASES))
def test_network_fixture(name):
    cin, cout, blind, n, size = C.NETWORK_CASES[name]
    params, x, dout = C.network_inputs(name)
    gold = C.load_golden(name)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    xg = x.clone().requires_grad_(True)
    out = O.noise_network_forward(p, xg, blind)
    out.backward(dout)
    assert rel(out, gold["out"]) < 1e-6
    assert rel(xg.grad, gold["dx"]) < 5e-5
    grads = {k: p[k].grad for k in O.param_order(cin, cout, blind)}
    assert rel(C.grad_summary(grads), gold["grad_summary"]) < 5e-5
    assert rel(grads["encode_block_1.0.weight"], gold["g_first_w"]) < 5e-5
    assert rel(grads["output_conv.weight"], gold["g_out_w"]) < 5e-5


def test_blindspot_property():
    """d out[:, :, h, w] / d in[:, :, h, w] == 0 exactly for the blind-spot network."""
    params, x, _ = C.network_inputs("net_blind_rgb")
    xg = x[:1].clone().requires_grad_(True)
    out = O.noise_network_forward(params, xg, True)
    out[0, :, 13, 17].sum().backward()
    assert xg.grad[0, :, 13, 17].abs().max() == 0
    assert (xg.grad[0].abs().sum(0) > 0).sum() > 900       # ... while almost every other pixel is seen


@pytest.mark.parametrize("name", list(C.PIPELINE_CASES))
def test_pipeline_fixture(name):
    d = C.pipeline_inputs(name)
    gold = C.load_golden(name)
    out, g, ge, gs = oracle_case(d)
    assert rel(out["loss"], gold["loss"]) < 1e-6
    names = list(d["params"].keys())
    assert rel(C.grad_summary({k: g[k] for k in names}), gold["grad_summary"]) < 1e-4
    if d["algorithm"] == "ssdn":
        assert rel(out["pme"], gold["out"]) < 1e-5 and rel(out["mu"], gold["mu"]) < 1e-6
        assert rel(out["model_std"], gold["model_std"]) < 1e-5
        assert rel(out["noise_std"].reshape(-1), gold["noise_std"].reshape(-1)) < 1e-6
    else:
        assert rel(out["out"], gold["out"]) < 1e-6
    if gs is not None:
        assert rel(gs, gold["g_est_sigma"]) < 1e-5
    if ge is not None:
        assert rel(ge["output_conv.weight"], gold["g_est_out_w"]) < 1e-4
    assert rel(O.psnr(out["pme"] if "pme" in out else out["out"], d["clean"]), gold["psnr"]) < 1e-5


def test_posterior_limits():
    """Known answers: Sigma_x -> 0 gives pme -> mu; Sigma_x -> inf gives pme -> y; diagonal RGB equals 3 x mono."""
    n, h = 1, 4
    mu, y = torch.rand(n, 3, h, h), torch.rand(n, 3, h, h)
    sig = torch.full((n, 1, 1, 1), 0.1)
    small = torch.cat([mu, torch.zeros(n, 6, h, h)], 1)
    assert rel(O.ssdn_posterior(small, y, sig, True)["pme"], mu) < 1e-3
    a = torch.zeros(n, 6, h, h)
    a[:, [0, 3, 5]] = 300.0
    assert rel(O.ssdn_posterior(torch.cat([mu, a], 1), y, sig, True)["pme"], y) < 1e-3
    a[:, [0, 3, 5]] = torch.rand(n, 3, h, h) + 0.1
    rgb = O.ssdn_posterior(torch.cat([mu, a], 1).double(), y.double(), sig.double(), True)
    for c, k in enumerate((0, 3, 5)):
        mono = O.ssdn_posterior(torch.cat([mu[:, c:c + 1], a[:, k:k + 1]], 1).double(), y[:, c:c + 1].double(), sig.double(), True)
        assert rel(rgb["pme"][:, c:c + 1], mono["pme"]) < 1e-4


def test_optimiser_fixture():
    g = C.load_golden("optimiser")
    its = 2_000_000
    for i, v in zip(g["lr_points"].tolist(), g["lr_values"].tolist()):
        assert abs(O.effective_lrate(int(i), its) - v) < 1e-12
    # effective schedule: up over the first 10 %, flat, down over the last 30 %
    assert O.effective_lrate(0, its) == 0 and abs(O.effective_lrate(its // 20, its) - 1.5e-4) < 1e-9
    assert abs(O.effective_lrate(its // 2, its) - 3e-4) < 1e-12 and abs(O.effective_lrate(int(its * 0.85), its) - 7.5e-5) < 1e-9
    gen = torch.Generator().manual_seed(5)
    p = torch.randn(1000, generator=gen)
    m, v = torch.zeros(1000), torch.zeros(1000)
    for k in range(4):
        gr = torch.randn(1000, generator=gen) * (10.0 ** (k - 2))
        O.adam_step(p, gr, m, v, k + 1, 3e-4 * (k + 1))
        assert rel(p, g["adam_traj"][k]) < 1e-6


def test_training_trajectory_fixture():
    d = C.pipeline_inputs("ssdn_known_rgb")
    gold = C.load_golden("trajectory_ssdn_known_rgb")
    tr = O.CpuTrainer("ssdn", "known", 3)
    tr.params = {k: v.clone().requires_grad_(True) for k, v in d["params"].items()}
    tr.leaves = list(tr.params.values())
    tr.m = [torch.zeros_like(t) for t in tr.leaves]
    tr.v = [torch.zeros_like(t) for t in tr.leaves]
    for k in range(3):
        out = tr.step(d["noisy"], d["noise_values"], lr=3e-4)
        assert rel(out["loss"], gold["losses"][k]) < 5e-5


def _load_wt_golden():
    import numpy as np
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wt_ssdn_gauss25_sigma_known.npz"))
    params = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p.")}
    return params, {k: torch.from_numpy(z[k]) for k in z.files if not k.startswith("p.")}


def _psnr(a, b):
    return -10.0 * torch.log10(((a - b) ** 2).mean(dim=(1, 2, 3)))


def test_oracle_reproduces_trained_checkpoint_outputs():
    """Trained reference weights (models/final-ssdn-gauss25-sigma_known.wt, tests/golden/make_wt_golden.py): the oracle
    reproduces the reference's posterior mean, network mean, loss and PSNRs (20.3 dB in -> 30.2 dB out)."""
    params, g = _load_wt_golden()
    out = O.ssdn_pipeline(params, g["noisy"], g["sigma"], "known")
    for k in ("pme", "mu"):
        assert ((out[k] - g[k]).abs().max() / g[k].abs().max()).item() < 1e-5, k
    assert ((out["loss"].view(-1) - g["loss"]).abs().max() / g["loss"].abs().max()).item() < 1e-5
    assert (_psnr(out["pme"], g["clean"]) - g["psnr_pme"]).abs().max().item() < 1e-3
    assert (g["psnr_pme"] > g["psnr_in"] + 9.0).all()          # the checkpoint really denoises: ~ +10 dB
