"""GPU: operator-level parity of the C-ABI against the CPU oracle (same seeded inputs)."""
import pytest
import torch

import cases as C
import ssdn_oracle as O
from util import TOL, rel

pytestmark = pytest.mark.gpu


def test_index_ops_bit_exact(engine):
    g = C.load_golden("index_ops")
    assert torch.equal(engine.rot4_stack(g["x"].cuda()).cpu(), g["rot4"])             # fixture = reference output
    assert torch.equal(engine.shift_unrot_concat(g["y"].cuda()).cpu(), g["unrot"])
    gen = torch.Generator().manual_seed(9)
    for n, c, s in ((1, 1, 32), (3, 3, 64), (2, 96, 32), (32, 3, 64), (5, 7, 2)):
        x = torch.rand(n, c, s, s, generator=gen)
        assert torch.equal(engine.rot4_stack(x.cuda()).cpu(), O.rot4_stack(x))
        y = torch.rand(4 * n, c, s, s, generator=gen)
        assert torch.equal(engine.shift_unrot_concat(y.cuda()).cpu(), O.shift_unrot_concat(y))
    with pytest.raises(ValueError):
        engine.rot4_stack(torch.rand(1, 1, 4, 6).cuda())


CONV_CASES = [  # n, cin, h, w, cout, k, blind
    (2, 48, 16, 16, 48, 3, True), (2, 96, 32, 32, 96, 3, True), (1, 3, 64, 64, 48, 3, True), (2, 144, 16, 16, 96, 3, False),
    (2, 99, 32, 32, 96, 3, True), (4, 48, 2, 2, 48, 3, True), (2, 384, 32, 32, 384, 1, False), (2, 96, 32, 32, 9, 1, False),
    (3, 1, 32, 32, 48, 3, False), (1, 96, 4, 4, 96, 3, True), (2, 96, 32, 64, 2, 1, False), (1, 144, 128, 128, 96, 3, True)]


@pytest.mark.parametrize("n,cin,h,w,cout,k,blind", CONV_CASES)
def test_conv_forward_dgrad_wgrad(engine, n, cin, h, w, cout, k, blind):
    gen = torch.Generator().manual_seed(n * 1000 + cin + h)
    x = torch.randn(n, cin, h, w, generator=gen)
    wt = torch.randn(cout, cin, k, k, generator=gen) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=gen)
    dy = torch.randn(n, cout, h, w, generator=gen)
    conv = O.shift_conv2d if blind else O.conv2d_same
    xg, wg, bg = x.double().requires_grad_(True), wt.double().requires_grad_(True), b.double().requires_grad_(True)
    z = conv(xg, wg, bg)
    z.backward(dy.double())
    y = engine.conv2d_forward(x.cuda(), wt.cuda(), b.cuda(), blind=blind, lrelu=True)
    assert rel(y, O.lrelu(z)) < 2e-5
    y0 = engine.conv2d_forward(x.cuda(), wt.cuda(), None, blind=blind, lrelu=False)
    assert rel(y0, conv(x.double(), wt.double(), None)) < 2e-5
    assert rel(engine.conv2d_backward_data(dy.cuda(), wt.cuda(), blind=blind), xg.grad) < 2e-5
    dw, db = engine.conv2d_backward_weight(x.cuda(), dy.cuda(), k, blind=blind)
    assert rel(dw, wg.grad) < 2e-5 and rel(db, bg.grad) < 2e-5


def test_conv_linearity_at_full_size(engine):
    """Size-independent property at the BASELINE batch (128 rotated images, 96 -> 96, 64 x 64): conv(a x + b y) = a conv(x) + b conv(y)."""
    gen = torch.Generator().manual_seed(4)
    x, y = torch.randn(128, 96, 64, 64, generator=gen).cuda(), torch.randn(128, 96, 64, 64, generator=gen).cuda()
    wt = (torch.randn(96, 96, 3, 3, generator=gen) / 30).cuda()
    f = lambda t: engine.conv2d_forward(t, wt, None, blind=True, lrelu=False)  # noqa: E731
    lhs, rhs = f(2.0 * x - 0.5 * y), 2.0 * f(x) - 0.5 * f(y)
    assert rel(lhs, rhs) < 2e-5
    assert f(x)[:, :, 0].abs().max() > 0 and torch.equal(f(torch.zeros_like(x)), torch.zeros_like(x))


def test_operators_refuse_cpu_tensors(engine):
    with pytest.raises(engine.EngineError):
        engine.conv2d_forward(torch.rand(1, 3, 32, 32), torch.rand(4, 3, 3, 3))
    with pytest.raises(engine.EngineError):
        engine.rot4_stack(torch.rand(1, 1, 4, 4))


@pytest.mark.parametrize("c,cs,known", [(3, 1, True), (3, 1, False), (3, 3, False), (1, 1, True), (1, 1, False)])
def test_posterior_forward_backward(engine, c, cs, known):
    gen = torch.Generator().manual_seed(20 + c + cs)
    n, h = 3, 32
    co = c + c * (c + 1) // 2
    net_out = torch.randn(n, co, h, h, generator=gen) * 0.3
    net_out[:, c:] += 0.5                                        # keep Sigma_x reasonably conditioned
    noisy = torch.rand(n, c, h, h, generator=gen)
    raw = (torch.rand(n, cs, 1, 1, generator=gen) * 0.2 + 0.05) if known else (torch.rand(n, cs, 1, 1, generator=gen) * 2 + 1.0)
    gloss = torch.rand(n, 1, generator=gen)
    no = net_out.double().requires_grad_(True)
    rw = raw.double().requires_grad_(True)
    sig = torch.max(rw, torch.tensor(1e-3, dtype=torch.float64)) if known else O.softplus_sigma(rw)
    ref = O.ssdn_posterior(no, noisy.double(), sig, known)
    (ref["loss"] * gloss.double()).sum().backward()
    pme, loss, mstd, nstd = engine.posterior_forward(net_out.cuda(), noisy.cuda(), raw.reshape(n, cs).cuda(), known)
    assert rel(loss, ref["loss"]) < 1e-5 and rel(pme, ref["pme"]) < 1e-4 and rel(mstd, ref["model_std"]) < 1e-4
    assert rel(nstd.reshape(-1), ref["noise_std"].reshape(-1)) < 1e-5
    dnet, dsig = engine.posterior_backward(net_out.cuda(), noisy.cuda(), raw.reshape(n, cs).cuda(), gloss.reshape(-1).cuda(), known)
    assert rel(dnet, no.grad) < 1e-4
    if not known:
        assert rel(dsig.reshape(-1), rw.grad.reshape(-1)) < 1e-4
    else:
        assert dsig is None


@pytest.mark.parametrize("c,cs,known", [(3, 1, True), (3, 1, False), (3, 3, False), (1, 1, True), (1, 1, False)])
def test_posterior_poisson_forward_backward(engine, c, cs, known):
    """Poisson noise model (denoiser.py:285-297): sigma = sqrt(max(mu, 1e-3) / lambda) or sqrt(max(mu, 1e-3) * estimate) per
    pixel; values, the per-pixel noise level and the gradients (through sigma into mu and into the estimate) vs autograd
    of the oracle formulation in fp64.  mu straddles the 1e-3 clamp."""
    gen = torch.Generator().manual_seed(40 + c + cs)
    n, h = 3, 32
    co = c + c * (c + 1) // 2
    net_out = torch.randn(n, co, h, h, generator=gen) * 0.3
    net_out[:, :c] = torch.rand(n, c, h, h, generator=gen) * 0.9 - 0.05      # means mostly in (0, 0.85), some below the clamp
    net_out[:, c:] += 0.5
    noisy = torch.rand(n, c, h, h, generator=gen)
    raw = (torch.rand(n, cs, 1, 1, generator=gen) * 40 + 10) if known else (torch.rand(n, cs, 1, 1, generator=gen) * 2 + 1.0)
    gloss = torch.rand(n, 1, generator=gen)
    no = net_out.double().requires_grad_(True)
    rw = raw.double().requires_grad_(True)
    base = torch.max(no[:, :c], torch.tensor(1e-3, dtype=torch.float64))
    sig = (base / rw) ** 0.5 if known else (base * O.softplus_sigma(rw)) ** 0.5
    ref = O.ssdn_posterior(no, noisy.double(), sig, known)
    (ref["loss"] * gloss.double()).sum().backward()
    pme, loss, mstd, nstd = engine.posterior_forward(net_out.cuda(), noisy.cuda(), raw.reshape(n, cs).cuda(), known, poisson=True)
    assert nstd.shape == (n, h, h)
    assert rel(loss, ref["loss"]) < 1e-5 and rel(pme, ref["pme"]) < 1e-4 and rel(mstd, ref["model_std"]) < 1e-4
    assert rel(nstd, ref["noise_std"]) < 1e-5
    dnet, dsig = engine.posterior_backward(net_out.cuda(), noisy.cuda(), raw.reshape(n, cs).cuda(), gloss.reshape(-1).cuda(), known, poisson=True)
    assert rel(dnet, no.grad) < 1e-4
    if not known:
        assert rel(dsig.reshape(-1), rw.grad.reshape(-1)) < 1e-4
    else:
        assert dsig is None


def test_mse_and_masked_mse(engine):
    gen = torch.Generator().manual_seed(31)
    a, b = torch.rand(4, 3, 32, 32, generator=gen), torch.rand(4, 3, 32, 32, generator=gen)
    gloss = torch.rand(4, generator=gen)
    ag = a.double().requires_grad_(True)
    ref = ((ag - b.double()) ** 2).reshape(4, -1).mean(1, keepdim=True)
    (ref.reshape(-1) * gloss.double()).sum().backward()
    assert rel(engine.mse_forward(a.cuda(), b.cuda()), ref) < 1e-5
    assert rel(engine.mse_backward(a.cuda(), b.cuda(), gloss.cuda()), ag.grad) < 1e-5
    coords = torch.randint(0, 32, (4, 64, 2), generator=gen)
    coords[0, 5] = coords[0, 4]                                   # duplicates accumulate
    ag = a.double().requires_grad_(True)
    ref = O.masked_mse(coords, ag, b.double())
    (ref.reshape(-1) * gloss.double()).sum().backward()
    c0 = coords[0].cuda().contiguous()
    assert rel(engine.masked_mse_forward(a.cuda(), b.cuda(), c0), ref) < 1e-5
    assert rel(engine.masked_mse_backward(a.cuda(), b.cuda(), c0, gloss.cuda()), ag.grad) < 1e-5


def test_spatial_mean(engine):
    x = torch.rand(5, 1, 64, 64)
    assert rel(engine.spatial_mean_forward(x.cuda()), x.mean(dim=(2, 3), keepdim=True)) < 1e-6
    g = torch.rand(5, 1, 1, 1)
    assert rel(engine.spatial_mean_backward(g.cuda(), x.shape), (g / 4096).expand_as(x)) < 1e-6


def test_adam_matches_torch_optim(engine):
    g = C.load_golden("optimiser")
    gen = torch.Generator().manual_seed(5)
    p = torch.randn(1000, generator=gen).cuda()
    m, v = torch.zeros(1000).cuda(), torch.zeros(1000).cuda()
    for k in range(4):
        gr = (torch.randn(1000, generator=gen) * (10.0 ** (k - 2))).cuda()
        engine.adam_step(p, gr, m, v, 3e-4 * (k + 1), k + 1)
        assert rel(p, g["adam_traj"][k]) < 1e-6                   # fixture = torch.optim.Adam as train.py:107 configures it
    p2, m2, v2 = p.clone(), m.clone(), v.clone()
    engine.adam_step(p, 2 * gr, m, v, 1e-3, 5, grad_scale=0.5)
    engine.adam_step(p2, gr, m2, v2, 1e-3, 5)
    assert torch.equal(p, p2)
