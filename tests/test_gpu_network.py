"""GPU: whole-network parity (forward vs reference fixtures; backward kernel-exact on identical masks; end to end)."""
import pytest
import torch

import cases as C
import ssdn_oracle as O
from oracle_trace import oracle_trace, upload_activations
from util import TOL, as_accurate_as_reference, rel, rel_l2

pytestmark = pytest.mark.gpu


def _flat(params, order):
    return torch.cat([params[k].reshape(-1) for k in order]).cuda()


def _split(grads, params, order):
    out, off = {}, 0
    for k in order:
        n = params[k].numel()
        out[k] = grads[off:off + n].cpu().reshape(params[k].shape)
        off += n
    return out


@pytest.mark.parametrize("name", list(C.NETWORK_CASES))
def test_forward_matches_reference_fixture(engine, name):
    cin, cout, blind, n, size = C.NETWORK_CASES[name]
    params, x, _ = C.network_inputs(name)
    plan = engine.NetPlan(n, cin, cout, size, size, blind, "cuda")
    out = plan.forward(_flat(params, O.param_order(cin, cout, blind)), x.cuda(), training=False)
    plan.check()
    assert rel(out, C.load_golden(name)["out"]) < TOL


@pytest.mark.parametrize("name", list(C.NETWORK_CASES))
def test_backward_kernels_on_oracle_activations(engine, name):
    """Every backward kernel (dgrad, wgrad, bias, pool / upsample / un-rotate backward) against autograd of the oracle,
    with the oracle's forward activations uploaded so that LeakyReLU masks and pool arg-maxes are identical."""
    cin, cout, blind, n, size = C.NETWORK_CASES[name]
    params, x, dout = C.network_inputs(name)
    order = O.param_order(cin, cout, blind)
    flat = _flat(params, order)
    plan = engine.NetPlan(n, cin, cout, size, size, blind, "cuda")
    plan.forward(flat, x.cuda(), training=True)
    po = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    oo, T = oracle_trace(po, x, blind)
    oo.backward(dout)
    upload_activations(plan, T)
    grads = _split(plan.backward(flat, dout.cuda()), params, order)
    plan.check()
    for k in order:
        assert rel(grads[k], po[k].grad) < TOL, k
    gold = C.load_golden(name)                                   # and against the reference's own gradients
    assert rel(grads["output_conv.weight"], gold["g_out_w"]) < TOL and rel(grads["encode_block_1.0.weight"], gold["g_first_w"]) < TOL
    assert rel(C.grad_summary(grads), gold["grad_summary"]) < TOL


@pytest.mark.parametrize("name", ["net_blind_rgb", "net_plain_rgb"])
def test_end_to_end_gradients_as_accurate_as_reference(engine, name):
    """Engine forward + backward with its own activations.  A pre-activation within rounding error of zero flips a
    LeakyReLU derivative (x10) in ANY fp32 implementation, so the yardstick is the fp64 oracle: the engine must be as
    close to it as the fp32 reference is."""
    cin, cout, blind, n, size = C.NETWORK_CASES[name]
    params, x, dout = C.network_inputs(name)
    order = O.param_order(cin, cout, blind)
    flat = _flat(params, order)
    plan = engine.NetPlan(n, cin, cout, size, size, blind, "cuda")
    plan.forward(flat, x.cuda(), training=True)
    grads = _split(plan.backward(flat, dout.cuda()), params, order)
    plan.check()
    g32, g64 = {}, {}
    for dt, store in ((torch.float32, g32), (torch.float64, g64)):
        po = {k: v.to(dt).clone().requires_grad_(True) for k, v in params.items()}
        O.noise_network_forward(po, x.to(dt), blind).backward(dout.to(dt))
        store.update({k: po[k].grad for k in order})
    worst = 0.0
    for k in order:
        ok, (e_eng, e_ref) = as_accurate_as_reference(grads[k], g32[k], g64[k], slack=4.0, norm=rel_l2)
        worst = max(worst, e_eng)
        assert ok, (k, e_eng, e_ref)
    assert worst < 5e-2


def test_blindspot_property_on_engine(engine):
    """out[:, :, h, w] must not depend on in[:, :, h, w]: bit-identical outputs when only that pixel changes."""
    params, x, _ = C.network_inputs("net_blind_rgb")
    flat = _flat(params, O.param_order(3, 9, True))
    plan = engine.NetPlan(2, 3, 9, 32, 32, True, "cuda")
    a = plan.forward(flat, x.cuda(), training=False).clone()
    x2 = x.clone()
    x2[:, :, 13, 17] += 0.37
    b = plan.forward(flat, x2.cuda(), training=False)
    assert torch.equal(a[:, :, 13, 17], b[:, :, 13, 17])
    assert (a != b).float().mean() > 0.5                         # ... while most other outputs do change


def test_batch_sharding_invariance_at_baseline_size(engine):
    """BASELINE size (32 x 3 x 64 x 64, blind-spot): each half-batch run alone gives bit-identical outputs, i.e. samples
    are independent units - the property the data-parallel sharding relies on."""
    torch.manual_seed(0)
    p = O.init_params(3, 9, True)
    flat = _flat(p, O.param_order(3, 9, True))
    _, noisy = O.synthetic_batch(32, 3, 64, seed=1234)
    full = engine.NetPlan(32, 3, 9, 64, 64, True, "cuda").forward(flat, noisy.cuda(), training=False)
    half = engine.NetPlan(16, 3, 9, 64, 64, True, "cuda")
    lo = half.forward(flat, noisy[:16].cuda(), training=False).clone()
    hi = half.forward(flat, noisy[16:].cuda(), training=False)
    assert torch.equal(full[:16], lo) and torch.equal(full[16:], hi)
    assert torch.isfinite(full).all()


def test_shape_validation(engine):
    with pytest.raises(ValueError):
        engine.NetPlan(1, 3, 9, 40, 40, True, "cuda")
    with pytest.raises(ValueError):
        engine.NetPlan(1, 3, 9, 64, 32, True, "cuda")
    plan = engine.NetPlan(1, 3, 3, 64, 32, False, "cuda")        # plain network: rectangles are fine
    params = O.init_params(3, 3, False, generator=torch.Generator().manual_seed(1))
    x = torch.rand(1, 3, 64, 32)
    out = plan.forward(_flat(params, O.param_order(3, 3, False)), x.cuda(), training=False)
    assert rel(out, O.noise_network_forward(params, x, False)) < TOL
