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


ACT_BUFFERS = [("e1a", "encode_block_1.0", 48), ("e1", "encode_block_1.2", 48), ("e2", "encode_block_2.0", 48), ("e3", "encode_block_3.0", 48),
               ("e4", "encode_block_4.0", 48), ("e5", "encode_block_5.0", 48), ("d_a5", "decode_block_5.0", 96), ("d_a4", "decode_block_4.0", 96),
               ("d_a3", "decode_block_3.0", 96), ("d_a2", "decode_block_2.0", 96), ("d_a1", "decode_block_1.0", 96), ("h2", "output_block.2", 96)]


@pytest.mark.parametrize("name", ["net_blind_rgb", "net_plain_rgb", "net_blind_rgb_64"])
def test_end_to_end_gradients_and_mask_flips(engine, name):
    """Engine forward + backward on its OWN activations versus autograd of the fp32 oracle.

    Given identical activations every gradient matches to 1e-4 (test above).  End to end, the only additional
    difference is the LeakyReLU derivative (1.0 vs 0.1) of pre-activations that are within the forward tolerance of
    zero: the engine's 3xTF32 forward agrees with the reference to ~4e-5 of the layer's range, so an activation that
    small may come out with the other sign.  This test pins exactly that: activations agree to TOL, sign flips happen
    ONLY on elements with |a| <= TOL * max|a| and are rare, and the gradients stay within a few 1e-3 in relative L2."""
    cin, cout, blind, n, size = C.NETWORK_CASES[name]
    params, x, dout = C.network_inputs(name)
    order = O.param_order(cin, cout, blind)
    flat = _flat(params, order)
    plan = engine.NetPlan(n, cin, cout, size, size, blind, "cuda")
    plan.forward(flat, x.cuda(), training=True)
    grads = _split(plan.backward(flat, dout.cuda()), params, order)
    plan.check()
    po = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    oo, T = oracle_trace(po, x, blind)
    oo.backward(dout)
    total = flips = 0
    for buf, layer, ch in ACT_BUFFERS:
        ref = T[layer][1].detach()
        got = plan.debug_read(buf, ch).cpu()
        scale = ref.abs().max()
        assert (got - ref).abs().max() <= TOL * scale, buf
        flipped = (got > 0) != (ref > 0)
        assert (ref[flipped].abs() <= TOL * scale).all(), buf       # flips only where the reference activation is ~0
        total += ref.numel()
        flips += int(flipped.sum())
    assert flips <= 1e-4 * total, (flips, total)
    for k in order:
        assert rel_l2(grads[k], po[k].grad) < 2e-2, k
    gold = C.load_golden(name)
    assert rel_l2(C.grad_summary(grads)[:, :2], gold["grad_summary"][:, :2]) < 2e-2


def test_blindspot_property_on_engine(engine):
    """out[:, :, h, w] must not depend on in[:, :, h, w]: bit-identical outputs when only that pixel changes.
    (Both passes run under the SAME operand scales - those left behind by a first pass over x - because the per-tensor
    power-of-two scales of the fp16 operand planes are the one global quantity every output rounding depends on.)"""
    params, x, _ = C.network_inputs("net_blind_rgb")
    flat = _flat(params, O.param_order(3, 9, True))
    plan = engine.NetPlan(2, 3, 9, 32, 32, True, "cuda")
    plan.forward(flat, x.cuda(), training=False)
    a = plan.forward(flat, x.cuda(), training=False).clone()
    x2 = x.clone()
    x2[:, :, 13, 17] += 0.37
    b = plan.forward(flat, x2.cuda(), training=False)
    assert torch.equal(a[:, :, 13, 17], b[:, :, 13, 17])
    assert (a != b).float().mean() > 0.5                         # ... while most other outputs do change


def test_batch_sharding_invariance_at_baseline_size(engine):
    """BASELINE size (32 x 3 x 64 x 64, blind-spot): each half-batch run alone gives the outputs of the full batch, i.e.
    samples are independent units - the property the data-parallel sharding relies on.  Bit-identical when the per-tensor
    operand scales agree (they come from the batch's maxima); to rounding (1e-6 of the output range) otherwise."""
    torch.manual_seed(0)
    p = O.init_params(3, 9, True)
    flat = _flat(p, O.param_order(3, 9, True))
    _, noisy = O.synthetic_batch(32, 3, 64, seed=1234)
    full = engine.NetPlan(32, 3, 9, 64, 64, True, "cuda").forward(flat, noisy.cuda(), training=False)
    half = engine.NetPlan(16, 3, 9, 64, 64, True, "cuda")
    lo = half.forward(flat, noisy[:16].cuda(), training=False).clone()
    hi = half.forward(flat, noisy[16:].cuda(), training=False)
    scale = full.abs().max()
    assert (full[:16] - lo).abs().max() <= 1e-6 * scale and (full[16:] - hi).abs().max() <= 1e-6 * scale
    assert torch.isfinite(full).all()
    # same plan, same data twice: the second and third passes run under identical scales and are bit-identical
    hi2 = half.forward(flat, noisy[16:].cuda(), training=False).clone()
    hi3 = half.forward(flat, noisy[16:].cuda(), training=False)
    assert torch.equal(hi2, hi3)


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


def test_forward_at_patch_128(engine):
    """BASELINE config 5 geometry (128 x 128 patches, 3-box halo windows): blind-spot forward vs the oracle."""
    p = O.init_params(3, 9, True, generator=torch.Generator().manual_seed(3))
    x = torch.rand(1, 3, 128, 128, generator=torch.Generator().manual_seed(4))
    plan = engine.NetPlan(1, 3, 9, 128, 128, True, "cuda")
    out = plan.forward(_flat(p, O.param_order(3, 9, True)), x.cuda(), training=False)
    plan.check()
    assert rel(out, O.noise_network_forward(p, x, True)) < TOL


@pytest.mark.parametrize("blind", [True, False])
def test_input_pack_im2col_operand(engine, blind):
    """pack_input_kernel: the first convolution's im2col operand xcol[pixel][c * 9 + kh * 3 + kw] = x_rot(c, i + kh - sh, j + kw - 1)
    (sh = 2: half-plane ShiftConv2d, models/noise_network.py:241-260; 1: plain conv), zero outside the image, for the rotation stack
    of utils/data.py:42-67 - and the input's own channel slot of the last concat buffer.  The values are the inputs themselves
    (two fp16 planes reproduce an fp32 to 2^-22), the padding is exactly zero."""
    n, size = 2, 32
    x = torch.rand(n, 3, size, size, generator=torch.Generator().manual_seed(5)) * 1.3 - 0.1
    params = O.init_params(3, 9 if blind else 3, blind, generator=torch.Generator().manual_seed(1))
    cout = 9 if blind else 3
    plan = engine.NetPlan(n, 3, cout, size, size, blind, "cuda")
    flat = _flat(params, O.param_order(3, cout, blind))
    plan.forward(flat, x.cuda(), training=False)
    plan.forward(flat, x.cuda(), training=False)                 # second pass: settled operand scales
    xr = O.rot4_stack(x) if blind else x                          # [4n or n, 3, H, W]
    sh = 2 if blind else 1
    padded = torch.nn.functional.pad(xr, (1, 1, sh, 2 - sh))
    cols = torch.stack([padded[:, :, kh:kh + size, kw:kw + size] for kh in range(3) for kw in range(3)], 2)   # [B, 3, 9, H, W]
    want = cols.reshape(xr.shape[0], 27, size, size)
    got = plan.debug_read("xcol", 32).cpu()
    assert got.shape == (xr.shape[0], 32, size, size)
    assert (got[:, :27] - want).abs().max() <= 1e-6 and torch.equal(got[:, 27:], torch.zeros_like(got[:, 27:]))
    assert torch.equal(got[:, :27][want == 0], torch.zeros_like(got[:, :27][want == 0]))                  # padding: exact zeros
    slot = plan.debug_read("cat1", 112).cpu()[:, 96:]
    assert (slot[:, :3] - xr).abs().max() <= 1e-6 and torch.equal(slot[:, 3:], torch.zeros_like(slot[:, 3:]))
