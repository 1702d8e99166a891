"""Test helper: oracle forward that keeps every intermediate, and upload of those intermediates into
the engine's internal buffers so that a backward pass can be checked kernel by kernel with identical
LeakyReLU masks (a pre-activation within rounding error of zero otherwise flips a 0.1/1.0 derivative
and hides real errors behind legitimate float noise)."""
import torch
import ssdn_oracle as O


def oracle_trace(p, x, blind):
    """Returns (output, T) with T[name] = (pre-activation z with retain_grad, activation)."""
    conv = O.shift_conv2d if blind else O.conv2d_same
    T = {}

    def cl(name, t, c=conv):
        z = c(t, p[name + ".weight"], p[name + ".bias"])
        z.retain_grad()
        a = O.lrelu(z)
        T[name] = (z, a)
        return a

    if blind:
        x = O.rot4_stack(x)
    T["x"] = (None, x)
    t = cl("encode_block_1.0", x)
    t = cl("encode_block_1.2", t)
    pools = [O.maxpool2(t, blind)]
    for i in (2, 3, 4, 5):
        pools.append(O.maxpool2(cl(f"encode_block_{i}.0", pools[-1]), blind))
    T["pools"] = (None, pools)
    t = O.upsample2(cl("encode_block_6.0", pools[4]))
    for i, skip in ((5, pools[3]), (4, pools[2]), (3, pools[1]), (2, pools[0])):
        t = torch.cat((t, skip), 1)
        T[f"cat{i}"] = (None, t)
        t = cl(f"decode_block_{i}.0", t)
        t = cl(f"decode_block_{i}.2", t)
        t = O.upsample2(t)
    t = torch.cat((t, x), 1)
    T["cat1"] = (None, t)
    t = cl("decode_block_1.0", t)
    t = cl("decode_block_1.2", t)
    if blind:
        t = O.shift_unrot_concat(t)
    T["head_in"] = (None, t)
    t = cl("output_block.0", t, O.conv2d_same)
    t = cl("output_block.2", t, O.conv2d_same)
    return O.conv2d_same(t, p["output_conv.weight"], p["output_conv.bias"]), T


def upload_activations(plan, T):
    w = lambda name, t: plan.debug_write(name, t.detach().cuda())
    for i in (1, 2, 3, 4, 5):
        w(f"cat{i}", T[f"cat{i}"][1])
        w(f"d_a{i}", T[f"decode_block_{i}.0"][1])
    w("e1a", T["encode_block_1.0"][1])
    w("e1", T["encode_block_1.2"][1])
    for i in (2, 3, 4, 5):
        w(f"e{i}", T[f"encode_block_{i}.0"][1])
    w("p5", T["pools"][1][4])
    w("head_in", T["head_in"][1])
    w("h1", T["output_block.0"][1])
    w("h2", T["output_block.2"][1])
