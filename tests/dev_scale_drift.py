"""Dev tool (CPU, not a test): how fast do the per-tensor maxima move from one training step to the next?

    python tests/dev_scale_drift.py > profiles/r01_scale_drift.txt

Companion of tests/dev_split_numerics.py.  A two-term fp16 operand split needs a per-tensor power-of-two scale that puts
max|x| into [2^6, 2^15]; the cheapest way to have it BEFORE a tensor is written is to reuse the maximum observed for
the same tensor in the previous step ("delayed scaling").  That is safe if the maxima move by much less than the
window between consecutive steps - also during the first steps from a fresh initialisation at the full learning rate,
where they move fastest.  This script trains the blind-spot network from scratch on the CPU oracle (batch 4 x 3 x 64 x 64,
Adam lr 3e-4 without ramp-up, new noisy patches every step) and records, per layer and step, max|activation| and
max|dZ|, then prints the largest step-to-step ratio in binades."""
import math
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
sys.path.insert(0, HERE)
import ssdn_oracle as O  # noqa: E402
from oracle_trace import oracle_trace  # noqa: E402

torch.set_num_threads(8)
STEPS = int(os.environ.get("STEPS", "40"))


def main():
    torch.manual_seed(0)
    p = {k: v.clone().requires_grad_(True) for k, v in O.init_params(3, 9, True).items()}
    opt = torch.optim.Adam(list(p.values()), lr=3e-4, betas=(0.9, 0.99))
    hist = {}
    losses = []
    for step in range(STEPS):
        clean, noisy = O.synthetic_batch(4, 3, 64, seed=1000 + step)
        sigma = torch.full((4, 1, 1, 1), 25.0 / 255.0)
        opt.zero_grad()
        out, T = oracle_trace(p, noisy, True)
        out.retain_grad()
        loss = O.ssdn_posterior(out, noisy, sigma, True)["loss"].mean()
        loss.backward()
        losses.append(float(loss))
        for name, (z, a) in T.items():
            if z is None:
                continue
            hist.setdefault("act  " + name, []).append(float(a.detach().abs().max()))
            hist.setdefault("dZ   " + name, []).append(float(z.grad.abs().max()))
        hist.setdefault("dZ   output_conv", []).append(float(out.grad.abs().max()))
        for k, v in p.items():
            if k.endswith("weight"):
                hist.setdefault("W    " + k[:-7], []).append(float(v.detach().abs().max()))
        opt.step()
    print(f"{STEPS} steps from a fresh initialisation, Adam lr 3e-4 (no ramp-up), batch 4 x 3 x 64 x 64; loss {losses[0]:.4f} -> {losses[-1]:.4f}")
    print(f"{'tensor':<28} {'max at step 0':>14} {'max at end':>12} {'largest step-to-step move':>28} {'total range':>12}   (binades)")
    worst = 0.0
    for name in sorted(hist):
        h = hist[name]
        moves = [abs(math.log2(h[i + 1] / h[i])) for i in range(len(h) - 1) if h[i] > 0 and h[i + 1] > 0]
        rng = math.log2(max(h) / min(x for x in h if x > 0))
        worst = max(worst, max(moves))
        print(f"{name:<28} {h[0]:14.3e} {h[-1]:12.3e} {max(moves):28.2f} {rng:12.2f}")
    print(f"largest move of any tensor maximum between consecutive steps: {worst:.2f} binades "
          f"(window of the fp16 split with identical accuracy: 9 binades, [2^6, 2^15]; aimed at 2^11 -> 4 binades either side)")


if __name__ == "__main__":
    main()
