"""Developer tool (GPU box): per-launch time / TFLOP/s of the tensor-core kernels in one training step of config 2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "selfsupervised-denoising_b200"))
import torch, bench
import ssdn
from ssdn import _engine as E
from ssdn.datasets import NoisyDataset
from ssdn.train import FlatAdam, train_step
den = ssdn.Denoiser(bench.make_cfg("known"), device="cuda"); opt = FlatAdam(den); opt.param_groups[0]["lr"] = 3e-4
clean, noisy, _ = bench.synthetic(32, 1234)
data = [noisy.cuda(), torch.zeros(0), {NoisyDataset.Metadata.INPUT_NOISE_VALUES: torch.full((32, 1, 1, 1), 25 / 255).cuda()}]
for _ in range(3): train_step(den, opt, data)
torch.cuda.synchronize()
E.profile_begin(); train_step(den, opt, data); E.profile_end()
names_f = ["enc1a", "enc1b", "enc2", "enc3", "enc4", "enc5", "enc6", "dec5a", "dec5b", "dec4a", "dec4b", "dec3a", "dec3b", "dec2a", "dec2b", "dec1a", "dec1b", "head1", "head2", "out"]
recs = E.profile_records()
fi = 0
tot = {}
for kind, ms, fl, by in recs:
    tag = names_f[fi] if kind == "conv_fwd" and fi < 20 else ""
    fi += kind == "conv_fwd"
    t = tot.setdefault(kind, [0, 0.0, 0.0, 0.0]); t[0] += 1; t[1] += ms; t[2] += fl; t[3] += by
    print(f"{kind:14s} {tag:6s} {ms*1e3:8.1f} us  {fl/1e9:8.2f} GFLOP  {fl/(ms*1e-3)/1e12 if ms > 0 else 0:7.1f} TFLOP/s  {by/1e6:8.1f} MB  {by/(ms*1e-3)/1e9 if ms > 0 else 0:7.0f} GB/s")
print("---- per kind")
for kind, (n, ms, fl, by) in tot.items():
    print(f"{kind:14s} x{n:3d} {ms*1e3:8.1f} us  {fl/(ms*1e-3)/1e12 if ms > 0 else 0:7.1f} TFLOP/s  {by/(ms*1e-3)/1e9 if ms > 0 else 0:7.0f} GB/s")
print(f"total {sum(t[1] for t in tot.values())*1e3:.1f} us in {sum(t[0] for t in tot.values())} launches")
