"""Developer tool (one GPU): how much a K-step Adam trajectory moves when the batch is processed as two shards whose
gradients are summed (what two data-parallel ranks do) instead of as one batch - per layer, and the per-layer relative
difference of the first gradient.  SSDN_LIB=<path> loads another build of the engine for A/B comparisons."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "selfsupervised-denoising_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from ssdn import _engine as E
if os.environ.get("SSDN_LIB"):
    E.LIB_PATH = os.environ["SSDN_LIB"]
import ssdn
from ssdn.params import PipelineOutput
from ssdn.train import FlatAdam
from dist_worker import make_cfg, global_batch, shard, rel_l2
K, n, size = 3, 8, 32
dev = torch.device("cuda", 0)
def run(split):
    torch.manual_seed(0)
    den = ssdn.Denoiser(make_cfg("ssdn", "known", 3), device=dev)
    opt = FlatAdam(den); opt.param_groups[0]["lr"] = 3e-4
    batch = global_batch("ssdn", n, size)
    first = None
    for _ in range(K):
        parts = [shard(batch, r, 2) for r in range(2)] if split else [batch]
        total = None
        for d in parts:
            d = [d[0].to(dev), d[1].to(dev) if d[1].numel() else d[1], {k: v.to(dev) for k, v in d[2].items()}]
            opt.zero_grad()
            out = den.run_pipeline(d)
            torch.mean(out[PipelineOutput.LOSS]).backward()
            g = den.flat_gradients().clone()
            total = g if total is None else total + g
        den.flat_gradients().copy_(total)
        if first is None: first = (total / len(parts)).clone()
        opt.step(grad_scale=1.0 / len(parts))
    torch.cuda.synchronize()
    return den, first
a, ga = run(False)
b, gb = run(True)
print(f"first gradient rel L2 (whole vector): {rel_l2(gb, ga):.2e}")
off = 0
for (name, pa), pb in zip(a.named_parameters(), b.parameters()):
    nel = pa.numel()
    ge = rel_l2(gb[off:off + nel], ga[off:off + nel]) if float(ga[off:off + nel].abs().max()) > 0 else 0.0
    gmed = float(ga[off:off + nel].abs().median())
    w = rel_l2(pb, pa) if pa.dim() > 1 and float(pa.abs().max()) > 0 else float((pa - pb).abs().max())
    print(f"{name:60s} grad rel {ge:.2e}  median|g| {gmed:.2e}  weights after {K} steps {w:.2e}")
    off += nel
