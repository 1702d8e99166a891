"""Developer tool (one GPU): the graphed training step at the per-GPU batch sizes of STRONG scaling (32 / world patches), i.e.
what bench.py's `strong` field measures on each rank, without needing the other GPUs (the all-reduce aside).
usage: python tests/dev_small_batch.py [batches...]   (default 4 8 16)"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "selfsupervised-denoising_b200"))
import torch, bench
dev = torch.device("cuda", 0)
for n in [int(a) for a in sys.argv[1:]] or [4, 8, 16]:
    case = bench.Case("known", n, dev, 0, 1, True)
    for _ in range(4): case.step_resident()
    case.capture()
    ts = [bench.timed_region(case.step_resident, 30, dev, False) / 30 * 1e3 for _ in range(3)]
    print(f"batch {n:3d}: {statistics.median(ts):.3f} ms per step ({n / statistics.median(ts) * 1e3:.0f} patches/s per GPU; runs {[round(t, 3) for t in ts]})", flush=True)
    del case
