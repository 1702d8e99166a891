"""Developer tool (GPU box): device time and HBM rate of the on-GPU input pipeline kernel (ssdn_noisy_crops)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "selfsupervised-denoising_b200"))
import torch
from ssdn import _engine as E
imgs = torch.randint(0, 256, (512, 3, 256, 256), dtype=torch.uint8, device="cuda")      # 100 MB image cache
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
for n in (32, 1024, 8192):
    for _ in range(3): E.noisy_crops(imgs, n, 64, 1, 0, 25 / 255)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    a.record()
    for s in range(reps): E.noisy_crops(imgs, n, 64, 1, s, 25 / 255)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / reps
    byts = n * 64 * 64 * 3 * (1 + 4 + 4) + n * 3 * 4
    print(f"noisy_crops n={n:5d} 64x64x3: {us:8.1f} us per batch (incl. 3 torch.empty), {n / us * 1e6:12.0f} patches/s, "
          f"{byts / us / 1e3:8.1f} GB/s algorithmic = {byts / us / 1e3 / peak * 100:5.1f}% of measured HBM peak {peak:.0f} GB/s")
