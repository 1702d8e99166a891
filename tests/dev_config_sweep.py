"""Developer tool (GPU box): one training step of every BASELINE.json configuration shape; loss of the first two samples
against the CPU oracle, device-side error flags, finiteness, step time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("selfsupervised-denoising_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
import ssdn
import ssdn_oracle as O
from ssdn.datasets import NoisyDataset
from ssdn.params import PipelineOutput
from ssdn.train import FlatAdam, train_step
from util import make_cfg
M = NoisyDataset.Metadata
CASES = [("cfg1 n2c mono 32 bs4", "n2c", None, 1, 4, 32, None), ("cfg2 ssdn known 64 bs32", "ssdn", "known", 3, 32, 64, "one"),
         ("cfg3 ssdn var 64 bs32", "ssdn", "var", 3, 32, 64, "one"), ("cfg4 n2v 64 bs32", "n2v", None, 3, 32, 64, None),
         ("cfg5 ssdn var gauss5_50 128 bs16", "ssdn", "var", 3, 16, 128, "perchannel"), ("odd: ssdn known 96 bs5", "ssdn", "known", 3, 5, 96, "one")]
for name, algo, mode, c, n, size, sig in CASES:
    torch.manual_seed(0)
    den = ssdn.Denoiser(make_cfg(algo, mode or "known", c, "gauss5_50" if sig == "perchannel" else "gauss25"), device="cuda")
    opt = FlatAdam(den); opt.param_groups[0]["lr"] = 3e-4
    clean, noisy = O.synthetic_batch(n, c, size, seed=5)
    md = {M.CLEAN: clean}
    g = torch.Generator().manual_seed(1)
    if sig == "one": md[M.INPUT_NOISE_VALUES] = torch.full((n, 1, 1, 1), 25 / 255)
    if sig == "perchannel": md[M.INPUT_NOISE_VALUES] = (torch.rand(n, c, 1, 1, generator=g) * 45 + 5) / 255
    ref = torch.zeros(0)
    if algo == "n2c": ref = clean
    if algo == "n2v":
        ref = (clean + torch.randn(clean.shape, generator=g) * 25 / 255).clamp(0, 1)
        md[M.MASK_COORDS] = torch.randint(0, size, (n, 64, 2), generator=g)
    data = [noisy, ref, md]
    main = den.get_model(ssdn.Denoiser.MODEL, False)
    params = {k: v.detach().cpu().clone() for k, v in main.state_dict().items() if not k.startswith("output_block.4")}
    est = {k: v.detach().cpu().clone() for k, v in den.get_model(ssdn.Denoiser.SIGMA_ESTIMATOR, False).state_dict().items()
           if not k.startswith("output_block.4")} if mode == "var" else None
    out = train_step(den, opt, data)
    torch.cuda.synchronize()
    for net in den._models.values():
        for plan in net._plans.values(): plan.check()
    loss = out[PipelineOutput.LOSS].detach().cpu().view(-1)
    k = min(2, n)

    def oracle(dt):
        cast = lambda t: t.to(dt) if torch.is_tensor(t) and t.is_floating_point() else t     # noqa: E731
        pp = {a: cast(b) for a, b in params.items()}
        with torch.no_grad():
            if algo == "ssdn":
                return O.ssdn_pipeline(pp, cast(noisy[:k]), cast(md[M.INPUT_NOISE_VALUES][:k]), mode,
                                       {a: cast(b) for a, b in est.items()} if est else None)["loss"].view(-1)
            if algo == "n2v":
                return O.mask_mse_pipeline(pp, cast(noisy[:k]), cast(ref[:k]), md[M.MASK_COORDS][:k])["loss"].view(-1)
            return O.mse_pipeline(pp, cast(noisy[:k]), cast(ref[:k]))["loss"].view(-1)
    o32, o64 = oracle(torch.float32), oracle(torch.float64)
    err = ((loss[:k].double() - o64).abs().max() / o64.abs().max()).item()            # engine vs exact arithmetic
    err32 = ((o32.double() - o64).abs().max() / o64.abs().max()).item()               # fp32 reference vs exact arithmetic
    for _ in range(3): train_step(den, opt, data)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): train_step(den, opt, data)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 100
    print(f"{name:36s} loss rel err vs fp64 oracle: engine {err:.2e}, fp32 reference path {err32:.2e}  finite {bool(torch.isfinite(loss).all())}  {ms:7.2f} ms/step  {n / ms * 1e3:8.0f} patches/s", flush=True)
    assert err < max(1e-4, 2 * err32) and torch.isfinite(loss).all()
print("all configurations ok")
