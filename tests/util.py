"""Shared helpers of the parity tests."""
import torch

import ssdn
from ssdn.datasets import NoisyDataset
from ssdn.params import ConfigValue, NoiseAlgorithm, NoiseValue

TOL = 1e-4          # BASELINE.json north_star: outputs within 1e-4 relative fp32 of the reference path


def rel(a, b):
    """max |a - b| / max |b|"""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def as_accurate_as_reference(engine_val, ref32, ref64, slack=3.0, floor=TOL, norm=rel):
    """Acceptance rule for quantities where the fp32 reference itself is unstable (sign of a pre-activation within
    rounding error of zero flips a LeakyReLU derivative; inverse of a near-singular covariance): the engine must be as
    close to the exact (fp64) result as the fp32 reference is, within `slack`, or within the plain tolerance."""
    e_engine, e_ref = norm(engine_val, ref64), norm(ref32, ref64)
    return e_engine <= max(floor, slack * e_ref), (e_engine, e_ref)


ALGO = {"ssdn": NoiseAlgorithm.SELFSUPERVISED_DENOISING, "n2c": NoiseAlgorithm.NOISE_TO_CLEAN, "n2n": NoiseAlgorithm.NOISE_TO_NOISE,
        "n2v": NoiseAlgorithm.NOISE_TO_VOID}
MODE = {"known": NoiseValue.KNOWN, "const": NoiseValue.UNKNOWN_CONSTANT, "var": NoiseValue.UNKNOWN_VARIABLE, None: NoiseValue.KNOWN}


def make_cfg(algorithm="ssdn", sigma_mode="known", channels=3, style="gauss25", diagonal=False):
    cfg = ssdn.cfg.base()
    cfg[ConfigValue.DIAGONAL_COVARIANCE] = diagonal
    cfg[ConfigValue.ALGORITHM] = ALGO[algorithm]
    cfg[ConfigValue.NOISE_STYLE] = style
    cfg[ConfigValue.NOISE_VALUE] = MODE[sigma_mode]
    cfg[ConfigValue.IMAGE_CHANNELS] = channels
    ssdn.cfg.infer(cfg, model_only=True)
    return cfg


def denoiser_for_case(d, device="cuda"):
    """Engine Denoiser loaded with the weights of a tests/golden pipeline case."""
    den = ssdn.Denoiser(make_cfg(d["algorithm"], d["sigma_mode"], d["channels"], d.get("noise_style", "gauss25"), d.get("diagonal", False)), device=device)
    den.get_model(ssdn.Denoiser.MODEL, False).load_state_dict(d["params"], strict=False)
    if "est_params" in d:
        den.get_model(ssdn.Denoiser.SIGMA_ESTIMATOR, False).load_state_dict(d["est_params"], strict=False)
    if "est_sigma" in d:
        den.l_params[ssdn.Denoiser.ESTIMATED_SIGMA].data.copy_(d["est_sigma"])
    return den


def data_for_case(d):
    M = NoisyDataset.Metadata
    md = {M.CLEAN: d["clean"]}
    if "noise_values" in d:
        md[M.INPUT_NOISE_VALUES] = d["noise_values"]
    if "coords" in d:
        md[M.MASK_COORDS] = d["coords"]
    return [d["noisy"], d.get("ref", torch.zeros(0)), md]


def oracle_case(d, dtype=torch.float32):
    """Oracle forward + backward of mean(loss) for a pipeline case; returns (outputs, grads main, grads est, grad sigma)."""
    import ssdn_oracle as O
    cast = lambda t: t.to(dtype) if t.is_floating_point() else t  # noqa: E731
    p = {k: cast(v).clone().requires_grad_(True) for k, v in d["params"].items()}
    ep = {k: cast(v).clone().requires_grad_(True) for k, v in d["est_params"].items()} if "est_params" in d else None
    es = cast(d["est_sigma"]).clone().requires_grad_(True) if "est_sigma" in d else None
    if d["algorithm"] == "ssdn":
        out = O.ssdn_pipeline(p, cast(d["noisy"]), cast(d["noise_values"]), d["sigma_mode"], ep, es, noise_style=d.get("noise_style", "gauss"),
                              diagonal=d.get("diagonal", False))
    elif d["algorithm"] == "n2v":
        out = O.mask_mse_pipeline(p, cast(d["noisy"]), cast(d["ref"]), d["coords"])
    else:
        out = O.mse_pipeline(p, cast(d["noisy"]), cast(d["ref"]))
    out["loss"].mean().backward()
    return out, {k: v.grad for k, v in p.items()}, ({k: v.grad for k, v in ep.items()} if ep else None), (es.grad if es is not None else None)
