"""Developer check (GPU box): Denoiser pipelines vs golden fixtures (reference outputs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "selfsupervised-denoising_b200")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import cases as C
import ssdn_oracle as O
import ssdn
from ssdn.params import *
from ssdn.datasets import NoisyDataset

def rel(a, b): return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()
ALGO = {"ssdn": NoiseAlgorithm.SELFSUPERVISED_DENOISING, "n2c": NoiseAlgorithm.NOISE_TO_CLEAN, "n2v": NoiseAlgorithm.NOISE_TO_VOID}
MODE = {"known": NoiseValue.KNOWN, "const": NoiseValue.UNKNOWN_CONSTANT, "var": NoiseValue.UNKNOWN_VARIABLE, None: NoiseValue.KNOWN}
for name in C.PIPELINE_CASES:
    d = C.pipeline_inputs(name); gold = C.load_golden(name)
    cfg = ssdn.cfg.base(); cfg[ConfigValue.ALGORITHM] = ALGO[d["algorithm"]]; cfg[ConfigValue.NOISE_STYLE] = "gauss25"
    cfg[ConfigValue.NOISE_VALUE] = MODE[d["sigma_mode"]]; cfg[ConfigValue.IMAGE_CHANNELS] = d["channels"]
    ssdn.cfg.infer(cfg, model_only=True)
    den = ssdn.Denoiser(cfg, device="cuda")
    den.get_model(ssdn.Denoiser.MODEL, False).load_state_dict(d["params"], strict=False)
    if "est_params" in d: den.get_model(ssdn.Denoiser.SIGMA_ESTIMATOR, False).load_state_dict(d["est_params"], strict=False)
    if "est_sigma" in d: den.l_params[ssdn.Denoiser.ESTIMATED_SIGMA].data.copy_(d["est_sigma"])
    M = NoisyDataset.Metadata
    md = {M.CLEAN: d["clean"]}
    if "noise_values" in d: md[M.INPUT_NOISE_VALUES] = d["noise_values"]
    if "coords" in d: md[M.MASK_COORDS] = d["coords"]
    out = den.run_pipeline([d["noisy"], d.get("ref", torch.zeros(0)), md])
    out[PipelineOutput.LOSS].mean().backward()
    torch.cuda.synchronize()
    msg = f"{name}: loss rel {rel(out[PipelineOutput.LOSS].cpu(), gold['loss']):.2e} out rel {rel(out[PipelineOutput.IMG_DENOISED].detach().cpu(), gold['out']):.2e}"
    if "model_std" in gold:
        msg += f" model_std {rel(out[PipelineOutput.MODEL_STD_DEV].cpu(), gold['model_std']):.2e} noise_std {rel(out[PipelineOutput.NOISE_STD_DEV].detach().cpu().reshape(-1), gold['noise_std'].reshape(-1)):.2e} mu {rel(out[PipelineOutput.IMG_MU].detach().cpu(), gold['mu']):.2e}"
    main = den.get_model(ssdn.Denoiser.MODEL, False)
    grads = {k: p.grad.cpu() for k, p in main.named_parameters()}
    msg += f" | g_out_w {rel(grads['output_conv.weight'], gold['g_out_w']):.2e} g_first_w {rel(grads['encode_block_1.0.weight'], gold['g_first_w']):.2e} summary {rel(C.grad_summary(grads), gold['grad_summary']):.2e}"
    if "g_est_sigma" in gold: msg += f" g_est_sigma {rel(den.l_params[ssdn.Denoiser.ESTIMATED_SIGMA].grad.cpu(), gold['g_est_sigma']):.2e}"
    if "g_est_out_w" in gold:
        est = den.get_model(ssdn.Denoiser.SIGMA_ESTIMATOR, False)
        eg = {k: p.grad.cpu() for k, p in est.named_parameters()}
        msg += f" g_est_out_w {rel(eg['output_conv.weight'], gold['g_est_out_w']):.2e} est_summary {rel(C.grad_summary(eg), gold['est_grad_summary']):.2e}"
    print(msg, flush=True)
