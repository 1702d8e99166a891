"""Reference-side half of the ``.training`` wire-format tests (build container only; run as a subprocess because the
reference package and this repo's drop-in share the name ``ssdn``).

    python tests/ref_training_io.py write <file.training> <known|const|var|n2c>   # the UNMODIFIED reference trains 2 steps, saves
    python tests/ref_training_io.py read  <file.training>                         # the UNMODIFIED reference resumes, trains 1 step

Both print one JSON line describing what the reference holds, for the test to compare with this package's view of
the same file.  Follows the reference trainer's own sequence (train.py:100-107,114-125,197-221,711-745)."""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from _ref_shim import import_reference  # noqa: E402

ssdn = import_reference()
from ssdn.datasets import FixedLengthSampler, NoisyDataset  # noqa: E402
from ssdn.params import ConfigValue, HistoryValue, NoiseAlgorithm, NoiseValue, PipelineOutput, StateValue  # noqa: E402
from ssdn.train import DenoiserTrainer  # noqa: E402

torch.set_num_threads(4)
BATCH, PATCH = 2, 32


def make_cfg(kind: str):
    cfg = ssdn.cfg.base()
    cfg[ConfigValue.IMAGE_CHANNELS] = 1
    cfg[ConfigValue.NOISE_STYLE] = "gauss25"
    cfg[ConfigValue.TRAIN_ITERATIONS] = 40
    cfg[ConfigValue.TRAIN_MINIBATCH_SIZE] = BATCH
    cfg[ConfigValue.TRAIN_PATCH_SIZE] = PATCH
    if kind == "n2c":
        cfg[ConfigValue.ALGORITHM] = NoiseAlgorithm.NOISE_TO_CLEAN
    else:
        cfg[ConfigValue.ALGORITHM] = NoiseAlgorithm.SELFSUPERVISED_DENOISING
        cfg[ConfigValue.NOISE_VALUE] = {"known": NoiseValue.KNOWN, "const": NoiseValue.UNKNOWN_CONSTANT, "var": NoiseValue.UNKNOWN_VARIABLE}[kind]
    ssdn.cfg.infer(cfg, model_only=True)
    return cfg


def batch(step: int, channels: int):
    g = torch.Generator().manual_seed(100 + step)
    clean = torch.rand(BATCH, channels, PATCH, PATCH, generator=g)
    noisy = (clean + torch.randn(clean.shape, generator=g) * (25 / 255)).clamp(0, 1)
    md = {NoisyDataset.Metadata.INPUT_NOISE_VALUES: torch.full((BATCH, 1, 1, 1), 25 / 255), NoisyDataset.Metadata.CLEAN: clean,
          NoisyDataset.Metadata.IMAGE_SHAPE: torch.tensor([[channels, PATCH, PATCH]] * BATCH)}
    return [noisy, clean, md]


def step(trainer: DenoiserTrainer, data):
    """train.py:195-221, minus PSNR bookkeeping."""
    trainer.denoiser.train()
    opt = trainer.optimizer
    opt.zero_grad()
    out = trainer.denoiser.run_pipeline(data)
    torch.mean(out[PipelineOutput.LOSS]).backward()
    opt.step()
    hist = trainer.state[StateValue.HISTORY][HistoryValue.TRAIN]
    n = data[NoisyDataset.INPUT].shape[0]
    with torch.no_grad():
        hist["n"] += n
        hist["loss"] += out[PipelineOutput.LOSS]
    trainer.state[StateValue.ITERATION] += n
    return float(out[PipelineOutput.LOSS].mean())


def describe(trainer: DenoiserTrainer, order) -> dict:
    opt = trainer._optimizer
    params = list(trainer.denoiser.parameters())
    st = opt.state_dict()["state"]
    hist = trainer.state[StateValue.HISTORY]
    return {
        "iteration": trainer.state[StateValue.ITERATION],
        "lr": trainer.learning_rate,
        "n_params": len(params),
        "param_numel": [p.numel() for p in params],
        "param_sum": [float(p.detach().double().sum()) for p in params],
        "steps": [int(float(st[i]["step"])) for i in sorted(st)],
        "exp_avg_sum": [float(st[i]["exp_avg"].double().sum()) for i in sorted(st)],
        "exp_avg_sq_sum": [float(st[i]["exp_avg_sq"].double().sum()) for i in sorted(st)],
        "betas": list(opt.param_groups[0]["betas"]), "eps": opt.param_groups[0]["eps"],
        "order": list(order.order), "order_index": order.index,
        "train_n": hist[HistoryValue.TRAIN]["n"],
        "train_loss_mean": float(hist[HistoryValue.TRAIN]["loss"].accumulated()) if not hist[HistoryValue.TRAIN]["loss"].empty() else None,
        "timing_keys": sorted(hist[HistoryValue.TIMINGS].keys()),
    }


def main():
    mode, path = sys.argv[1], sys.argv[2]
    if mode == "write":
        torch.manual_seed(7)
        trainer = DenoiserTrainer(make_cfg(sys.argv[3]), state={}, runs_dir=os.path.dirname(path), run_dir="run")
        trainer.new_target()
        trainer.train_sampler = FixedLengthSampler(list(range(10)), num_samples=40, shuffled=True)
        order = iter(trainer.train_sampler)
        channels = trainer.cfg[ConfigValue.IMAGE_CHANNELS]
        trainer.state[StateValue.HISTORY][HistoryValue.TIMINGS]["total"].update()
        losses = [step(trainer, batch(s, channels)) for s in range(2)]
        for _ in range(6):                      # a loader prefetches ahead of what was processed: the saved cursor must not follow it
            next(order)
        torch.save(trainer.state_dict(), path)
        info = describe(trainer, trainer.train_sampler.last_iter())
        info["order_index"] = trainer.state[StateValue.ITERATION]       # what state_dict() writes (train.py:722-723)
        info["losses"] = losses
    else:
        trainer = DenoiserTrainer(None, runs_dir=os.path.dirname(path), run_dir="run")
        trainer.load_state_dict(torch.load(path, map_location="cpu", weights_only=False))
        info = describe(trainer, trainer._train_iter)
        # what the reference touches right after a resume (train.py:166-181): keys it has not created, the integer counters
        hist = trainer.state[StateValue.HISTORY]
        hist[HistoryValue.TIMINGS]["last_print"].update()
        samples = hist[HistoryValue.EVAL]["n"] + hist[HistoryValue.TRAIN]["n"]
        info["samples"] = samples
        trainer.reset_metrics()
        channels = trainer.cfg[ConfigValue.IMAGE_CHANNELS]
        info["resumed_loss"] = step(trainer, batch(2, channels))
        info["steps_after"] = [int(float(s["step"])) for s in trainer._optimizer.state_dict()["state"].values()]
        info["param_sum_after"] = [float(p.detach().double().sum()) for p in trainer.denoiser.parameters()]
    print("JSON " + json.dumps(info))


if __name__ == "__main__":
    main()
