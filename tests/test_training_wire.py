"""CPU: ``.training`` checkpoint wire format in both directions against the UNMODIFIED reference trainer
(reference train.py:711-745, 871-909).  The reference runs in a subprocess (tests/ref_training_io.py):
it shares the package name ``ssdn`` with this repo's drop-in.  Build container only - the GPU box has no /root/reference."""
import json
import os
import subprocess
import sys
from collections import defaultdict

import pytest
import torch

import ssdn
from ssdn.datasets import FixedLengthSampler
from ssdn.params import ConfigValue, HistoryValue, NoiseAlgorithm, NoiseValue, StateValue
from ssdn.train import DenoiserTrainer, FlatAdam

HERE = os.path.dirname(os.path.abspath(__file__))
needs_reference = pytest.mark.skipif(not os.path.isdir("/root/reference/ssdn"), reason="the reference only exists in the build container")


def _reference(*args) -> dict:
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}       # the subprocess must import the REFERENCE's ssdn
    out = subprocess.run([sys.executable, os.path.join(HERE, "ref_training_io.py"), *map(str, args)], capture_output=True, text=True,
                         timeout=900, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("JSON ")][-1]
    return json.loads(line[5:])


def _sums(tensors):
    return [float(t.detach().double().sum()) for t in tensors]


def _split(flat, params):
    out, off = [], 0
    for p in params:
        out.append(flat[off:off + p.numel()])
        off += p.numel()
    assert off == flat.numel()
    return out


@needs_reference
@pytest.mark.parametrize("kind", ["known", "const", "var", "n2c"])
def test_reference_training_file_resumes_here(kind, tmp_path):
    """A snapshot written by the reference trainer after two of its own Adam steps restores here: parameters, the Adam
    moments in flat layout, the step counter, the image counter, the LR it implies, the sampling order and the metric
    history."""
    path = tmp_path / "model_00000004.training"
    ref = _reference("write", path, kind)
    trainer = DenoiserTrainer(None, runs_dir=str(tmp_path), run_dir="run")
    trainer.load_state_dict(str(path), device="cpu")
    params = list(trainer.denoiser.parameters())
    assert [p.numel() for p in params] == ref["param_numel"]
    assert _sums(params) == pytest.approx(ref["param_sum"], rel=1e-12, abs=1e-12)
    opt = trainer._optimizer
    assert opt.step_count == 2 and set(ref["steps"]) == {2}
    assert opt.exp_avg.numel() == sum(ref["param_numel"])
    assert _sums(_split(opt.exp_avg, params)) == pytest.approx(ref["exp_avg_sum"], rel=1e-12, abs=1e-15)
    assert _sums(_split(opt.exp_avg_sq, params)) == pytest.approx(ref["exp_avg_sq_sum"], rel=1e-12, abs=1e-18)
    assert list(opt.param_groups[0]["betas"]) == ref["betas"] and opt.param_groups[0]["eps"] == ref["eps"]
    assert trainer.state[StateValue.ITERATION] == ref["iteration"] == 4
    assert trainer.learning_rate == pytest.approx(ref["lr"], rel=1e-12)
    hist = trainer.state[StateValue.HISTORY]
    assert hist[HistoryValue.TRAIN]["n"] == ref["train_n"] == 4
    assert float(hist[HistoryValue.TRAIN]["loss"].accumulated()) == pytest.approx(ref["train_loss_mean"], rel=1e-6)
    assert isinstance(hist[HistoryValue.TIMINGS], defaultdict) and "total" in hist[HistoryValue.TIMINGS]
    # the restored order continues where the PROCESSED images ended, not where the loader's prefetch had got to
    sampler = FixedLengthSampler(list(range(10)), num_samples=40, shuffled=True)
    trainer.attach_sampler(sampler)
    assert list(iter(sampler)) == ref["order"][ref["order_index"]:] and ref["order_index"] == 4
    # and what we write back is, tensor for tensor, what the reference wrote
    again = trainer.state_dict()
    orig = torch.load(path, map_location="cpu", weights_only=False)
    assert set(again) == set(orig)
    assert set(again["denoiser"]) == set(orig["denoiser"])
    assert all(torch.equal(again["denoiser"][k], v) for k, v in orig["denoiser"].items() if torch.is_tensor(v))
    assert again["train_order_iter"]["order"] == orig["train_order_iter"]["order"]
    assert again["train_order_iter"]["index"] == orig["train_order_iter"]["index"]
    assert set(again["optimizer"]["state"]) == set(orig["optimizer"]["state"])
    for i, s in orig["optimizer"]["state"].items():
        assert float(again["optimizer"]["state"][i]["step"]) == float(s["step"])
        assert torch.equal(again["optimizer"]["state"][i]["exp_avg"], s["exp_avg"])
        assert torch.equal(again["optimizer"]["state"][i]["exp_avg_sq"], s["exp_avg_sq"])
    assert again["optimizer"]["param_groups"][0]["params"] == orig["optimizer"]["param_groups"][0]["params"]


def _cfg(kind: str):
    cfg = ssdn.cfg.base()
    cfg[ConfigValue.IMAGE_CHANNELS] = 1
    cfg[ConfigValue.NOISE_STYLE] = "gauss25"
    cfg[ConfigValue.TRAIN_ITERATIONS] = 40
    cfg[ConfigValue.TRAIN_MINIBATCH_SIZE] = 2
    cfg[ConfigValue.TRAIN_PATCH_SIZE] = 32
    if kind == "n2c":
        cfg[ConfigValue.ALGORITHM] = NoiseAlgorithm.NOISE_TO_CLEAN
    else:
        cfg[ConfigValue.ALGORITHM] = NoiseAlgorithm.SELFSUPERVISED_DENOISING
        cfg[ConfigValue.NOISE_VALUE] = {"known": NoiseValue.KNOWN, "const": NoiseValue.UNKNOWN_CONSTANT, "var": NoiseValue.UNKNOWN_VARIABLE}[kind]
    return cfg


def _fabricated_trainer(kind, tmp_path):
    """A trainer in the state it has after 3 optimiser steps of batch 2.  The moments are fabricated (the Adam kernel needs
    the GPU; tests/test_gpu_pipeline.py covers a real interrupted run) - what is under test here is the file."""
    torch.manual_seed(11)
    trainer = DenoiserTrainer(_cfg(kind), runs_dir=str(tmp_path), run_dir="run")
    trainer.new_target(device="cpu")
    sampler = FixedLengthSampler(list(range(10)), num_samples=40, shuffled=True)
    trainer.attach_sampler(sampler)
    order = iter(sampler)
    for _ in range(9):
        next(order)
    opt = trainer._optimizer
    total = sum(p.numel() for p in trainer.denoiser.parameters())
    g = torch.Generator().manual_seed(5)
    opt.step_count = 3
    opt.exp_avg = torch.randn(total, generator=g) * 1e-3
    opt.exp_avg_sq = torch.rand(total, generator=g) * 1e-5
    trainer.state[StateValue.ITERATION] = 6
    hist = trainer.state[StateValue.HISTORY]
    hist[HistoryValue.TRAIN]["n"] += 6
    hist[HistoryValue.TRAIN]["loss"] += torch.full((6, 1), 0.25)
    hist[HistoryValue.TIMINGS]["total"].update()
    return trainer, sampler


@needs_reference
@pytest.mark.parametrize("kind", ["known", "const", "var"])
def test_training_file_written_here_resumes_in_reference(kind, tmp_path):
    """The other direction: the unmodified reference loads a snapshot written by this package (torch.optim.Adam accepts the
    optimiser state, the trainer finds every key and container type it indexes after a resume) and trains on from it."""
    trainer, sampler = _fabricated_trainer(kind, tmp_path)
    path = trainer.snapshot()
    assert path.endswith(os.path.join("run", "training", "model_00000006.training"))
    params = list(trainer.denoiser.parameters())
    ref = _reference("read", path)
    assert ref["param_numel"] == [p.numel() for p in params]
    assert ref["param_sum"] == pytest.approx(_sums(params), rel=1e-12, abs=1e-12)
    assert set(ref["steps"]) == {3} and len(ref["steps"]) == len(params)
    assert ref["exp_avg_sum"] == pytest.approx(_sums(_split(trainer._optimizer.exp_avg, params)), rel=1e-12, abs=1e-15)
    assert ref["exp_avg_sq_sum"] == pytest.approx(_sums(_split(trainer._optimizer.exp_avg_sq, params)), rel=1e-12, abs=1e-18)
    assert ref["betas"] == [0.9, 0.99] and ref["eps"] == 1e-8
    assert ref["iteration"] == 6 and ref["train_n"] == 6 and ref["samples"] == 6
    assert ref["lr"] == pytest.approx(trainer.learning_rate, rel=1e-12)
    assert ref["train_loss_mean"] == pytest.approx(0.25)
    assert ref["order"] == sampler.last_iter().order and ref["order_index"] == 6
    assert "total" in ref["timing_keys"]
    # the reference then took one Adam step of its own from that state
    assert set(ref["steps_after"]) == {4}
    assert all(abs(v) < 1e6 for v in ref["param_sum_after"]) and ref["resumed_loss"] == ref["resumed_loss"]
    assert ref["param_sum_after"] != ref["param_sum"]


def test_training_file_round_trip_without_reference(tmp_path):
    """Same file through this package alone (runs everywhere): snapshot -> resume_run restores optimiser, counters, order."""
    trainer, sampler = _fabricated_trainer("const", tmp_path)
    trainer.snapshot()
    resumed = ssdn.train.resume_run(os.path.join(str(tmp_path), "run"), device="cpu")
    assert resumed.state[StateValue.ITERATION] == 6 and resumed._optimizer.step_count == 3
    assert torch.equal(resumed._optimizer.exp_avg, trainer._optimizer.exp_avg)
    assert torch.equal(resumed._optimizer.exp_avg_sq, trainer._optimizer.exp_avg_sq)
    assert all(torch.equal(a, b) for a, b in zip(resumed.denoiser.parameters(), trainer.denoiser.parameters()))
    assert resumed.cfg == trainer.cfg and resumed.learning_rate == trainer.learning_rate
    assert resumed.state[StateValue.HISTORY][HistoryValue.TIMINGS]["total"].last_time is None       # stale clocks dropped
    again = FixedLengthSampler(list(range(10)), num_samples=40, shuffled=True)
    resumed.attach_sampler(again)
    assert list(iter(again)) == sampler.last_iter().order[6:]
    with pytest.raises(ValueError):
        ssdn.train.resume_run(os.path.join(str(tmp_path), "run"), iteration=8, device="cpu")


def test_flat_adam_accepts_old_torch_and_flat_layouts(tmp_path):
    """Integer ``step`` (torch < 1.12, the version the reference was written for), list betas, and the flat layout
    that early snapshots of this package used."""
    trainer, _ = _fabricated_trainer("known", tmp_path)
    opt = trainer._optimizer
    sd = opt.state_dict()
    for s in sd["state"].values():
        s["step"] = 3
    sd["param_groups"] = [{"lr": 1e-4, "betas": [0.9, 0.99], "eps": 1e-8, "weight_decay": 0, "amsgrad": False,
                           "params": sd["param_groups"][0]["params"]}]
    other = FlatAdam(trainer.denoiser)
    other.load_state_dict(sd)
    assert other.step_count == 3 and torch.equal(other.exp_avg, opt.exp_avg) and other.param_groups[0]["betas"] == (0.9, 0.99)
    other.load_state_dict({"step": 5, "exp_avg": opt.exp_avg, "exp_avg_sq": opt.exp_avg_sq, "param_groups": [{"lr": 1e-3, "betas": (0.9, 0.99), "eps": 1e-8}]})
    assert other.step_count == 5
    fresh = FlatAdam(trainer.denoiser)
    assert fresh.state_dict()["state"] == {}                     # no step taken yet: torch.optim.Adam's empty state
    other.load_state_dict(fresh.state_dict())
    assert other.step_count == 0 and other.exp_avg is None
    bad = opt.state_dict()
    bad["param_groups"][0]["weight_decay"] = 0.1
    with pytest.raises(NotImplementedError):
        other.load_state_dict(bad)
    bad = opt.state_dict()
    bad["param_groups"][0]["params"] = bad["param_groups"][0]["params"][:-1]
    with pytest.raises(ValueError):
        other.load_state_dict(bad)


def test_trainer_housekeeping_intervals_on_its_own_data(tmp_path, monkeypatch, caplog):
    """The reference-style run: the trainer reads a folder data set through its own loaders and, at multiples of the
    configured intervals, evaluates, reports + resets the metric window and snapshots; it ends with a snapshot and
    ``final-<config>.wt`` (train.py:155-181, 223-231).  The engine calls are replaced by stand-ins (no GPU here): what is
    under test is the host-side sequencing around them."""
    import logging
    import numpy as np
    from PIL import Image
    from ssdn.params import PipelineOutput
    data_dir = tmp_path / "kodak_tiny"
    data_dir.mkdir()
    rng = np.random.default_rng(1)
    for i in range(3):
        Image.fromarray(rng.integers(0, 256, (40, 48, 3), dtype=np.uint8), mode="RGB").save(data_dir / "im{}.png".format(i))
    cfg = _cfg("known")
    cfg[ConfigValue.IMAGE_CHANNELS] = 3
    cfg[ConfigValue.TRAIN_DATA_PATH] = cfg[ConfigValue.TEST_DATA_PATH] = str(data_dir)
    cfg[ConfigValue.DATALOADER_WORKERS] = 0
    cfg[ConfigValue.TRAIN_ITERATIONS], cfg[ConfigValue.TRAIN_MINIBATCH_SIZE], cfg[ConfigValue.TRAIN_PATCH_SIZE] = 12, 2, 32
    cfg[ConfigValue.EVAL_INTERVAL], cfg[ConfigValue.PRINT_INTERVAL], cfg[ConfigValue.SNAPSHOT_INTERVAL] = 8, 4, 6
    trainer = DenoiserTrainer(cfg, runs_dir=str(tmp_path), run_dir="run")
    trainer.new_target(device="cpu")
    steps, evals = [], []

    def fake_step(denoiser, optimizer, data, world_size):
        steps.append((trainer.state[StateValue.ITERATION], optimizer.param_groups[0]["lr"], tuple(data[0].shape)))
        return {PipelineOutput.LOSS: torch.full((data[0].shape[0], 1), 0.5), PipelineOutput.IMG_DENOISED: data[0]}

    def fake_pipeline(data, **kwargs):
        evals.append(tuple(data[0].shape))
        return {PipelineOutput.IMG_DENOISED: data[0], PipelineOutput.INPUTS: data}

    monkeypatch.setattr(ssdn.train, "train_step", fake_step)
    monkeypatch.setattr(trainer.denoiser, "run_pipeline", fake_pipeline)
    monkeypatch.setattr(ssdn.cfg, "test_length", lambda name: 4)                 # 240 passes over 'kodak' would only take time
    with caplog.at_level(logging.INFO, logger="ssdn.train"):
        trainer.train()
    assert [s[0] for s in steps] == [0, 2, 4, 6, 8, 10] and all(s[2] == (2, 3, 32, 32) for s in steps)
    assert steps[0][1] == 0.0 and steps[-1][1] < 3e-4                            # the ramps follow the image counter
    assert len(evals) == 2 * 2 and all(e == (2, 3, 64, 64) for e in evals)      # at images 0 and 8: 4 padded test images, 2 per batch
    assert set(trainer.last_eval) == {"psnr_out"} and 5.0 < trainer.last_eval["psnr_out"] < 40.0
    lines = [r.getMessage() for r in caplog.records if r.name == "ssdn.train"]
    assert [ln[:10] for ln in lines] == ["[00000000]", "[00000004]", "[00000008]", "[00000012]"]
    assert "loss=0.5000" in lines[1] and "n=4" in lines[1] and "VALID psnr_out=" in lines[1]
    run = tmp_path / "run"
    assert sorted(p.name for p in (run / "training").iterdir()) == ["model_00000000.training", "model_00000006.training", "model_00000012.training"]
    assert [p.name for p in run.glob("final-*.wt")] == ["final-ssdn-gauss25-sigma_known.wt"]
    final = torch.load(run / "final-ssdn-gauss25-sigma_known.wt", map_location="cpu", weights_only=False)
    assert "cfg" in final and "state" not in final
    resumed = ssdn.train.resume_run(str(run), device="cpu")
    assert resumed.state[StateValue.ITERATION] == 12 and len(resumed._train_iter.order) == 12 and resumed._train_iter.index == 12
    # with caller-supplied batches nothing of this happens unless asked for
    quiet = DenoiserTrainer(_cfg("known"), runs_dir=str(tmp_path), run_dir="quiet")
    quiet.new_target(device="cpu")
    batch = [torch.rand(2, 1, 32, 32), torch.zeros(0), {}]
    quiet.train([batch, batch])
    assert quiet.state[StateValue.ITERATION] == 4 and not (tmp_path / "quiet").exists()
