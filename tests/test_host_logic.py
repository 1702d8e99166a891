"""CPU: host-side mirror of the reference interface (config, enums, module / state-dict schema, schedule, data plumbing)."""
import os
import pickle

import pytest
import torch

import ssdn
import ssdn_oracle as O
from ssdn.datasets import FixedLengthSampler, NoisyDataset, SamplingOrder
from ssdn.models import NoiseNetwork, Shift2d, Crop2d
from ssdn.params import ConfigValue, DatasetType, NoiseAlgorithm, NoiseValue, Pipeline, PipelineOutput, StateValue, HistoryValue
from util import make_cfg


def test_enum_values_match_reference_pickles():
    # values the reference derives with auto(): pickled by value inside checkpoints (denoiser.py:399-403)
    assert [e.value for e in ConfigValue] == list(range(1, 27))
    assert ConfigValue.TRAIN_ITERATIONS.value == 15 and ConfigValue.PIN_DATA_MEMORY.value == 26
    assert DatasetType.HDF5.value == 1 and StateValue.ITERATION.value == 3 and HistoryValue.TIMINGS.value == 3
    assert PipelineOutput.INPUTS.value == 1 and PipelineOutput.IMG_DENOISED.value == "out"
    assert NoiseValue("var") is NoiseValue.UNKNOWN_VARIABLE and NoiseAlgorithm("n2v") is NoiseAlgorithm.NOISE_TO_VOID
    assert pickle.loads(pickle.dumps(ConfigValue.NOISE_STYLE)) is ConfigValue.NOISE_STYLE


def test_cfg_inference_and_names():
    cfg = make_cfg("ssdn", "known")
    assert cfg[ConfigValue.PIPELINE] == Pipeline.SSDN and cfg[ConfigValue.BLINDSPOT] is True
    assert ssdn.cfg.config_name(cfg) == "ssdn-gauss25-sigma_known"
    assert ssdn.cfg.config_name(make_cfg("ssdn", "var", channels=1)) == "ssdn-gauss25-sigma_var-mono"
    cfg = make_cfg("n2v")
    assert cfg[ConfigValue.PIPELINE] == Pipeline.MASK_MSE and cfg[ConfigValue.BLINDSPOT] is False
    assert ssdn.cfg.config_name(cfg) == "n2v-gauss25"
    base = ssdn.cfg.base()
    assert base[ConfigValue.TRAIN_MINIBATCH_SIZE] == 4 and base[ConfigValue.TRAIN_PATCH_SIZE] == 64 and base[ConfigValue.LEARNING_RATE] == 3e-4
    with pytest.raises(ValueError):
        ssdn.cfg.infer({ConfigValue.ALGORITHM: NoiseAlgorithm.NOISE_TO_CLEAN, ConfigValue.TRAIN_DATA_PATH: "/data/unknown.h5"})
    c = {ConfigValue.ALGORITHM: NoiseAlgorithm.NOISE_TO_CLEAN, ConfigValue.TRAIN_DATA_PATH: "/data/ilsvrc_val.h5"}
    ssdn.cfg.infer(c)
    assert c[ConfigValue.TRAIN_DATASET_NAME] == "ilsvrc" and c[ConfigValue.TRAIN_DATASET_TYPE] == DatasetType.HDF5


def test_learning_rate_schedule_effective_ramps():
    cfg = make_cfg()
    its = cfg[ConfigValue.TRAIN_ITERATIONS]
    lr = lambda i: ssdn.train.learning_rate(cfg, i)  # noqa: E731
    assert lr(0) == 0 and abs(lr(its // 20) - 1.5e-4) < 1e-9 and abs(lr(its // 10) - 3e-4) < 1e-12
    assert abs(lr(int(its * 0.7)) - 3e-4) < 1e-12 and abs(lr(int(its * 0.85)) - 7.5e-5) < 1e-9 and lr(its) < 1e-12
    for i in (0, 1234, its // 3, its - 5):
        assert abs(lr(i) - O.effective_lrate(i, its)) < 1e-15


@pytest.mark.parametrize("blind,cin,cout,count", [(True, 3, 9, 1269129), (True, 1, 2, 1265858), (False, 3, 1, 1102177),
                                                   (False, 3, 3, 1102371), (False, 1, 1, 1099585)])
def test_noise_network_schema_and_init(blind, cin, cout, count):
    torch.manual_seed(3)
    net = NoiseNetwork(cin, cout, blindspot=blind)
    torch.manual_seed(3)
    ref = O.init_params(cin, cout, blind)                       # oracle init == reference init (pinned in make_golden.py)
    assert sum(p.numel() for p in net.parameters()) == count
    assert [n for n, _ in net.named_parameters()] == O.param_order(cin, cout, blind)
    sd = net.state_dict()
    assert set(sd) == set(ref) | {"output_block.4.weight", "output_block.4.bias"}
    assert all(torch.equal(sd[k], ref[k]) for k in ref)
    assert net.blindspot is blind and NoiseNetwork.input_wh_mul() == 32
    flat = net.flat_parameters()
    assert flat.numel() == count and torch.equal(flat[: ref["encode_block_1.0.weight"].numel()], ref["encode_block_1.0.weight"].reshape(-1))
    net.encode_block_1[0].weight.data.add_(1.0)                 # parameters are views of the flat buffer
    assert torch.equal(flat[:10], (ref["encode_block_1.0.weight"].reshape(-1) + 1.0)[:10])
    net.load_state_dict({k: v.clone() for k, v in ref.items()}, strict=False)
    assert torch.equal(net.flat_parameters()[:10], ref["encode_block_1.0.weight"].reshape(-1)[:10])


def test_noise_network_refuses_cpu():
    from ssdn._engine import EngineError
    with pytest.raises(EngineError):
        NoiseNetwork(3, 3)(torch.rand(1, 3, 32, 32))


def test_denoiser_state_dict_schema_and_roundtrip(tmp_path):
    cfg = make_cfg("ssdn", "var")
    den = ssdn.Denoiser(cfg, device="cpu")
    assert sum(p.numel() for p in den.parameters()) == 2371306
    sd = den.state_dict()
    assert "cfg" in sd and sd["cfg"] is cfg
    assert "models.denoiser_model.module.encode_block_1.0.weight" in sd and "_models.denoiser_model.output_conv.bias" in sd
    assert "models.sigma_estimation_model.module.output_block.4.weight" in sd
    assert "cfg" not in den.state_dict(params_only=True)
    path = tmp_path / "m.wt"
    torch.save(sd, path)
    den2 = ssdn.Denoiser.from_state_dict(torch.load(path, map_location="cpu", weights_only=False), device="cpu")
    for (k1, p1), (k2, p2) in zip(den.named_parameters(), den2.named_parameters()):
        assert k1 == k2 and torch.equal(p1, p2)
    const = ssdn.Denoiser(make_cfg("ssdn", "const"), device="cpu")
    assert list(const.l_params) == ["estimated_sigma"] and const.l_params["estimated_sigma"].shape == (1, 1, 1, 1)
    assert [n for n, _ in const.named_parameters()][-1] == "l_params.estimated_sigma"
    flat = const.flat_parameters()
    assert flat.numel() == 1269130 and const.flat_gradients().numel() == 1269130
    assert ssdn.Denoiser.MODEL == "denoiser_model" and ssdn.Denoiser.SIGMA_ESTIMATOR == "sigma_estimation_model"


def test_unsupported_pipeline_raises():
    cfg = make_cfg("ssdn", "known")
    cfg[ConfigValue.PIPELINE] = "nonsense"
    den = ssdn.Denoiser.__new__(ssdn.Denoiser)
    torch.nn.Module.__init__(den)
    den.cfg = cfg
    den.device = torch.device("cpu")
    with pytest.raises(NotImplementedError):
        den.run_pipeline([torch.zeros(1, 3, 32, 32)])


def test_shift_and_crop_modules():
    x = torch.rand(2, 3, 6, 7)
    assert torch.equal(Shift2d((1, 0))(x), O.shift2d(x, 1, 0)) and torch.equal(Shift2d((0, -2))(x), O.shift2d(x, 0, -2))
    assert torch.equal(Crop2d((1, 2, 0, 3))(x), x[:, :, 0:3, 1:5])
    for a in (0, 90, 180, 270):
        assert torch.equal(ssdn.utils.rotate(x[..., :6], a), O.rotate(x[..., :6], a))
    with pytest.raises(NotImplementedError):
        ssdn.utils.rotate(x, 45)


class _Index(torch.utils.data.Dataset):
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return i


def test_sampler_sequential_and_shuffled():
    ds = _Index(5)
    assert list(FixedLengthSampler(ds, 12)) == [0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 0, 1]
    assert list(FixedLengthSampler(ds, 3)) == [0, 1, 2] and list(FixedLengthSampler(ds)) == [0, 1, 2, 3, 4]
    order = list(FixedLengthSampler(ds, 13, shuffled=True))
    assert len(order) == 13 and sorted(order[:5]) == [0, 1, 2, 3, 4] and sorted(order[5:10]) == [0, 1, 2, 3, 4]
    loader = torch.utils.data.DataLoader(ds, batch_size=2, drop_last=True, sampler=FixedLengthSampler(ds, 7))
    assert [b.tolist() for b in loader] == [[0, 1], [2, 3], [4, 0]]


def test_sampler_order_resume():
    ds = _Index(10)
    s = FixedLengthSampler(ds, 25, shuffled=True)
    it = iter(s)
    first = [next(it) for _ in range(7)]
    saved = s.last_iter().state_dict()
    rest = list(it)
    s2 = FixedLengthSampler(ds, 25, shuffled=True)
    s2.for_next_iter(SamplingOrder.from_state_dict(saved))
    assert list(iter(s2)) == rest and len(first) + len(rest) == 25


class _Images(torch.utils.data.Dataset):
    def __init__(self, shapes):
        self.items = [torch.rand(*s) for s in shapes]

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return (self.items[i], i)


def test_noisy_dataset_triples_and_padding():
    torch.manual_seed(0)
    M = NoisyDataset.Metadata
    ds = NoisyDataset(_Images([(3, 40, 50), (3, 33, 20)]), "gauss25", NoiseAlgorithm.SELFSUPERVISED_DENOISING, pad_uniform=True, pad_multiple=32,
                      square=True)
    inp, ref, md = ds[1]
    assert inp.shape == (3, 64, 64) and ref.numel() == 0 and md[M.CLEAN].shape == (3, 64, 64)
    assert md[M.IMAGE_SHAPE].tolist() == [3, 33, 20] and abs(float(md[M.INPUT_NOISE_VALUES]) - 25 / 255) < 1e-7
    assert torch.equal(NoisyDataset.unpad(md[M.CLEAN], md), ds.child[1][0])
    n2c = NoisyDataset(_Images([(1, 32, 32)]), "gauss25_nc", NoiseAlgorithm.NOISE_TO_CLEAN)
    inp, ref, md = n2c[0]
    assert torch.equal(ref, n2c.child[0][0]) and (inp < 0).any()                        # _nc: not clipped
    n2v = NoisyDataset(_Images([(3, 64, 64)]), "gauss25", NoiseAlgorithm.NOISE_TO_VOID, training_mode=True)
    inp, ref, md = n2v[0]
    assert md[M.MASK_COORDS].shape[1] == 2 and 40 <= md[M.MASK_COORDS].shape[0] <= 64 and ref.shape == inp.shape
    rng = NoisyDataset(_Images([(3, 32, 32)]), "gauss5_50", NoiseAlgorithm.SELFSUPERVISED_DENOISING)
    _, _, md = rng[0]
    assert md[M.INPUT_NOISE_VALUES].shape == (3, 1, 1)                                    # per-channel sigma
    assert (md[M.INPUT_NOISE_VALUES] >= 5 / 255).all() and (md[M.INPUT_NOISE_VALUES] <= 50 / 255).all()
    with pytest.raises(NotImplementedError):
        ssdn.utils.noise.add_style(torch.rand(1, 3, 8, 8), "speckle3")


def test_metric_accumulates_batch_means():
    m = ssdn.utils.Metric()
    m += torch.tensor([[1.0, 3.0], [5.0, 7.0]])
    m += torch.tensor([[9.0, 11.0]])
    assert m.n == 3 and abs(float(m.accumulated()) - 6.0) < 1e-6
    assert ssdn.utils.seconds_to_dhms(3661) == "01h01m01s"
    assert ssdn.utils.seconds_to_dhms(0) == "" and ssdn.utils.seconds_to_dhms(59.9) == "59s"
    assert ssdn.utils.seconds_to_dhms(90061.5) == "01d01h01m01s" and ssdn.utils.seconds_to_dhms(60, trim=False) == "00d00h01m00s"
    whole = ssdn.utils.Metric(batched=False)
    whole += torch.full((3, 2), 2.0)
    assert whole.n == 1 and float(whole.accumulated()) == 2.0
    keep = ssdn.utils.Metric(collapse=False)
    keep += torch.ones(4, 3)
    assert keep.accumulated(reset=True).tolist() == [1.0, 1.0, 1.0] and keep.empty() and keep.accumulated() is None
    history = ssdn.utils.MetricDict()
    history["psnr"] += torch.ones(2)
    assert list(history) == ["psnr"] and history["psnr"].n == 2


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference checkpoints only exist in the build container")
@pytest.mark.parametrize("name", ["final-ssdn-gauss25-sigma_known.wt", "final-ssdn-gauss25-sigma_var.wt", "final-ssdn-gauss25-sigma_const.wt",
                                  "final-n2v-gauss25.wt", "final-n2c-gauss25.wt"])
def test_reference_checkpoints_load_and_round_trip(name, tmp_path):
    """Checkpoint wire format: the reference's shipped ``.wt`` files (pickled ConfigValue enums,
    ``models.<id>.module.`` / ``_models.<id>.`` alias keys, learnable sigma) load into this package unchanged, every tensor
    survives, and what we save has exactly the reference's key set."""
    st = torch.load(os.path.join("/root/reference/models", name), map_location="cpu", weights_only=False)
    den = ssdn.Denoiser.from_state_dict(st, device="cpu")
    sd = den.state_dict()
    assert set(sd) == set(st)
    assert all(torch.equal(sd[k], v) for k, v in st.items() if torch.is_tensor(v))
    assert sd["cfg"] == st["cfg"]
    path = tmp_path / "again.wt"
    torch.save(sd, path)
    again = torch.load(path, map_location="cpu", weights_only=False)
    assert set(again) == set(st) and all(torch.equal(again[k], v) for k, v in st.items() if torch.is_tensor(v))


def test_image_cache_from_folder_dataset(tmp_path):
    """Folder of 8-bit images -> the uint8 cache of the on-GPU input pipeline, bit-exact; unequal sizes are refused."""
    import numpy as np
    from PIL import Image
    from ssdn.datasets import UnlabelledImageFolderDataset
    from ssdn.datasets.gpu_pipeline import image_cache_from_dataset
    rng = np.random.default_rng(0)
    arrs = [rng.integers(0, 256, (24, 40, 3), dtype=np.uint8) for _ in range(3)]
    for i, a in enumerate(arrs):
        Image.fromarray(a, mode="RGB").save(tmp_path / "img{}.png".format(i))
    ds = UnlabelledImageFolderDataset(str(tmp_path), output_format=None)        # to_tensor layout: C x H x W
    cache = image_cache_from_dataset(ds)
    assert cache.dtype == torch.uint8 and cache.shape == (3, 3, 24, 40)
    assert all(np.array_equal(cache[i].permute(1, 2, 0).numpy(), a) for i, a in enumerate(arrs))
    assert image_cache_from_dataset(ds, limit=2).shape[0] == 2
    Image.fromarray(rng.integers(0, 256, (30, 40, 3), dtype=np.uint8), mode="RGB").save(tmp_path / "odd.png")
    with pytest.raises(ValueError):
        image_cache_from_dataset(UnlabelledImageFolderDataset(str(tmp_path), output_format=None))
    import importlib.util
    if importlib.util.find_spec("h5py") is None:                                 # not installed in this image: a clear error, not a crash
        with pytest.raises(ImportError):
            ssdn.datasets.HDF5Dataset(str(tmp_path / "missing.h5"))
