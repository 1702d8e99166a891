"""Differential battery for the host-side interface (CPU only; run as a subprocess).

    python tests/host_battery.py reference     # imports the UNMODIFIED reference package (build container only)
    python tests/host_battery.py mirror        # imports this repo's drop-in package of the same name

The battery below is written ONCE against the reference's public API (module paths, call signatures, metadata keys)
and runs unchanged on both packages; tests/test_host_differential.py compares the two JSON documents.  Everything here
is host logic next to the hot path: configuration inference and naming, the learning-rate ramp, metric bookkeeping,
image helpers, the noisy-dataset wrapper (padding, metadata, RNG consumption), samplers, Noise2Void masking and loss,
shift / crop modules, the network's parameter schema and seeded initialisation, the denoiser's state-dict schema."""
import hashlib
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
MODE = sys.argv[1]
if MODE == "reference":
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from _ref_shim import import_reference
    ssdn = import_reference()
else:
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "selfsupervised-denoising_b200"))
    import ssdn

from ssdn.datasets import FixedLengthSampler, NoisyDataset, SamplingOrder  # noqa: E402
from ssdn.params import ConfigValue, NoiseAlgorithm, NoiseValue, Pipeline  # noqa: E402

torch.set_num_threads(2)
OUT = {}


def plain(v):
    """JSON-able, order-preserving description of a value."""
    if torch.is_tensor(v):
        v = v.detach().contiguous()
        d = {"shape": list(v.shape), "dtype": str(v.dtype), "sha256": hashlib.sha256(v.numpy().tobytes()).hexdigest()[:24],
             "sum": float(v.double().sum())}
        if v.numel() <= 16:
            d["data"] = v.double().flatten().tolist()
        return d
    if isinstance(v, (list, tuple)):
        return [plain(x) for x in v]
    if isinstance(v, dict):
        return {str(getattr(k, "name", k)): plain(x) for k, x in v.items()}
    if hasattr(v, "name") and hasattr(v, "value"):
        return "{}.{}".format(type(v).__name__, v.name)
    if isinstance(v, (int, float, str, bool)) or v is None:
        return v
    return repr(v)


def attempt(fn):
    try:
        return plain(fn())
    except Exception as e:                    # the error TYPE is part of the interface
        return "raises " + type(e).__name__


# ---------------------------------------------------------------- configuration
def cfg_for(algorithm, noise_value, style, channels):
    cfg = ssdn.cfg.base()
    cfg[ConfigValue.ALGORITHM] = algorithm
    cfg[ConfigValue.NOISE_VALUE] = noise_value
    cfg[ConfigValue.NOISE_STYLE] = style
    cfg[ConfigValue.IMAGE_CHANNELS] = channels
    ssdn.cfg.infer(cfg, model_only=True)
    return cfg


OUT["cfg.base"] = plain(ssdn.cfg.base())
OUT["cfg.DEFAULT_RUN_DIR"] = ssdn.cfg.DEFAULT_RUN_DIR
for alg in NoiseAlgorithm:
    OUT["cfg.infer_pipeline." + alg.name] = attempt(lambda: ssdn.cfg.infer_pipeline(alg))
    OUT["cfg.infer_blindspot." + alg.name] = attempt(lambda: ssdn.cfg.infer_blindspot(alg))
    for nv in NoiseValue:
        for style in ("gauss25", "gauss5_50_nc", "poisson30"):
            cfg = cfg_for(alg, nv, style, 3)
            OUT["cfg.infer.{}.{}.{}".format(alg.name, nv.name, style)] = plain(cfg)
            OUT["cfg.name.{}.{}.{}".format(alg.name, nv.name, style)] = ssdn.cfg.config_name(cfg)


def datasets_cfg(train, test):
    cfg = cfg_for(NoiseAlgorithm.SELFSUPERVISED_DENOISING, NoiseValue.KNOWN, "gauss25", 3)
    cfg[ConfigValue.TRAIN_DATA_PATH], cfg[ConfigValue.TEST_DATA_PATH] = train, test
    ssdn.cfg.infer_datasets(cfg)
    return {k.name: cfg[k] for k in (ConfigValue.TRAIN_DATASET_NAME, ConfigValue.TRAIN_DATASET_TYPE, ConfigValue.TEST_DATASET_NAME,
                                     ConfigValue.TEST_DATASET_TYPE)}


for train, test in (("/data/ilsvrc_val.h5", "/data/kodak"), ("/x/BSDS300/images/train", "/x/set14.h5"), (None, "/data/Kodak/"),
                    ("/data/ilsvrc.hdf5", None), ("/data/unknown_things", "/data/kodak")):
    OUT["cfg.infer_datasets.{}.{}".format(train, test)] = attempt(lambda: datasets_cfg(train, test))
for name in ("kodak", "bsds300", "set14", "ilsvrc", "nothing"):
    OUT["cfg.test_length." + name] = attempt(lambda: ssdn.cfg.test_length(name))

# ---------------------------------------------------------------- bookkeeping helpers
OUT["lr"] = [[ssdn.utils.compute_ramped_lrate(i, 1000, up, down, 3e-4) for i in range(0, 1001, 25)]
             for up, down in ((0.1, 0.3), (0.3, 0.1), (0.0, 0.0), (0.5, 0.5), (0.0, 1.0))]
OUT["dhms"] = [ssdn.utils.seconds_to_dhms(s, trim) for s in (0, 0.4, 1, 59.99, 60, 61, 3599, 3600, 86399, 86400, 90061.7, 1e7) for trim in (True, False)]
OUT["separator"] = [ssdn.utils.separator(), ssdn.utils.separator(7)]


def metric_run(batched, collapse, values):
    m = ssdn.utils.Metric(batched=batched, collapse=collapse)
    seen = [m.empty(), m.accumulated()]
    for v in values:
        m += v
        seen.append(m.accumulated())
    seen.append([m.n, m.accumulated(reset=True), m.empty(), m.n])
    return seen


g = torch.Generator().manual_seed(0)
vals = [torch.rand(4, 3, 2, generator=g), torch.rand(2, 3, 2, generator=g), torch.rand(1, 3, 2, generator=g)]
for batched in (True, False):
    for collapse in (True, False):
        OUT["metric.{}.{}".format(batched, collapse)] = attempt(lambda: metric_run(batched, collapse, vals if batched else [v[0] for v in vals]))
md = ssdn.utils.MetricDict()
md["a"] += torch.ones(3)
md["b"] += torch.zeros(2, 2)
OUT["metricdict"] = [list(md.keys()), md["a"].n, md["b"].n, plain(md["b"].accumulated())]
t = ssdn.utils.TrackedTime()
OUT["trackedtime"] = [t.total, t.last_time, sorted(vars(t))]

# ---------------------------------------------------------------- image helpers
x = torch.rand(2, 3, 4, 4, generator=g) * 1.4 - 0.2
OUT["clip.float"] = plain(ssdn.utils.clip_img(x))
OUT["clip.uint8like"] = attempt(lambda: ssdn.utils.clip_img((x * 300).to(torch.int32)))
sq = torch.arange(2 * 3 * 4 * 4, dtype=torch.float32).reshape(2, 3, 4, 4)
for angle in (0, 90, 180, 270, 45, -90, 360):
    OUT["rotate.{}".format(angle)] = attempt(lambda: ssdn.utils.rotate(sq, angle).contiguous())
OUT["mse2psnr"] = plain(ssdn.utils.mse2psnr(torch.tensor([1e-4, 0.01, 0.5])))
OUT["mse2psnr.int"] = attempt(lambda: ssdn.utils.mse2psnr(torch.tensor([1.0, 100.0]), False))
a, b = torch.rand(3, 3, 8, 8, generator=g), torch.rand(3, 3, 8, 8, generator=g)
OUT["psnr"] = plain(ssdn.utils.calculate_psnr(a, b))
OUT["psnr.chw"] = attempt(lambda: ssdn.utils.calculate_psnr(a[0], b[0], "CHW"))


# ---------------------------------------------------------------- noisy dataset wrapper
class Images:
    def __init__(self, shapes):
        gg = torch.Generator().manual_seed(3)
        self.items = [torch.rand(*s, generator=gg) for s in shapes]

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i], i


shapes = [(3, 40, 56), (3, 64, 33), (3, 32, 32)]
for alg in NoiseAlgorithm:
    for style in ("gauss25", "gauss5_50", "gauss10_nc", "poisson30", "poisson5_50"):
        for kwargs in ({}, {"pad_multiple": 32, "square": True}, {"pad_uniform": True, "pad_multiple": 32}, {"training_mode": True}):
            def run():
                torch.manual_seed(5)
                ds = NoisyDataset(Images(shapes), style, alg, **kwargs)
                rows = []
                for i in range(len(ds)):
                    inp, ref, meta = ds[i]
                    rows.append([inp, ref, meta])
                    if i == 1:
                        un = NoisyDataset.unpad(inp, meta)
                        rows.append(list(un.shape))
                return rows
            OUT["dataset.{}.{}.{}".format(alg.name, style, sorted(kwargs))] = attempt(run)
OUT["dataset.badstyle"] = attempt(lambda: NoisyDataset(Images(shapes), "speckle4", NoiseAlgorithm.NOISE_TO_CLEAN)[0])
ds = NoisyDataset(Images(shapes), "gauss25", NoiseAlgorithm.NOISE_TO_CLEAN, pad_multiple=32, square=True)
loader = torch.utils.data.DataLoader(NoisyDataset(Images([(3, 32, 32)] * 4), "gauss25", NoiseAlgorithm.SELFSUPERVISED_DENOISING), batch_size=2)
torch.manual_seed(9)
batch = next(iter(loader))
OUT["dataset.collated"] = plain([batch[0].shape, batch[1].shape, {k.name: list(v.shape) for k, v in batch[2].items()}])
OUT["dataset.unpad.batch"] = attempt(lambda: [list(t.shape) for t in NoisyDataset.unpad(batch[0], batch[2])])
OUT["dataset.unpad.index"] = attempt(lambda: list(NoisyDataset.unpad(batch[0], batch[2], 1).shape))
OUT["dataset.constants"] = [NoisyDataset.INPUT, NoisyDataset.REFERENCE, NoisyDataset.METADATA, [(m.name, m.value) for m in NoisyDataset.Metadata]]

# ---------------------------------------------------------------- samplers
for shuffled in (False, True):
    for num in (None, 3, 10, 25):
        torch.manual_seed(1)
        s = FixedLengthSampler(list(range(10)), num_samples=num, shuffled=shuffled)
        OUT["sampler.{}.{}".format(shuffled, num)] = [len(s), list(iter(s))]
torch.manual_seed(2)
s = FixedLengthSampler(list(range(6)), num_samples=14, shuffled=True)
it = iter(s)
head = [next(it) for _ in range(5)]
state = s.last_iter().state_dict()
s2 = FixedLengthSampler(list(range(6)), num_samples=14, shuffled=True)
s2.for_next_iter(SamplingOrder.from_state_dict(state))
OUT["sampler.resume"] = [head, plain(state), list(iter(s2)), len(s2.last_iter())]

# ---------------------------------------------------------------- Noise2Void masking and loss
torch.manual_seed(4)
img = torch.rand(3, 32, 32)
masked, coords = ssdn.utils.n2v_ups.manipulate(img, 5)
OUT["n2v.manipulate"] = plain([masked, coords, torch.equal(img, masked)])
OUT["n2v.manipulate.even"] = attempt(lambda: ssdn.utils.n2v_ups.manipulate(img, 4))
out_img, ref_img = torch.rand(2, 3, 32, 32), torch.rand(2, 3, 32, 32)
OUT["n2v.loss"] = attempt(lambda: ssdn.utils.n2v_loss.loss_mask_mse(torch.stack([coords, coords]), out_img, ref_img))

# ---------------------------------------------------------------- noise styles
for style in ("gauss25", "gauss25_nc", "gauss5_50", "gauss0.1", "gauss0.05_0.2_nc", "poisson30", "poisson5_50_nc", "speckle3"):
    def run():
        torch.manual_seed(6)
        noisy, coeff = ssdn.utils.noise.add_style(torch.rand(2, 3, 8, 8), style)
        return [noisy, coeff if not torch.is_tensor(coeff) else coeff]
    OUT["noise." + style] = attempt(run)

# ---------------------------------------------------------------- shift / crop modules, network and denoiser schemas
from ssdn.models import Crop2d, NoiseNetwork, Shift2d  # noqa: E402

z = torch.arange(2 * 1 * 4 * 5, dtype=torch.float32).reshape(2, 1, 4, 5)
for shift in ((1, 0), (0, 1), (-1, 0), (0, -2), (2, 1)):
    OUT["shift2d.{}".format(shift)] = attempt(lambda: Shift2d(shift)(z))
for crop in ((0, 0, 1, 0), (1, 1, 0, 0), (0, 2, 0, 1)):
    OUT["crop2d.{}".format(crop)] = attempt(lambda: Crop2d(crop)(z))
for blind, cin, cout, zero in ((True, 3, 9, False), (False, 3, 1, True), (False, 1, 1, False), (True, 1, 2, False)):
    torch.manual_seed(0)
    net = NoiseNetwork(in_channels=cin, out_channels=cout, blindspot=blind, zero_output_weights=zero)
    sd = net.state_dict()
    OUT["net.{}.{}.{}.{}".format(blind, cin, cout, zero)] = {
        "keys": {k: list(v.shape) for k, v in sd.items()},
        "param_order": [n for n, _ in net.named_parameters()],
        "sums": {k: float(v.double().sum()) for k, v in sd.items()},
        "abs_sums": {k: float(v.double().abs().sum()) for k, v in sd.items()},
        "blindspot": net.blindspot, "wh_mul": NoiseNetwork.input_wh_mul(),
        "rng_after": float(torch.rand(1)),
    }
from ssdn.denoiser import Denoiser  # noqa: E402

for alg, nv in ((NoiseAlgorithm.SELFSUPERVISED_DENOISING, NoiseValue.KNOWN), (NoiseAlgorithm.SELFSUPERVISED_DENOISING, NoiseValue.UNKNOWN_CONSTANT),
                (NoiseAlgorithm.SELFSUPERVISED_DENOISING, NoiseValue.UNKNOWN_VARIABLE), (NoiseAlgorithm.NOISE_TO_VOID, NoiseValue.KNOWN),
                (NoiseAlgorithm.SELFSUPERVISED_DENOISING_MEAN_ONLY, NoiseValue.KNOWN)):
    for ch in (1, 3):
        torch.manual_seed(0)
        den = Denoiser(cfg_for(alg, nv, "gauss25", ch), device="cpu")
        sd = den.state_dict()
        OUT["denoiser.{}.{}.{}".format(alg.name, nv.name, ch)] = {
            "keys": {k: (list(v.shape) if torch.is_tensor(v) else type(v).__name__) for k, v in sd.items()},
            "params_only": sorted(den.state_dict(params_only=True)),
            "n_parameters": len(list(den.parameters())), "numel": sum(p.numel() for p in den.parameters()),
            "config_name": den.config_name(),
            "constants": [Denoiser.MODEL, Denoiser.SIGMA_ESTIMATOR, Denoiser.ESTIMATED_SIGMA],
            "sum": float(sum(p.double().sum() for p in den.parameters())),
        }
OUT["pipeline.enum"] = [(p.name, p.value) for p in Pipeline]

# ---------------------------------------------------------------- axis-order strings, PIL conversion
from ssdn.utils import data_format as DF  # noqa: E402

OUT["data_format"] = [[DF.permute_tuple(a, b) for a, b in (("CWH", "CHW"), ("BCHW", "BHWC"), ("HWC", "CHW"), ("CHW", "CHW"), ("BCWH", "BCHW"))],
                      DF.batch("CHW"), DF.batch("BCHW"), DF.unbatch("BHWC"), DF.PIL_FORMAT, DF.PIL_BATCH_FORMAT,
                      sorted(k for k in vars(DF.DataFormat) if k.isupper())]
OUT["data_format.bad"] = attempt(lambda: DF.permute_tuple("CHW", "BCHW"))


def pil_plain(im):
    return {"size": list(im.size), "mode": im.mode, "sha256": hashlib.sha256(im.tobytes()).hexdigest()[:24]}


gi = torch.Generator().manual_seed(8)
OUT["tensor2image.rgb"] = pil_plain(ssdn.utils.tensor2image(torch.rand(3, 6, 9, generator=gi) * 1.2 - 0.1))
OUT["tensor2image.grey"] = pil_plain(ssdn.utils.tensor2image(torch.rand(1, 6, 9, generator=gi)))
OUT["tensor2image.hwc"] = pil_plain(ssdn.utils.tensor2image(torch.rand(6, 9, 3, generator=gi), "HWC"))
OUT["tensor2image.batch"] = pil_plain(ssdn.utils.tensor2image(torch.rand(3, 3, 6, 9, generator=gi)))      # a batch becomes a grid
OUT["tensor2image.two_channels"] = attempt(lambda: ssdn.utils.tensor2image(torch.rand(2, 6, 9)))

# ---------------------------------------------------------------- folder data set and the loaders the trainer builds from a configuration
import tempfile  # noqa: E402

from PIL import Image  # noqa: E402

from ssdn.datasets import UnlabelledImageFolderDataset  # noqa: E402

if MODE == "reference":        # the header-only size reader the reference imports is absent here; give the stub the same contract
    sys.modules["imagesize"].get = lambda path: Image.open(path).size
root = tempfile.mkdtemp(prefix="battery_")
kodak = os.path.join(root, "kodak_synth")
os.makedirs(os.path.join(kodak, "more"))
gi = torch.Generator().manual_seed(12)


def write(path, size, mode):
    w, h = size
    bands = {"RGB": 3, "L": 1, "RGBA": 4}[mode]
    arr = torch.randint(0, 256, (h, w, bands), generator=gi, dtype=torch.uint8).numpy()
    Image.fromarray(arr.squeeze() if bands == 1 else arr, mode=mode).save(path)


write(os.path.join(kodak, "b.png"), (56, 40), "RGB")
write(os.path.join(kodak, "a.PNG"), (33, 64), "RGB")
write(os.path.join(kodak, "c.bmp"), (32, 32), "L")
write(os.path.join(kodak, "more", "d.png"), (24, 20), "RGBA")
open(os.path.join(kodak, "notes.txt"), "w").write("not an image")
for recursive in (False, True):
    for channels in (3, 1):
        fd = UnlabelledImageFolderDataset(kodak, recursive=recursive, channels=channels)
        OUT["folder.{}.{}".format(recursive, channels)] = [len(fd), [os.path.relpath(f, kodak) for f in fd.files], [plain(fd[i]) for i in range(len(fd))],
                                                          [plain(fd.image_size(i)) for i in range(len(fd))]]
OUT["folder.empty"] = attempt(lambda: UnlabelledImageFolderDataset(os.path.join(kodak, "more"), extensions=[".jpg"]))
OUT["set_color_channels"] = [pil_plain(ssdn.utils.set_color_channels(Image.open(os.path.join(kodak, f)), ch)) for f in ("b.png", "c.bmp") for ch in (1, 3)]
from torchvision.transforms import RandomCrop  # noqa: E402

torch.manual_seed(21)
fd = UnlabelledImageFolderDataset(kodak, recursive=True, transform=RandomCrop(48, pad_if_needed=True, padding_mode="reflect"))
OUT["folder.randomcrop"] = [[plain(fd[i]) for i in range(len(fd))], plain(fd.image_size(1)), plain(fd.image_size(1, ignore_transform=True))]
torch.manual_seed(22)
OUT["noise_transform"] = plain(ssdn.utils.transforms.NoiseTransform("gauss25")(torch.rand(2, 3, 8, 8)))

from ssdn.train import DenoiserTrainer  # noqa: E402

for alg, nv in ((NoiseAlgorithm.SELFSUPERVISED_DENOISING, NoiseValue.KNOWN), (NoiseAlgorithm.NOISE_TO_VOID, NoiseValue.KNOWN)):
    def loaders():
        cfg = ssdn.cfg.base()
        cfg[ConfigValue.ALGORITHM], cfg[ConfigValue.NOISE_VALUE], cfg[ConfigValue.NOISE_STYLE] = alg, nv, "gauss25"
        cfg[ConfigValue.TRAIN_DATA_PATH] = cfg[ConfigValue.TEST_DATA_PATH] = kodak
        cfg[ConfigValue.DATALOADER_WORKERS] = 0
        cfg[ConfigValue.TRAIN_ITERATIONS], cfg[ConfigValue.TRAIN_MINIBATCH_SIZE], cfg[ConfigValue.TEST_MINIBATCH_SIZE] = 10, 3, 2
        trainer = DenoiserTrainer(cfg, state={}, runs_dir=root)
        described = [plain({k.name: trainer.cfg[k] for k in (ConfigValue.TRAIN_DATASET_NAME, ConfigValue.TRAIN_DATASET_TYPE,
                                                              ConfigValue.TEST_DATASET_NAME, ConfigValue.TEST_DATASET_TYPE)})]
        torch.manual_seed(31)
        loader, dataset, sampler = trainer.train_data()
        described.append([len(loader), len(dataset), len(sampler), [plain(list(b)) for b in loader]])
        torch.manual_seed(32)
        loader, dataset, sampler = trainer.test_data()
        it = iter(loader)
        described.append([len(loader), len(dataset), len(sampler), plain(dataset.max_image_size), [plain(list(next(it))) for _ in range(3)]])
        trainer.set_test_data(os.path.join(root, "kodak_synth", "more"))
        described.append(plain({k.name: trainer.cfg[k] for k in (ConfigValue.TEST_DATASET_NAME, ConfigValue.TEST_DATASET_TYPE)}))
        described.append(os.path.relpath(trainer.cfg[ConfigValue.TEST_DATA_PATH], root))
        return described
    OUT["trainer.loaders." + alg.name] = attempt(loaders)

import shutil  # noqa: E402

shutil.rmtree(root, ignore_errors=True)
print("JSON " + json.dumps(OUT))
