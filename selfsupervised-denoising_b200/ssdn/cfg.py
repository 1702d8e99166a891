"""Configuration dictionaries keyed by :class:`ssdn.params.ConfigValue` (reference: ssdn/ssdn/cfg.py).

Same defaults and inference rules as the reference so that run names and checkpoints are interchangeable."""
import os
from typing import Dict

from ssdn.params import ConfigValue, DatasetType, NoiseAlgorithm, Pipeline

DEFAULT_RUN_DIR = "runs"

# Defaults of the reference (cfg.py:10-38), keyed by ConfigValue member name.
# NB: train.py hands the two ramp fractions to compute_ramped_lrate in swapped positions (see ssdn.train.learning_rate);
# the stored numbers are kept identical to the reference so the EFFECTIVE schedule is identical too.
_DEFAULTS = {getattr(ConfigValue, name): value for name, value in dict(
    TRAIN_ITERATIONS=2000000, TRAIN_MINIBATCH_SIZE=4, TEST_MINIBATCH_SIZE=2, IMAGE_CHANNELS=3, TRAIN_PATCH_SIZE=64, LEARNING_RATE=3e-4,
    LR_RAMPDOWN_FRACTION=0.1, LR_RAMPUP_FRACTION=0.3, EVAL_INTERVAL=10000, PRINT_INTERVAL=1000, SNAPSHOT_INTERVAL=10000,
    DATALOADER_WORKERS=4, PIN_DATA_MEMORY=False, DIAGONAL_COVARIANCE=False, TRAIN_DATA_PATH=None, TRAIN_DATASET_TYPE=None,
    TRAIN_DATASET_NAME=None, TEST_DATA_PATH=None, TEST_DATASET_TYPE=None, TEST_DATASET_NAME=None).items()}


def base() -> Dict:
    return dict(_DEFAULTS)


class DatasetName:
    BSD = "bsd"
    IMAGE_NET = "ilsvrc"
    KODAK = "kodak"
    SET14 = "set14"


_PATH_HINTS = {"bsds300": DatasetName.BSD, "ilsvrc": DatasetName.IMAGE_NET, "kodak": DatasetName.KODAK, "set14": DatasetName.SET14}


def _dataset_name_from_path(path: str) -> str:
    hits = [name for hint, name in _PATH_HINTS.items() if hint in path.lower()]
    if not hits:
        raise ValueError("Could not infer dataset from path.")
    if len(hits) > 1:
        raise ValueError("Matched multiple datasets with dataset path.")
    return hits[0]


def infer_datasets(cfg: Dict):
    """Fill in dataset name (from a known substring of the path) and type (directory -> folder, file -> HDF5)."""
    for path_k, name_k, type_k in ((ConfigValue.TRAIN_DATA_PATH, ConfigValue.TRAIN_DATASET_NAME, ConfigValue.TRAIN_DATASET_TYPE),
                                   (ConfigValue.TEST_DATA_PATH, ConfigValue.TEST_DATASET_NAME, ConfigValue.TEST_DATASET_TYPE)):
        path = cfg.get(path_k)
        if path is None:
            continue
        if cfg.get(name_k) is None:
            cfg[name_k] = _dataset_name_from_path(path)
        if cfg.get(type_k) is None:
            cfg[type_k] = DatasetType.FOLDER if os.path.isdir(path) else DatasetType.HDF5


def test_length(dataset_name: str) -> int:
    """Images to evaluate so that every test set contributes a similar pixel count."""
    return {DatasetName.BSD: 300, DatasetName.KODAK: 240, DatasetName.SET14: 280}[dataset_name]


def infer_pipeline(algorithm: NoiseAlgorithm) -> Pipeline:
    table = {NoiseAlgorithm.SELFSUPERVISED_DENOISING: Pipeline.SSDN,
             NoiseAlgorithm.SELFSUPERVISED_DENOISING_MEAN_ONLY: Pipeline.MSE,
             NoiseAlgorithm.NOISE_TO_NOISE: Pipeline.MSE, NoiseAlgorithm.NOISE_TO_CLEAN: Pipeline.MSE,
             NoiseAlgorithm.NOISE_TO_VOID: Pipeline.MASK_MSE}
    if algorithm not in table:
        raise NotImplementedError("Algorithm does not have a default pipeline.")
    return table[algorithm]


def infer_blindspot(algorithm: NoiseAlgorithm) -> bool:
    if algorithm in (NoiseAlgorithm.SELFSUPERVISED_DENOISING, NoiseAlgorithm.SELFSUPERVISED_DENOISING_MEAN_ONLY):
        return True
    if algorithm in (NoiseAlgorithm.NOISE_TO_NOISE, NoiseAlgorithm.NOISE_TO_CLEAN, NoiseAlgorithm.NOISE_TO_VOID):
        return False
    raise NotImplementedError("Not known if algorithm requires blindspot.")


def infer(cfg: Dict, model_only: bool = False) -> Dict:
    if cfg.get(ConfigValue.PIPELINE) is None:
        cfg[ConfigValue.PIPELINE] = infer_pipeline(cfg[ConfigValue.ALGORITHM])
    if cfg.get(ConfigValue.BLINDSPOT) is None:
        cfg[ConfigValue.BLINDSPOT] = infer_blindspot(cfg[ConfigValue.ALGORITHM])
    if not model_only:
        infer_datasets(cfg)
    return cfg


def config_name(cfg: Dict) -> str:
    """'<algorithm>[-<pipeline>_pipeline][-blindspot...]-<noise style>[-sigma_<mode>][-mono][-diag]'"""
    cfg = infer(cfg)
    algo = cfg[ConfigValue.ALGORITHM]
    parts = [algo.value]
    if cfg[ConfigValue.PIPELINE] != infer_pipeline(algo):
        parts.append(cfg[ConfigValue.PIPELINE].value + "_pipeline")
    if cfg[ConfigValue.BLINDSPOT] != infer_blindspot(algo):
        parts.append("blindspot" if cfg[ConfigValue.BLINDSPOT] else "blindspot_disabled")
    parts.append(cfg[ConfigValue.NOISE_STYLE])
    ssdn_pipe = cfg[ConfigValue.PIPELINE] == Pipeline.SSDN
    if ssdn_pipe:
        parts.append("sigma_" + cfg[ConfigValue.NOISE_VALUE].value)
    if cfg[ConfigValue.IMAGE_CHANNELS] == 1:
        parts.append("mono")
    if ssdn_pipe and cfg[ConfigValue.DIAGONAL_COVARIANCE]:
        parts.append("diag")
    return "-".join(parts)
