"""Enumerations shared by the whole package.

The names AND values mirror the reference's ``ssdn/ssdn/params.py`` because reference checkpoints
pickle these enums by value inside ``state_dict()["cfg"]`` (denoiser.py:399-403); keeping them makes
``models/*.wt`` files of the reference loadable here and vice versa.  Values the reference derives with
``auto()`` are written out explicitly so that reordering this file can never change them."""
from enum import Enum


class NoiseAlgorithm(Enum):
    SELFSUPERVISED_DENOISING = "ssdn"
    SELFSUPERVISED_DENOISING_MEAN_ONLY = "ssdn_u_only"
    NOISE_TO_NOISE = "n2n"
    NOISE_TO_CLEAN = "n2c"
    NOISE_TO_VOID = "n2v"


class NoiseValue(Enum):
    UNKNOWN_CONSTANT = "const"
    UNKNOWN_VARIABLE = "var"
    KNOWN = "known"


class Pipeline(Enum):
    MSE = "mse"
    SSDN = "ssdn"
    MASK_MSE = "mask_mse"


class Blindspot(Enum):
    ENABLED = "blindspot"
    DISABLED = "normal"


class ConfigValue(Enum):
    INFER_CFG = 1
    ALGORITHM = 2
    BLINDSPOT = 3
    PIPELINE = 4
    IMAGE_CHANNELS = 5
    NOISE_STYLE = 6
    LEARNING_RATE = 7
    LR_RAMPUP_FRACTION = 8
    LR_RAMPDOWN_FRACTION = 9
    NOISE_VALUE = 10
    DIAGONAL_COVARIANCE = 11
    EVAL_INTERVAL = 12
    PRINT_INTERVAL = 13
    SNAPSHOT_INTERVAL = 14
    TRAIN_ITERATIONS = 15
    DATALOADER_WORKERS = 16
    TRAIN_DATASET_NAME = 17
    TRAIN_DATASET_TYPE = 18
    TRAIN_DATA_PATH = 19
    TRAIN_PATCH_SIZE = 20
    TRAIN_MINIBATCH_SIZE = 21
    TEST_DATASET_NAME = 22
    TEST_DATASET_TYPE = 23
    TEST_DATA_PATH = 24
    TEST_MINIBATCH_SIZE = 25
    PIN_DATA_MEMORY = 26


class DatasetType(Enum):
    HDF5 = 1
    FOLDER = 2


class StateValue(Enum):
    INITIALISED = 1
    MODE = 2
    ITERATION = 3
    REFERENCE = 4
    HISTORY = 5


class HistoryValue(Enum):
    TRAIN = 1
    EVAL = 2
    TIMINGS = 3


class PipelineOutput(Enum):
    INPUTS = 1
    LOSS = "loss"
    IMG_DENOISED = "out"
    IMG_MU = "out_mu"
    NOISE_STD_DEV = "noise_std"
    MODEL_STD_DEV = "model_std"
