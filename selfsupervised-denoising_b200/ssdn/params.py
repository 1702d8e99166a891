"""Enumerations shared by the whole package, built from one table.

Member names AND values are an interchange format: reference checkpoints pickle these enums (by module path, class name
and member name) inside ``state_dict()["cfg"]`` (denoiser.py:399-403) and run-directory names embed the string values, so
``ssdn.params`` must expose the same classes with the same members as the reference's ``ssdn/ssdn/params.py`` for the
shipped ``models/*.wt`` files to load here and for files written here to load there.  Integer members are numbered 1..N
in table order (what the reference's ``auto()`` yields); ``tests/test_host_logic.py`` pins every value and round-trips the
reference's checkpoints."""
from enum import Enum

_STRING_VALUED = {
    "NoiseAlgorithm": (("SELFSUPERVISED_DENOISING", "ssdn"), ("SELFSUPERVISED_DENOISING_MEAN_ONLY", "ssdn_u_only"), ("NOISE_TO_NOISE", "n2n"),
                       ("NOISE_TO_CLEAN", "n2c"), ("NOISE_TO_VOID", "n2v")),
    "NoiseValue": (("UNKNOWN_CONSTANT", "const"), ("UNKNOWN_VARIABLE", "var"), ("KNOWN", "known")),
    "Pipeline": (("MSE", "mse"), ("SSDN", "ssdn"), ("MASK_MSE", "mask_mse")),
    "Blindspot": (("ENABLED", "blindspot"), ("DISABLED", "normal")),
    # the first member is numbered, the rest name the keys of the pipeline output dictionary
    "PipelineOutput": (("INPUTS", 1), ("LOSS", "loss"), ("IMG_DENOISED", "out"), ("IMG_MU", "out_mu"), ("NOISE_STD_DEV", "noise_std"),
                       ("MODEL_STD_DEV", "model_std")),
}
_NUMBERED = {
    "ConfigValue": "INFER_CFG ALGORITHM BLINDSPOT PIPELINE IMAGE_CHANNELS NOISE_STYLE LEARNING_RATE LR_RAMPUP_FRACTION LR_RAMPDOWN_FRACTION "
                   "NOISE_VALUE DIAGONAL_COVARIANCE EVAL_INTERVAL PRINT_INTERVAL SNAPSHOT_INTERVAL TRAIN_ITERATIONS DATALOADER_WORKERS "
                   "TRAIN_DATASET_NAME TRAIN_DATASET_TYPE TRAIN_DATA_PATH TRAIN_PATCH_SIZE TRAIN_MINIBATCH_SIZE TEST_DATASET_NAME "
                   "TEST_DATASET_TYPE TEST_DATA_PATH TEST_MINIBATCH_SIZE PIN_DATA_MEMORY",
    "DatasetType": "HDF5 FOLDER",
    "StateValue": "INITIALISED MODE ITERATION REFERENCE HISTORY",
    "HistoryValue": "TRAIN EVAL TIMINGS",
}


def _build():
    made = {}
    for cls, members in _STRING_VALUED.items():
        made[cls] = Enum(cls, list(members), module=__name__, qualname=cls)
    for cls, names in _NUMBERED.items():
        made[cls] = Enum(cls, names.split(), module=__name__, qualname=cls, start=1)
    return made


globals().update(_build())
__all__ = sorted(list(_STRING_VALUED) + list(_NUMBERED))
