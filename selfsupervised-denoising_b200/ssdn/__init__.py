"""ssdn - drop-in, B200-native backend for the `ssdn` package of
COMP6248-Reproducability-Challenge/selfsupervised-denoising.

Module paths, class names, constructor signatures, state-dict keys and enum values follow the
reference; the arithmetic of the hot path runs in libssdn_b200.so (hand-written sm_100a kernels)."""
from ssdn.version import __version__  # noqa: F401
from ssdn import params, cfg, utils, models, datasets  # noqa: F401
from ssdn.denoiser import Denoiser  # noqa: F401
from ssdn import train  # noqa: F401
