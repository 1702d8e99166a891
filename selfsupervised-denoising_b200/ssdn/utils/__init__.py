"""Helpers shared by the pipelines: image ops and PSNR (data), bookkeeping (utils), noise models, Noise2Void sampling."""
from . import data_format, n2v_loss, n2v_ups, noise, transforms  # noqa: F401
from .data import *  # noqa: F401,F403
from .utils import *  # noqa: F401,F403
