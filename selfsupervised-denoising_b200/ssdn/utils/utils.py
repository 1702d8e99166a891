"""Host-side bookkeeping (reference: ssdn/ssdn/utils/utils.py): LR ramp, timers, metric accumulators."""
import math
import time
from collections import OrderedDict

import torch
from torch import Tensor

__all__ = ["compute_ramped_lrate", "TrackedTime", "seconds_to_dhms", "Metric", "MetricDict", "separator"]


def compute_ramped_lrate(i: int, iteration_count: int, ramp_up_fraction: float, ramp_down_fraction: float,
                         learning_rate: float) -> float:
    """Cosine ramp-up over the first fraction of training and squared-cosine ramp-down over the last."""
    if ramp_up_fraction > 0.0 and i <= iteration_count * ramp_up_fraction:
        t = (i / ramp_up_fraction) / iteration_count
        learning_rate *= 0.5 - math.cos(t * math.pi) / 2
    if ramp_down_fraction > 0.0:
        start = iteration_count * (1 - ramp_down_fraction)
        if i >= start:
            t = ((i - start) / ramp_down_fraction) / iteration_count
            learning_rate *= (0.5 + math.cos(t * math.pi) / 2) ** 2
    return learning_rate


def separator(cols: int = 100) -> str:
    return "#" * cols


class TrackedTime:
    """Running total of wall-clock time between successive update() calls."""

    def __init__(self):
        self.total, self.last_time = 0, None

    def update(self):
        now = time.time()
        if self.last_time is not None:
            self.total += now - self.last_time
        self.last_time = now

    def forget(self):
        self.last_time = None


def seconds_to_dhms(seconds: float, trim: bool = True) -> str:
    parts = ((seconds // 86400, "d"), (seconds // 3600 % 24, "h"), (seconds // 60 % 60, "m"), (seconds % 60, "s"))
    out = ""
    for value, unit in parts:
        if trim and value < 1:
            continue
        trim = False
        out += "{:02}{}".format(int(value), unit)
    return out


class Metric:
    """Accumulates batch-summed values (batch on dim 0) and reports their mean."""

    def __init__(self, batched: bool = True, collapse: bool = True):
        self.batched, self.collapse = batched, collapse
        self.reset()

    def reset(self):
        self.total, self.n = None, 0

    def add(self, value: Tensor):
        n = value.shape[0] if self.batched else 1
        if self.collapse:
            dims = list(range(1 if self.batched else 0, value.dim()))
            if dims:
                value = value.mean(dim=dims)
        if self.batched:
            value = value.sum(dim=0)
        # a resumed history arrives on the CPU (map_location) while new values live on the training device
        self.total = value if self.total is None else self.total.to(value.device) + value
        self.n += n

    def __add__(self, value):
        self.add(value)
        return self

    def accumulated(self, reset: bool = False):
        if self.n == 0:
            return None
        acc = self.total / self.n
        if reset:
            self.reset()
        return acc

    def empty(self) -> bool:
        return self.n == 0

    def __str__(self):
        return str(self.accumulated())


class MetricDict(OrderedDict):
    """Dictionary that creates a Metric on first access of a key."""

    def __missing__(self, key):
        self[key] = value = Metric()
        return value
