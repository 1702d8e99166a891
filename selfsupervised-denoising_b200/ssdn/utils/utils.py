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
    """'01d02h03m04s'; with trim, leading units that are zero are left out."""
    minutes, secs = divmod(seconds, 60)
    hours, minutes = divmod(minutes, 60)
    days, hours = divmod(hours, 24)
    fields = [(days, "d"), (hours, "h"), (minutes, "m"), (secs, "s")]
    if trim:
        while fields and fields[0][0] < 1:
            fields.pop(0)
    return "".join("{:02}{}".format(int(v), unit) for v, unit in fields)


class Metric:
    """Accumulates batch-summed values (batch on dim 0) and reports their mean."""

    def __init__(self, batched: bool = True, collapse: bool = True):
        self.batched, self.collapse = batched, collapse
        self.reset()

    def reset(self):
        self.total, self.n = None, 0

    def add(self, value: Tensor):
        count = 1
        lead = 0
        if self.batched:
            count, lead = value.shape[0], 1
        if self.collapse and value.dim() > lead:
            value = value.mean(dim=tuple(range(lead, value.dim())))
        if self.batched:
            value = value.sum(dim=0)
        # a resumed history arrives on the CPU (map_location) while new values live on the training device
        self.total = value if self.total is None else self.total.to(value.device) + value
        self.n += count

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
