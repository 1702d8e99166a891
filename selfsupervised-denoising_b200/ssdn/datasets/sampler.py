"""Fixed-length, resumable sampling (reference: ssdn/ssdn/datasets/sampler.py).

A training run consumes a fixed number of samples; when that exceeds the dataset length the dataset
is cycled so that no sample is used more than once more than any other.  The drawn order is an
explicit list with a cursor, so it can be checkpointed and resumed mid-epoch."""
from __future__ import annotations

from typing import Dict, Iterator, List

import torch
from torch.utils.data import Dataset, Sampler


class SamplingOrder:
    """A materialised index order plus the position of the next index to hand out."""

    def __init__(self, order: List[int], index: int = 0):
        self.order, self.index = order, index

    def __iter__(self) -> "SamplingOrder":
        return self

    def __len__(self) -> int:
        return len(self.order)

    def __next__(self) -> int:
        if self.index >= len(self.order):
            raise StopIteration()
        self.index += 1
        return self.order[self.index - 1]

    def state_dict(self) -> Dict:
        return {"order": self.order, "index": self.index}

    @staticmethod
    def from_state_dict(state_dict: Dict) -> "SamplingOrder":
        return SamplingOrder(state_dict["order"], state_dict["index"])


class FixedLengthSampler(Sampler):
    def __init__(self, data_source: Dataset, num_samples: int = None, shuffled: bool = False):
        self.data_source, self._num_samples, self.shuffled = data_source, num_samples, shuffled
        self._next_iter = None
        self._last_iter = None

    @property
    def num_samples(self) -> int:
        return len(self.data_source) if self._num_samples is None else self._num_samples

    def sampler(self) -> Iterator[int]:
        size, left = len(self.data_source), self.num_samples
        if self.shuffled:
            while left > 0:                      # one fresh permutation per pass over the data
                take = min(left, size)
                yield from (int(i) for i in torch.randperm(size)[:take])
                left -= take
        else:
            for k in range(left):
                yield k % size

    def __iter__(self):
        if self._next_iter is not None:
            return self._next_iter
        self._last_iter = SamplingOrder(list(self.sampler()))
        return self._last_iter

    def __len__(self) -> int:
        return self.num_samples

    def for_next_iter(self, iter_order: SamplingOrder):
        """Inject a (restored) order to be used by the next iteration instead of drawing a new one."""
        self._next_iter = self._last_iter = iter_order

    def last_iter(self) -> SamplingOrder:
        return self._last_iter
