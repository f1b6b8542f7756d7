"""Iterator protocol of the delay patterns (reference src/augm_iterators/abstract_augm_iterator.py:13-35).

Besides the reference's ``__iter__ / __next__ / new_entries_count / reset`` protocol every iterator
offers ``offset_table()``: the whole (E, dim) integer offset table at once, which is what the GPU
augmentation kernel consumes (the reference walks the iterator once per input row)."""
from abc import ABCMeta, abstractmethod

import numpy as np


class AbstractAugmIterator(metaclass=ABCMeta):
    @abstractmethod
    def __init__(self, n, dim=1):
        self.n = n
        self.dim = dim
        self.reset()

    def __iter__(self):
        return self

    def __next__(self):
        if self._pos >= len(self._table):
            self.reset()
            raise StopIteration
        v = self._table[self._pos].copy()
        self._pos += 1
        return v

    @abstractmethod
    def new_entries_count(self):
        """number of augmented entries this pattern adds per input row"""

    @abstractmethod
    def _build_table(self):
        """(E, dim) float array of offsets in units of tau"""

    def reset(self):
        """(re)initialise the iterator state so that the same object is reusable"""
        self._table = self._build_table()
        self._pos = 0

    def offset_table(self):
        return self._build_table()
