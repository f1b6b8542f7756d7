"""Backward delays 0, -e_1..-e_dim, -2e_1.., ..., -n e_dim
(reference src/augm_iterators/backward_augm_iterator.py:20-37; count n*dim+1, :36-37)."""
import numpy as np

from .abstract_augm_iterator import AbstractAugmIterator


class BackwardAugmentation(AbstractAugmIterator):
    def __init__(self, n, dim=1):
        super().__init__(n, dim=dim)

    def _build_table(self):
        table = np.zeros((self.new_entries_count(), self.dim))
        row = 1
        for step in range(1, self.n + 1):
            for axis in range(self.dim):
                table[row, axis] = -step
                row += 1
        return table

    def new_entries_count(self):
        return self.n * self.dim + 1
