"""Symmetric delays 0, then for i = 1..n: -i e_1..-i e_dim, +i e_1..+i e_dim
(reference src/augm_iterators/even_augm_iterator.py:20-48; count 2*n*dim+1, :50-51)."""
import numpy as np

from .abstract_augm_iterator import AbstractAugmIterator


class EvenAugmentation(AbstractAugmIterator):
    def __init__(self, n, dim=1):
        super().__init__(n, dim=dim)

    def _build_table(self):
        table = np.zeros((self.new_entries_count(), self.dim))
        row = 1
        for step in range(1, self.n + 1):
            for sign in (-1.0, 1.0):
                for axis in range(self.dim):
                    table[row, axis] = sign * step
                    row += 1
        return table

    def new_entries_count(self):
        return 2 * self.n * self.dim + 1
