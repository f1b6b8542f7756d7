"""Delay patterns of the augmentation step: which multiples of tau each input is shifted by before the
low-fidelity function is evaluated (reference: src/augm_iterators/)."""
from .abstract_augm_iterator import AbstractAugmIterator
from .backward_augm_iterator import BackwardAugmentation
from .even_augm_iterator import EvenAugmentation

__all__ = ["AbstractAugmIterator", "BackwardAugmentation", "EvenAugmentation"]
