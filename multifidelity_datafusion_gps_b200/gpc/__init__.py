"""Generalised polynomial chaos on top of the multi-fidelity GP (reference: src/gpc/)."""
from .gpc_abstract import AbstractGPC
from .legendre_pce import LegendrePCE, total_degree_multi_index
from .mfgp_gpc import MFGP_GPC

__all__ = ["AbstractGPC", "LegendrePCE", "MFGP_GPC", "total_degree_multi_index"]
