"""``models.GPDFC`` module path of the reference (src/models/GPDFC.py); the class lives in ``_presets``."""
from ._presets import GPDFC  # noqa: F401
