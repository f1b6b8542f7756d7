"""``models.GPDF`` module path of the reference (src/models/GPDF.py); the class lives in ``_presets``."""
from ._presets import GPDF  # noqa: F401
