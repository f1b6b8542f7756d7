"""``models.NARGP`` module path of the reference (src/models/NARGP.py); the class lives in ``_presets``."""
from ._presets import NARGP  # noqa: F401
