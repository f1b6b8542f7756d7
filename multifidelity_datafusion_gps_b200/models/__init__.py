"""Model presets (NARGP, GPDF, GPDFC as in the reference's ``src.models``) and the multi-level recursion."""
from ._presets import GPDF, GPDFC, NARGP
from .multilevel import MultiLevelNARGP

__all__ = ["NARGP", "GPDF", "GPDFC", "MultiLevelNARGP"]
