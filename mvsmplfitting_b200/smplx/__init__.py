"""Drop-in for the reference's `smplx` package as far as the fitting path uses it
(code/smplx/__init__.py exports create_scale; init.py:101 calls it)."""
from .body_models_scale import SMPL, ModelOutput, create_scale, Struct  # noqa: F401
