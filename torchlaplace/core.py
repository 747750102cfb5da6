"""``torchlaplace.core`` (reference: torchlaplace/core.py:15-187) -> ``neurallaplace_b200.core``."""
from neurallaplace_b200.core import ILT_ALGORITHMS, _check_inputs, laplace_reconstruct

__all__ = ["ILT_ALGORITHMS", "laplace_reconstruct", "_check_inputs"]
