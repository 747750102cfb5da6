"""``torchlaplace.transformations`` (reference: torchlaplace/transformations.py:28-151) ->
``neurallaplace_b200.transformations`` (CUDA kernels ``nl_complex_to_sphere`` / ``nl_sphere_to_complex``)."""
from neurallaplace_b200.transformations import (complex_to_spherical, complex_to_spherical_riemann,
                                                spherical_riemann_to_complex, spherical_to_complex)

__all__ = [
    "spherical_riemann_to_complex", "complex_to_spherical_riemann", "spherical_to_complex", "complex_to_spherical"
]
