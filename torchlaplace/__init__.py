"""Drop-in import path of the reference package (``/root/reference/torchlaplace/__init__.py:1-2``).

``import torchlaplace`` / ``from torchlaplace import laplace_reconstruct`` resolve to the B200-native
implementation in ``neurallaplace_b200`` (hand-written sm_100a CUDA behind the C ABI of ``include/nl_b200.h``),
so the reference's examples, tests and experiment code import unmodified:

    torchlaplace.laplace_reconstruct                    -> neurallaplace_b200.core.laplace_reconstruct
    torchlaplace.core.{laplace_reconstruct, ILT_ALGORITHMS}
    torchlaplace.inverse_laplace.{InverseLaplaceTransformAlgorithmBase, Fourier, DeHoog, CME, FixedTablot,
                                   Stehfest, Euler, Gaver, real_vector_to_complex, ...}
    torchlaplace.transformations.{complex_to_spherical_riemann, spherical_riemann_to_complex,
                                   spherical_to_complex, complex_to_spherical}
    torchlaplace.data_utils.basic_collate_fn             (host-side batching helper the examples import)
"""
from neurallaplace_b200.core import laplace_reconstruct
from neurallaplace_b200.version import __version__

__all__ = ["laplace_reconstruct", "__version__"]
