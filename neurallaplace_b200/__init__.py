"""B200-native reconstruction hot path of ``torchlaplace`` (samholt/NeuralLaplace).

Drop-in for the reference's public API on this path:

    from neurallaplace_b200 import laplace_reconstruct            # torchlaplace.laplace_reconstruct
    from neurallaplace_b200.inverse_laplace import Fourier, ...     # torchlaplace.inverse_laplace.*
    from neurallaplace_b200.transformations import ...              # torchlaplace.transformations.*

Hand-written sm_100a CUDA behind a C ABI (include/nl_b200.h); PyTorch provides device memory,
streams, autograd bookkeeping and torch.distributed only.  No CPU fallback.
"""
from .core import ILT_ALGORITHMS, laplace_reconstruct
from .version import __version__

__all__ = ["laplace_reconstruct", "ILT_ALGORITHMS", "__version__"]
