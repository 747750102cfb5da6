"""``torchlaplace.inverse_laplace`` (reference: torchlaplace/inverse_laplace.py:49-1689) ->
``neurallaplace_b200.inverse_laplace``: same class names (reference spellings, e.g. ``FixedTablot``), constructor
arguments and methods; ``compute_s`` / ``line_integrate*`` / ``forward`` run as CUDA kernels."""
import sys

import torch

from neurallaplace_b200.inverse_laplace import (CME, TORCH_COMPLEX_DATATYPE, TORCH_FLOAT_DATATYPE, DeHoog, Euler,
                                                FixedTablot, Fourier, Gaver, InverseLaplaceTransformAlgorithmBase,
                                                Stehfest, real_vector_to_complex)

EPS = sys.float_info.epsilon  # inverse_laplace.py:49
# The reference keeps its constants on this module-global device (l.53).  Here every call follows the device of
# its tensors instead; the name is kept for callers that read it.
device = "cuda" if torch.cuda.is_available() else "cpu"


def complex_numpy_to_complex_torch(sn):
  """numpy complex array -> flat complex64 tensor through float32 parts (inverse_laplace.py:77-85)."""
  import numpy as np
  sn = np.asarray(sn)
  re = torch.tensor(np.real(sn), dtype=torch.float32).flatten()
  im = torch.tensor(np.imag(sn), dtype=torch.float32).flatten()
  return torch.complex(re, im)


__all__ = [
    "InverseLaplaceTransformAlgorithmBase", "FixedTablot", "Stehfest", "DeHoog", "Fourier", "Gaver", "Euler", "CME",
    "real_vector_to_complex", "complex_numpy_to_complex_torch", "TORCH_FLOAT_DATATYPE", "TORCH_COMPLEX_DATATYPE",
    "EPS", "device"
]
