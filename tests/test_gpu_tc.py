"""GPU (-m gpu): tcgen05 building blocks.  One bf16x3 UMMA GEMM per operand-layout combination the
fused tensor-core kernels use (K-major / MN-major views of the same shared-memory image, M=128 and
M=64 accumulators), against a float64 matmul."""
import ctypes

import pytest
import torch

from neurallaplace_b200 import _lib

pytestmark = pytest.mark.gpu


def _selftest(A, a_mn, B, b_mn, M, N, K):
  lib = _lib.load()
  fn = lib.nl_debug_umma_selftest
  fn.restype = ctypes.c_int
  fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                 ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                 ctypes.c_void_p]
  D = torch.full((M, N), float("nan"), device=A.device)
  st = fn(A.data_ptr(), A.shape[0], A.shape[1], a_mn, B.data_ptr(), B.shape[0], B.shape[1], b_mn, M, N, K,
          D.data_ptr(), torch.cuda.current_stream().cuda_stream)
  assert st == 0
  torch.cuda.synchronize()
  return D


CASES = [
    # name, A shape, a_mn, B shape, b_mn, M, N, K, rows of D that are meaningful
    ("layer2: H1[128x64] . W2[64x64]^T", (128, 64), 0, (64, 64), 0, 128, 64, 64, 128),
    ("layer3: H2[128x64] . W3c[80x64]^T", (128, 64), 0, (80, 64), 0, 128, 80, 64, 128),
    ("dH2: dO[128x80] . W3c[80x64]", (128, 80), 0, (80, 64), 1, 128, 64, 80, 128),
    ("dW3: dO[128x80]^T . H2[128x64]", (128, 80), 1, (128, 64), 1, 128, 64, 128, 80),
    ("dW2: Z2g[128x64]^T . H1[128x64] (M=64)", (128, 64), 1, (128, 64), 1, 64, 64, 128, 64),
    ("dW1p: Z1g[128x64]^T . Q[128x16] (M=64,N=16)", (128, 64), 1, (128, 16), 1, 64, 16, 128, 64),
    ("layer3 wide: H2[128x64] . W3c[128x64]^T", (128, 64), 0, (128, 64), 0, 128, 128, 64, 128),
]


@pytest.mark.parametrize("name,ashape,a_mn,bshape,b_mn,M,N,K,rows", CASES, ids=[c[0].split(":")[0] for c in CASES])
def test_umma_operand_layouts(name, ashape, a_mn, bshape, b_mn, M, N, K, rows):
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  dev = torch.device("cuda", 0)
  g = torch.Generator().manual_seed(len(name))
  A = torch.randn(ashape, generator=g).to(dev)
  B = torch.randn(bshape, generator=g).to(dev)
  D = _selftest(A, a_mn, B, b_mn, M, N, K)
  Aop = (A.t() if a_mn else A).double()  # [M', K]
  Bop = (B.t() if b_mn else B).double()  # [N', K]
  want = (Aop[:rows, :K] @ Bop[:N, :K].t())
  got = D[:rows].double()
  err = float((got - want).abs().max())
  scale = float(want.abs().max())
  assert err <= 2e-4 * scale, (name, err, scale)
