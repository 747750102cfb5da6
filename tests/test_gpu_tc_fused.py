"""GPU (-m gpu): the tcgen05 fused kernels (forced with NL_B200_IMPL=tc) against the CPU oracle.
Same bar as the SIMT path: normwise rel <= 1e-4 (float32) on x(t) and on every gradient."""
import pytest
import torch

from tests import helpers as hp
from tests.test_gpu_parity import _run_product

pytestmark = pytest.mark.gpu

CASES = [
    # alg, terms, B, T, D, K, sphere
    ("fourier", 33, 130, 7, 1, 2, True),
    ("fourier", 33, 256, 5, 2, 2, True),
    ("fourier", 33, 77, 3, 3, 3, True),
    ("fourier", 65, 200, 4, 1, 2, True),
    ("fourier", 129, 140, 3, 1, 2, True),
    ("euler", 33, 150, 6, 1, 2, True),
    ("cme", 33, 129, 5, 2, 2, True),
    ("fixed_tablot", 33, 128, 6, 1, 2, True),
    ("fourier", 33, 160, 5, 1, 2, False),
    ("fourier", 17, 300, 9, 1, 7, True),
    ("fourier", 33, 1500, 40, 1, 2, True),
    # d_obs x terms beyond the TMEM / shared-memory resident budget: W3 streamed (bulk-async ring), dW3 scratch flush
    ("fourier", 33, 200, 3, 4, 2, True),
    ("fourier", 33, 300, 5, 16, 2, True),
    ("fourier", 65, 150, 3, 5, 3, True),
    ("fourier", 129, 130, 2, 2, 2, True),
    ("euler", 33, 260, 4, 6, 2, True),
    ("fourier", 129, 2000, 12, 3, 2, True),
    ("fourier", 129, 140, 2, 16, 2, True),
]


# the backward either streams the hidden activations the forward stashed (default) or recomputes them
@pytest.mark.parametrize("stash", ["1", "0"])
@pytest.mark.parametrize("alg,terms,B,T,D,K,sphere", CASES)
def test_tc_against_oracle(alg, terms, B, T, D, K, sphere, stash, monkeypatch):
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  dev = torch.device("cuda", 0)
  monkeypatch.setenv("NL_B200_IMPL", "tc")
  monkeypatch.setenv("NL_B200_STASH", stash)
  from neurallaplace_b200 import _consts, _ops
  from oracle import nl_oracle as orc
  n = orc.IltConsts(alg, terms).n
  c = _consts.get(alg, terms, torch.float32)
  assert _ops.selected_impl(c, torch.float32, B, T, K, 64, D, 0, 0 if sphere else 1, 0) == 2
  assert _ops.selected_impl(c, torch.float32, B, T, K, 64, D, 0, 0 if sphere else 1, 1) == 2
  g = torch.Generator().manual_seed(99 + B + T)
  weights = orc.make_canonical_weights(n, D, K, seed=B + T)
  p = torch.randn(B, K, generator=g)
  t = torch.linspace(20.0 / T, 20.0, T)
  gout = torch.randn(B, T, D, generator=g)
  x_ref, g_ref = hp.oracle_run(weights, p, t, D, alg, sphere, terms, None, gout)
  x, grads = _run_product(weights, p, t, D, alg, sphere, terms, None, gout, dev)
  ex = hp.rel(x, x_ref)
  eg = [hp.rel(a, b) for a, b in zip(grads, g_ref)]
  print(f"\n[tc stash={stash}] {alg} terms={terms} B={B} T={T} D={D}: x {ex:.2e} grads " + " ".join(f"{e:.1e}" for e in eg))
  assert ex <= 1e-4, ex
  for i, e in enumerate(eg):
    assert e <= 1e-4, (i, e)
