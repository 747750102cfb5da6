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


DEHOOG_CASES = [
    # terms, B, T, D, K, sphere
    (33, 128, 8, 1, 2, True),
    (33, 192, 6, 2, 3, True),
    (17, 128, 9, 1, 2, True),
    (65, 128, 4, 1, 2, True),
    (33, 160, 5, 1, 7, False),
]


def _dehoog_protocol(weights, p, t, D, sphere, terms, gout, dev, G, what):
  """The robust float32 De Hoog comparison of tests/helpers.py (SURVEY section 8d's 4x factor on outlier-proof
  statistics): returns nothing, asserts."""
  B = p.shape[0]
  x32, g32 = hp.oracle_group_grads(weights, p, t, D, "dehoog", sphere, terms, gout, G)
  x64, g64 = hp.oracle_group_grads([w.double() for w in weights], p.double(), t.double(), D, "dehoog", sphere, terms,
                                   gout.double(), G)
  groups, x = [], None
  for gi in range(G):
    x, grads = _run_product(weights, p, t, D, "dehoog", sphere, terms, None, gout * hp.group_mask(B, G, gi), dev)
    groups.append([q.detach().cpu() for q in grads])
  rm, r90 = hp.pointwise_ratios(x, x32, x64)
  print(f"\n[{what}] terms={terms} B={B} T={t.numel()} D={D}: pointwise |x - x64| vs reference-f32: median x{rm:.2f} "
        f"p90 x{r90:.2f}   (normwise: ours {hp.rel(x, x64):.1e} reference-f32 {hp.rel(x32, x64):.1e})")
  assert rm <= hp.DEHOOG_FACTOR and r90 <= hp.DEHOOG_FACTOR, (rm, r90)
  for i, (em, er, ratio) in enumerate(hp.group_median_ratios(groups, g32, g64)):
    print(f"    grad{i}: median-over-groups normwise error ours {em:.2e} reference-f32 {er:.2e}  x{ratio:.2f}")
    assert em <= hp.DEHOOG_FACTOR * er + 1e-4, (i, em, er)


@pytest.mark.parametrize("terms,B,T,D,K,sphere", DEHOOG_CASES)
def test_tc_dehoog_against_oracle(terms, B, T, D, K, sphere, monkeypatch):
  """De Hoog with the MLP on the tensor cores (default float32 path from 16 batch rows): the forward evaluates
  layers 2-3 with a THREE-term operand split (hi|mid|lo, six bf16 passes: F to float32 accuracy), writes F planes,
  the per-thread quotient-difference recurrence turns them into x; the saved backward reads dL/dF.
  Bar: SURVEY section 8d -- error against the float64 oracle no more than 4x the reference-float32 error against
  the float64 oracle -- on the outlier-proof statistics of tests/helpers.py (x pointwise median / p90; gradients
  median over 8 row groups), the same protocol for the SIMT kernels (float32 FMA)."""
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  dev = torch.device("cuda", 0)
  from neurallaplace_b200 import _consts, _ops
  from oracle import nl_oracle as orc
  n = orc.IltConsts("dehoog", terms).n
  c = _consts.get("dehoog", terms, torch.float32)
  # auto mode picks the tensor-core kernels here, forward and backward, with an activation stash
  assert _ops.selected_impl(c, torch.float32, B, T, K, 64, D, 0, 0 if sphere else 1, 0) == 2
  assert _ops.selected_impl(c, torch.float32, B, T, K, 64, D, 0, 0 if sphere else 1, 1) == 2
  assert _ops.saved_bytes(c, torch.float32, B, T, K, 64, D, 0, 0 if sphere else 1) > 0
  g = torch.Generator().manual_seed(5 + B + T)
  weights = orc.make_canonical_weights(n, D, K, seed=B + T, sphere=sphere)
  p = torch.randn(B, K, generator=g)
  t = torch.linspace(20.0 / T, 20.0, T)
  gout = torch.randn(B, T, D, generator=g)
  _dehoog_protocol(weights, p, t, D, sphere, terms, gout, dev, 8, "tc dehoog")
  monkeypatch.setenv("NL_B200_IMPL", "simt")
  _dehoog_protocol(weights, p, t, D, sphere, terms, gout, dev, 8, "simt dehoog")


def test_tc_dehoog_without_stash_falls_back_for_backward(monkeypatch):
  """NL_B200_STASH=0: the forward still runs on the tensor cores (F plane in the workspace), the backward has no
  saved F / activations and runs on the SIMT kernels + the recurrence kernels; same bar as above."""
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  dev = torch.device("cuda", 0)
  monkeypatch.setenv("NL_B200_STASH", "0")
  from oracle import nl_oracle as orc
  terms, B, T, D, K = 33, 96, 6, 1, 2
  n = orc.IltConsts("dehoog", terms).n
  g = torch.Generator().manual_seed(11)
  weights = orc.make_canonical_weights(n, D, K, seed=7)
  p = torch.randn(B, K, generator=g)
  t = torch.linspace(0.5, 10.0, T)
  gout = torch.randn(B, T, D, generator=g)
  _dehoog_protocol(weights, p, t, D, True, terms, gout, dev, 4, "tc fwd / simt bwd dehoog")
