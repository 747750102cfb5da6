"""CPU: the ``torchlaplace`` drop-in import path: every name the reference exports for the hot
path resolves to the B200 implementation, ``basic_collate_fn`` reproduces the reference's batches (fixture
generated from the live reference by tests/golden/make_golden.py), and the CME table loader verifies the reference
file by git blob hash."""
import inspect
import os
import subprocess

import numpy as np
import pytest
import torch

from tests import helpers as hp


def test_every_boundary_name_resolves_to_the_b200_implementation():
  import neurallaplace_b200 as nb
  import torchlaplace
  import torchlaplace.core
  import torchlaplace.data_utils
  import torchlaplace.inverse_laplace as il
  import torchlaplace.transformations as tr
  assert torchlaplace.__file__.startswith(hp.ROOT)
  assert torchlaplace.laplace_reconstruct is nb.laplace_reconstruct
  assert torchlaplace.core.laplace_reconstruct is nb.laplace_reconstruct
  assert set(torchlaplace.core.ILT_ALGORITHMS) == {"fourier", "dehoog", "cme", "euler", "fixed_tablot", "stehfest"}
  assert isinstance(torchlaplace.__version__, str)
  import torchlaplace.version  # (reference: torchlaplace/__init__.py:2)
  assert torchlaplace.version.__version__ == torchlaplace.__version__
  # signature of the reference (core.py:25-36)
  sig = inspect.signature(torchlaplace.laplace_reconstruct)
  assert list(sig.parameters) == ["laplace_rep_func", "p", "t", "recon_dim", "ilt_algorithm", "use_sphere_projection",
                                  "ilt_reconstruction_terms", "options", "compute_deriv", "x0"]
  assert sig.parameters["ilt_algorithm"].default == "fourier" and sig.parameters["ilt_reconstruction_terms"].default == 33
  for name in ("InverseLaplaceTransformAlgorithmBase", "Fourier", "DeHoog", "CME", "FixedTablot", "Stehfest", "Euler",
               "Gaver"):
    cls = getattr(il, name)
    assert issubclass(cls, torch.nn.Module) and issubclass(cls, il.InverseLaplaceTransformAlgorithmBase)
  for cls in (il.Fourier, il.DeHoog, il.CME, il.FixedTablot, il.Stehfest, il.Euler, il.Gaver):
    for meth in ("compute_s", "line_integrate", "line_integrate_all_multi", "forward"):
      assert callable(getattr(cls, meth)), (cls.__name__, meth)
    params = inspect.signature(cls.__init__).parameters
    assert list(params)[1] == "ilt_reconstruction_terms" and params["ilt_reconstruction_terms"].default == 33
    assert "torch_float_datatype" in params and "torch_complex_datatype" in params
  for meth in ("compute_fixed_s", "fixed_line_integrate"):
    assert callable(getattr(il.DeHoog, meth))
  assert "time_max" in inspect.signature(il.FixedTablot.line_integrate).parameters
  assert "time_max" in inspect.signature(il.Fourier.compute_s).parameters
  for cls in (il.Fourier, il.FixedTablot):
    assert callable(cls.line_integrate_multi)
  for cls in (il.Fourier, il.CME, il.Euler, il.Gaver):
    assert callable(cls.line_integrate_all_multi_batch_time)
  assert il.TORCH_FLOAT_DATATYPE == torch.float32 and il.TORCH_COMPLEX_DATATYPE == torch.cfloat
  z = il.real_vector_to_complex(torch.tensor([1.0, 2.0]))
  assert z.is_complex() and torch.equal(z.imag, torch.zeros(2))
  c = il.complex_numpy_to_complex_torch(np.array([1.0 + 2.0j, 3.0 - 1.0j]))
  assert c.dtype == torch.cfloat and c.shape == (2,)
  for name in ("complex_to_spherical_riemann", "spherical_riemann_to_complex", "spherical_to_complex",
               "complex_to_spherical"):
    assert callable(getattr(tr, name))
  assert callable(torchlaplace.data_utils.basic_collate_fn)


def test_error_contract_through_the_shim():
  """core.py:163-180: the validation errors come before any CUDA work, so they are checkable on a CPU box."""
  import torchlaplace
  f = torch.nn.Linear(2, 2)
  p, t = torch.zeros(2, 2), torch.ones(3)
  with pytest.raises(RuntimeError):
    torchlaplace.laplace_reconstruct(lambda x: x, p, t)
  with pytest.raises(RuntimeError):
    torchlaplace.laplace_reconstruct(f, [1.0], t)
  with pytest.raises(ValueError, match="Invalid ILT algorithm"):
    torchlaplace.laplace_reconstruct(f, p, t, ilt_algorithm="talbot")
  with pytest.raises(ValueError, match="odd"):
    torchlaplace.laplace_reconstruct(f, p, t, ilt_reconstruction_terms=32)
  if not torch.cuda.is_available():
    # valid arguments, no GPU: the product path fails loudly instead of falling back
    with pytest.raises(RuntimeError, match="no CPU fallback"):
      torchlaplace.laplace_reconstruct(f, p, t)


def test_basic_collate_fn_matches_the_reference(golden_dir):
  from torchlaplace.data_utils import basic_collate_fn
  z = np.load(os.path.join(golden_dir, "collate.npz"))
  trajs = [torch.from_numpy(a) for a in z["trajs"]]
  ts = torch.from_numpy(z["ts"])
  cases = dict(interp=dict(), extrap=dict(extrap=True), extrap_steps=dict(extrap=True, observe_step=2, predict_step=3),
               hopper=dict(extrap=True, data="hopper", data_type="test"))
  for name, kw in cases.items():
    d = basic_collate_fn(list(trajs), ts, **kw)
    assert ",".join(d.keys()) == str(z[f"{name}_keys"])
    assert d["mode"] == str(z[f"{name}_mode"])
    assert d["mask_predicted_data"] is None and d["labels"] is None
    for k in ("observed_data", "observed_tp", "data_to_predict", "tp_to_predict", "observed_mask"):
      assert np.array_equal(d[k].numpy(), z[f"{name}_{k}"]), (name, k)
  # the outputs are copies (the reference clones): writing to them must not touch the inputs
  d = basic_collate_fn(list(trajs), ts)
  d["observed_data"].zero_()
  assert float(trajs[0].abs().sum()) > 0


def test_cme_table_is_verified_by_git_blob_hash(tmp_path):
  from neurallaplace_b200 import _iltcme
  path = tmp_path / "_iltcme.py"
  path.write_text("def cme_params_factory():\n"
                  "  return [dict(n=1, cv2=1.0, c=0.1, a=[0.1], b=[0.2], omega=0.3, mu1=1.0)]\n")
  sha = _iltcme.git_blob_sha1(str(path))
  try:
    want = subprocess.check_output(["git", "hash-object", str(path)]).decode().strip()
    assert sha == want
  except (OSError, subprocess.CalledProcessError):
    pass
  assert len(_iltcme.REFERENCE_BLOB_SHA1) == 40 and sha != _iltcme.REFERENCE_BLOB_SHA1
  with pytest.raises(ValueError, match="git blob"):
    _iltcme.load_table(str(path))  # not the reference's file: refused before it is executed
  v = _iltcme.table_version()
  _iltcme.load_table(str(path), verify=False)
  assert _iltcme.table_origin() == "unverified:" + sha and _iltcme.table_version() == v + 1
  from oracle import nl_oracle as orc
  _iltcme.set_table(orc.synthetic_cme_table())  # restore the session table
  assert _iltcme.table_origin() == "user"
