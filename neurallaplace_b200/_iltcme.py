"""CME parameter table plumbing (stands in for the reference's ``torchlaplace/_iltcme.py``).

The reference ships the numerically optimised CME parameters of Horvath et al. (2020) as a large
Python data module, ``cme_params_factory() -> [ {n, cv2, c, a[n], b[n], omega, mu1}, ... ]``
(schema: /root/reference/torchlaplace/inverse_laplace.py:1622-1634).  That file is ABSENT from
the reference checkout this repository was built against (.MISSING_LARGE_BLOBS:1), so the table
VALUES cannot be reproduced here ("parity unpinned" for CME table values).  The arithmetic that
consumes the table is implemented and parity-tested with tables of the same schema.

Provide the table in one of three ways:
  * ``set_table(list_of_dicts)``;
  * ``load_table(path)`` with a ``.json`` file (list of dicts) or the original ``_iltcme.py``;
  * environment variable ``NL_B200_CME_TABLE=<path>`` (read lazily on first use).
"""
from __future__ import annotations

import importlib.util
import json
import os

_table = None
_version = 0


def set_table(table):
  """Install a CME table (list of dicts with keys n, cv2, c, a, b, omega, mu1)."""
  global _table, _version
  for p in table:
    for key in ("n", "cv2", "c", "a", "b", "omega", "mu1"):
      if key not in p:
        raise ValueError(f"CME table entry lacks '{key}'")
    if len(p["a"]) != p["n"] or len(p["b"]) != p["n"]:
      raise ValueError("CME table entry: len(a) and len(b) must equal n")
  _table = list(table)
  _version += 1


def load_table(path):
  if path.endswith(".json"):
    with open(path, "r", encoding="utf-8") as f:
      set_table(json.load(f))
    return
  spec = importlib.util.spec_from_file_location("_nl_user_iltcme", path)
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  set_table(mod.cme_params_factory())


def table_version():
  return _version


def cme_params_factory():
  if _table is None:
    path = os.environ.get("NL_B200_CME_TABLE")
    if path:
      load_table(path)
  if _table is None:
    raise RuntimeError(
        "CME parameter table not available: the reference's torchlaplace/_iltcme.py is not redistributed "
        "here. Call neurallaplace_b200._iltcme.set_table(...) / load_table(path) or set NL_B200_CME_TABLE.")
  return _table
