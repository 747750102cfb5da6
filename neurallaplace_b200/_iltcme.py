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

Integrity: the reference's file is git blob ``bef335ebe59dfe7de17f34417a9eed27fd0c3091`` (SURVEY section 0
finding 1).  ``load_table`` of a ``.py`` file computes the file's git blob hash and, unless it matches,
raises (``verify=True``, the default for ``NL_B200_CME_TABLE``) or records the mismatch
(``verify=False``); ``table_origin()`` tells which table is installed: ``"reference"`` (hash verified),
``"unverified:<sha>"`` or ``"user"`` (``set_table`` / JSON).
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os

REFERENCE_BLOB_SHA1 = "bef335ebe59dfe7de17f34417a9eed27fd0c3091"

_table = None
_version = 0
_origin = None


def git_blob_sha1(path):
  """SHA-1 git assigns to the file's contents (``git hash-object``): sha1(b"blob <size>\0" + data)."""
  with open(path, "rb") as f:
    data = f.read()
  return hashlib.sha1(b"blob %d\0" % len(data) + data).hexdigest()


def table_origin():
  return _origin


def set_table(table):
  """Install a CME table (list of dicts with keys n, cv2, c, a, b, omega, mu1)."""
  global _table, _version
  for p in table:
    for key in ("n", "cv2", "c", "a", "b", "omega", "mu1"):
      if key not in p:
        raise ValueError(f"CME table entry lacks '{key}'")
    if len(p["a"]) != p["n"] or len(p["b"]) != p["n"]:
      raise ValueError("CME table entry: len(a) and len(b) must equal n")
  global _origin
  _table = list(table)
  _version += 1
  _origin = "user"


def load_table(path, verify=True):
  """Install the table from ``path``.  A ``.py`` file is taken to be the reference's ``_iltcme.py`` and checked
  against its git blob hash BEFORE it is executed: with ``verify=True`` a mismatch raises, with ``verify=False``
  it is recorded in ``table_origin()``.  A ``.json`` file is a user table (no hash to check against)."""
  global _origin
  if path.endswith(".json"):
    with open(path, "r", encoding="utf-8") as f:
      set_table(json.load(f))
    return
  sha = git_blob_sha1(path)
  if sha != REFERENCE_BLOB_SHA1 and verify:
    raise ValueError(f"{path}: git blob {sha} is not the reference's torchlaplace/_iltcme.py "
                     f"({REFERENCE_BLOB_SHA1}); pass verify=False to install it as an unverified table")
  spec = importlib.util.spec_from_file_location("_nl_user_iltcme", path)
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  set_table(mod.cme_params_factory())
  _origin = "reference" if sha == REFERENCE_BLOB_SHA1 else "unverified:" + sha


def table_version():
  return _version


def cme_params_factory():
  if _table is None:
    path = os.environ.get("NL_B200_CME_TABLE")
    if path:
      load_table(path, verify=os.environ.get("NL_B200_CME_VERIFY", "1") != "0")
  if _table is None:
    raise RuntimeError(
        "CME parameter table not available: the reference's torchlaplace/_iltcme.py is not redistributed "
        "here. Call neurallaplace_b200._iltcme.set_table(...) / load_table(path) or set NL_B200_CME_TABLE.")
  return _table
