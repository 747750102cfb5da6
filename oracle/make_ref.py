"""Recipe for ``oracle/_ref``: stage the UNMODIFIED reference package for the CPU baseline.  TEST / BENCH INFRASTRUCTURE.

The reference is pure Python, so there is nothing to compile: when ``/root/reference`` is present (the build
container) this copies the five files of its ``torchlaplace`` package that make up the hot path into
``oracle/_ref/torchlaplace/`` -- ``oracle/_ref/`` is git-ignored (the sources never enter the history) but travels
to the GPU box with the snapshot, like the built ``.so`` -- and writes the one module the checkout lacks,
``_iltcme.py`` (/root/reference/.MISSING_LARGE_BLOBS:1), as a stub returning the oracle's synthetic CME table.
``bench.py --impl reference`` and the ``cpu_baseline`` leg import the package from there
(``cpu_baseline.kind = "reference"``); without it they time the oracle port (``kind = "port"``).
Nothing under ``neurallaplace_b200/`` imports it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/torchlaplace"
DST = os.path.join(HERE, "_ref", "torchlaplace")
FILES = ("__init__.py", "core.py", "inverse_laplace.py", "transformations.py", "version.py")

STUB = '''"""Stub of the module the reference checkout lacks (.MISSING_LARGE_BLOBS): synthetic table, same schema."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from oracle.nl_oracle import synthetic_cme_table as cme_params_factory  # noqa: E402,F401
'''


def main():
  if not os.path.isdir(SRC):
    print("oracle/make_ref.py: /root/reference not present -- keeping whatever oracle/_ref holds")
    return 0
  os.makedirs(DST, exist_ok=True)
  for f in FILES:
    shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
  with open(os.path.join(DST, "_iltcme.py"), "w", encoding="utf-8") as fh:
    fh.write(STUB)
  print(f"oracle/make_ref.py: staged {len(FILES)} reference files + the _iltcme stub in {DST}")
  return 0


if __name__ == "__main__":
  sys.exit(main())
