import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
  return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session", autouse=True)
def _cme_table():
  # CME table values are not redistributable/recoverable (see neurallaplace_b200/_iltcme.py):
  # every test uses the oracle's synthetic table of the same schema, in product and oracle alike.
  from neurallaplace_b200 import _iltcme
  from oracle import nl_oracle as orc
  _iltcme.set_table(orc.synthetic_cme_table())
  yield
