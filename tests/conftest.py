import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
  import numpy as np
  return np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))


@pytest.fixture(scope="session")
def oracle():
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  import oracle as orc
  orc.lib()
  return orc
