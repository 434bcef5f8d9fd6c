import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
  with np.load(os.path.join(GOLDEN, name)) as f:
    return {k: f[k] for k in f.files}


@pytest.fixture(scope="session")
def golden_models():
  return load_golden("golden_models.npz")


@pytest.fixture(scope="session")
def golden_math():
  return load_golden("golden_math.npz")


@pytest.fixture(scope="session")
def golden_like():
  return load_golden("golden_like.npz")


@pytest.fixture(scope="session")
def golden_like2():
  return load_golden("golden_like2.npz")


@pytest.fixture(scope="session")
def golden_in1d():
  return load_golden("golden_inputs_1d.npz")


@pytest.fixture(scope="session")
def golden_inpix():
  return load_golden("golden_inputs_pix.npz")


@pytest.fixture(scope="session")
def golden_setup():
  return load_golden("golden_setup.npz")
