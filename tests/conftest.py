import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def cuda():
  import torch
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  return torch.device('cuda:0')


def rel_err(a, b):
  """max|a-b| / max(max|b|, tiny): the parity metric of SURVEY.md §8(d)."""
  import numpy as np
  a = np.asarray(a, np.float64)
  b = np.asarray(b, np.float64)
  scale = max(float(np.abs(b).max()) if b.size else 0.0, 1e-12)
  return float(np.abs(a - b).max() / scale) if b.size else 0.0


def oracle_fp64():
  """The model oracle re-typed to float64 (source rewrite: it hard-codes float32 in a few places).  Used as the
  "truth" where fp32 implementations cannot agree with each other to 1e-3 because the computation itself amplifies
  round-off (training-mode forward), and for finite-difference checks of the gradient oracle."""
  import types
  from oracle import model as OM
  src = open(OM.__file__).read()
  src = src.replace('torch.float32', 'torch.float64').replace('.float()', '.double()')
  src = src.replace('from . import hungarian as _hung', 'from oracle import hungarian as _hung')
  mod = types.ModuleType('oracle_model_fp64')
  exec(compile(src, 'oracle_model_fp64', 'exec'), mod.__dict__)
  return mod
