import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
  if p not in sys.path:
    sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
  import torch
  if torch.cuda.is_available():
    return
  skip = pytest.mark.skip(reason='no CUDA device')
  for item in items:
    if 'gpu' in item.keywords:
      item.add_marker(skip)


def load_golden(name):
  import torch
  return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def rel_err(a, b):
  """max|a-b| / max|b| — the relative measure every parity tolerance in this
  repo is stated in."""
  a, b = a.detach().double().cpu(), b.detach().double().cpu()
  return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))
