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


@pytest.fixture(autouse=True, scope='session')
def _exact_fp32_kernels_unless_a_test_says_otherwise():
  """The library's default arithmetic is 'bf16x3' (tensor cores).  The parity tests of the
  non-convolution kernels and of the exact FFMA path (test_gpu_ops.py, test_gpu_model.py, ...) are
  written against 'fp32'; tests of the tensor-core arithmetics select theirs explicitly
  (test_gpu_bf16x3.py: the benchmarked mode, held to the same 1e-3 bar) and restore 'fp32'."""
  from sg2im_b200 import ops
  ops.set_conv_math('fp32')
  yield


def load_golden(name):
  import torch
  return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def rel_err(a, b):
  """max|a-b| / max|b| — the relative measure every parity tolerance in this
  repo is stated in."""
  a, b = a.detach().double().cpu(), b.detach().double().cpu()
  return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))
