"""A small slice of tools/fuzz_tc_emulated.py in the suite: random shapes / tile heuristics /
cluster switches for every tensor-core kernel variant under the functional model."""
import shutil
import subprocess
import sys
import os

import pytest

from conftest import ROOT
from emul_device import HAVE_TC

pytestmark = pytest.mark.skipif(shutil.which('g++') is None or not HAVE_TC,
                                reason='needs g++ (C++20) and the CUDA headers')


def test_fuzz_slice():
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'fuzz_tc_emulated.py'),
                        '14' if os.environ.get('SG2IM_FULL_EMUL') == '1' else '7', '7'],
                       capture_output=True, text=True, timeout=900)
  assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
  assert 'all cases agree' in out.stdout
