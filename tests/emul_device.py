"""TEST INFRASTRUCTURE ONLY: run the PRODUCT's own op layer (sg2im_b200/ops.py — the ctypes
marshalling, stride / slice arithmetic and autograd Functions exactly as shipped) on CPU tensors
against the kernel sources compiled for the host (tests/emul/, -DSG2IM_EMUL).

``with emulated_device(): ...`` swaps the loaded library handle for the emulation build and
relaxes the two device checks of the op layer (`_chk`'s is_cuda test, the CUDA stream lookup);
nothing else of the product is touched.  When the CUDA headers are present (HAVE_TC) the
tensor-core kernels are part of the host build too — they execute against the functional
TMA / mbarrier / tcgen05 / TMEM / cluster model of tests/emul/tc_emul.h, so `set_conv_math('tf32')`
(the benchmarked configuration) runs end to end; without them a convolution that would take the
tensor-core path fails loudly (missing symbol).
"""
import contextlib
import ctypes
import os
import subprocess
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = next((d for d in (os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'include'),
                             '/usr/local/cuda/include') if os.path.exists(os.path.join(d, 'cuda.h'))), None)
HAVE_TC = CUDA_INC is not None
_built = {}


def build_lib():
  """Compile tests/emul/emul_*.cpp (= the .cu sources with -DSG2IM_EMUL) once per process."""
  if 'path' in _built:
    return _built['path']
  import glob
  out = os.path.join(tempfile.mkdtemp(prefix='sg2im_emul_'), 'libemul.so')
  src = sorted(glob.glob(os.path.join(ROOT, 'tests', 'emul', 'emul_*.cpp')))
  cmd = ['g++', '-std=c++20', '-O1', '-pthread', '-shared', '-fPIC', '-Wno-psabi', '-U_FORTIFY_SOURCE', '-DSG2IM_EMUL',
         '-I', os.path.join(ROOT, 'tests', 'emul'), '-I', os.path.join(ROOT, 'include'),
         '-I', os.path.join(ROOT, 'sg2im_b200', 'csrc')]
  if CUDA_INC is not None and '-DSG2IM_EMUL_THREADS' not in os.environ.get('SG2IM_EMUL_CXXFLAGS', ''):
    # with the CUDA headers (cuda.h: CUtensorMap) the tensor-core kernels build too, against the
    # functional tcgen05 / TMA / cluster model of tests/emul/tc_emul.h (fiber execution model only:
    # the OS-thread model used for the ASan / TSan runs builds the SIMT kernels alone)
    src += sorted(glob.glob(os.path.join(ROOT, 'tests', 'emul', 'emultc_*.cpp')))
    cmd += ['-I', CUDA_INC]
  cmd += src + ['-o', out]
  cmd[1:1] = os.environ.get('SG2IM_EMUL_CXXFLAGS', '').split()
  subprocess.check_call(cmd)
  _built['path'] = out
  return out


@contextlib.contextmanager
def emulated_device():
  from sg2im_b200 import _lib, ops
  lib = ctypes.CDLL(build_lib())
  for name, sig in _lib.SIGNATURES.items():
    if hasattr(lib, name):
      fn = getattr(lib, name)
      fn.argtypes = sig
      fn.restype = ctypes.c_int
  lib.sg2im_last_error_string.argtypes = []
  lib.sg2im_last_error_string.restype = ctypes.c_char_p

  def chk(t, dtype=torch.float32, name='tensor'):
    if t.dtype != dtype:
      raise RuntimeError('sg2im_b200: %s must be %s, got %s' % (name, dtype, t.dtype))
    return t

  saved = (_lib._lib, ops._chk, ops._stream, ops.CONV_MATH)
  _lib._lib, ops._chk, ops._stream, ops.CONV_MATH = lib, chk, (lambda: None), 'fp32'
  try:
    yield lib
  finally:
    _lib._lib, ops._chk, ops._stream, ops.CONV_MATH = saved
