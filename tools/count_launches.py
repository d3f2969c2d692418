"""Launch census of one training iteration at the benchmark size, WITHOUT a GPU: the shipped op
layer and module mirror run on CPU tensors with every kernel entry point stubbed to "return 0"
(only the *_supported queries are answered by the real host code, from the emulation build), so
the control flow, allocations and the library-call sequence are exactly those of a real step
while no arithmetic happens.  Prints library calls per iteration by entry point for the weight /
optimiser modes.  (Tool for DESIGN.md numbers; not part of the library.)"""
import collections
import contextlib
import ctypes
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402
import bench  # noqa: E402
from emul_device import build_lib  # noqa: E402
from sg2im_b200 import _lib, ops  # noqa: E402
from sg2im_b200.discriminators import AcCropDiscriminator, PatchDiscriminator  # noqa: E402
from sg2im_b200.model import Sg2ImModel  # noqa: E402
from sg2im_b200.synth import CONFIGS, make_vocab, synth_batch  # noqa: E402
from sg2im_b200.train_step import TrainStep  # noqa: E402


class NullLib(object):
  def __init__(self, real):
    self.real, self.calls = real, collections.Counter()

  def __getattr__(self, name):
    if name.endswith('_supported') or name in ('sg2im_last_error_string', 'sg2im_abi_version', 'sg2im_device_ok'):
      return getattr(self.real, name)
    def stub(*a):
      self.calls[name] += 1
      return 0
    return stub


def main():
  cfg = dict(CONFIGS[os.environ.get('WORKLOAD', 'vg128')])
  cfg['N'] = int(os.environ.get('N', 4))                 # launch counts do not depend on the batch size
  real = ctypes.CDLL(build_lib())
  for name, sig in _lib.SIGNATURES.items():
    if hasattr(real, name):
      getattr(real, name).argtypes = sig
  real.sg2im_last_error_string.restype = ctypes.c_char_p
  ops._chk = lambda t, dtype=torch.float32, name='tensor': t
  ops._stream = lambda: None
  import sg2im_b200.train_step as ts
  ts._all_finite = lambda value, group=None: True        # outputs are garbage: never skip the backward
  vocab = make_vocab(cfg['num_objs'], cfg['num_preds'])
  for label, kw, env in (('default (packed weights, torch Adam)', {}, {}),
                         ('pack-both', {}, {'PACK_BOTH': True}),
                         ('flat Adam', {'fused_adam': 'flat'}, {}),
                         ('kcc weights, torch Adam', {'weights': 'kcc'}, {}),
                         ('kcc weights, flat Adam', {'weights': 'kcc', 'fused_adam': 'flat'}, {}),
                         ("error-compensated ('tf32x3')", {}, {'MATH': 'tf32x3'})):
    null = NullLib(real)
    _lib._lib = null
    ops.PACK_BOTH = bool(env.get('PACK_BOTH'))
    ops.set_conv_math(env.get('MATH', 'tf32'))
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
      model = Sg2ImModel(vocab, **bench.model_kwargs(cfg))
      d_img = PatchDiscriminator(bench.D_ARCH, padding='valid')
      d_obj = AcCropDiscriminator(vocab, bench.D_ARCH, 'batch', 'leakyrelu-0.2', 32, 'valid')
    step = TrainStep(model, d_obj, d_img, **kw)
    batch = synth_batch(seed=0, **cfg)
    null.calls.clear()
    step.step(batch)
    c = null.calls
    total = sum(c.values())
    print('%-40s %4d library calls' % (label, total))
    for k, v in c.most_common(10):
      print('    %-32s %4d' % (k, v))
  ops.PACK_BOTH = False
  ops.set_conv_math('fp32')


if __name__ == '__main__':
  main()
