"""Which ATen ops still launch kernels inside one training step of the benchmarked configuration, and
from where: one eager step at the VG-128 benchmark size under a TorchDispatchMode spy; ops grouped by
(op, first frame inside sg2im_b200/ — or "autograd engine" when the dispatch comes from the backward
thread).  Usage (GPU): python tools/aten_sites.py > gpurun_out/aten_sites.txt"""
import collections
import contextlib
import io
import os
import sys
import traceback

import torch
from torch.utils._python_dispatch import TorchDispatchMode

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sg2im_b200 import ops  # noqa: E402
from sg2im_b200.model import Sg2ImModel  # noqa: E402
from sg2im_b200.discriminators import PatchDiscriminator, AcCropDiscriminator  # noqa: E402
from sg2im_b200.synth import CONFIGS, make_vocab, synth_batch  # noqa: E402
from sg2im_b200.train_step import TrainStep  # noqa: E402

VIEW = ('view', 'reshape', 'permute', 'transpose', 't.default', 'expand', 'unsqueeze', 'squeeze', 'slice', 'select',
        'as_strided', 'detach', 'alias', 'empty', 'unbind', 'split', 'narrow', '_unsafe_view', 'unfold', 'is_', 'size',
        'stride', 'numel', 'dim', 'sym_', 'lift', 'set_', 'resize', 'storage', '_local_scalar', 'item', 'result_type')


class Spy(TorchDispatchMode):
  def __init__(self):
    super().__init__()
    self.cnt = collections.Counter()

  def __torch_dispatch__(self, func, types, args=(), kwargs=None):
    name = str(func).replace('aten.', '')
    if not any(v in name for v in VIEW):
      site = 'autograd engine'
      for fr in reversed(traceback.extract_stack(limit=18)):
        if '/sg2im_b200/' in fr.filename:
          site = '%s:%d' % (fr.filename.split('/sg2im_b200/')[-1], fr.lineno)
          break
      self.cnt[(name, site)] += 1
    return func(*args, **(kwargs or {}))


def main():
  dev = torch.device('cuda:0')
  cfg = dict(CONFIGS['vg128'])
  ops.set_conv_math('bf16x3')
  vocab = make_vocab(cfg['num_objs'], cfg['num_preds'])
  torch.manual_seed(0)
  with contextlib.redirect_stdout(io.StringIO()):
    m = Sg2ImModel(vocab, image_size=cfg['image_size'], embedding_dim=128, gconv_dim=128, gconv_hidden_dim=512,
                   gconv_num_layers=5, refinement_dims=(1024, 512, 256, 128, 64), mask_size=16,
                   layout_noise_dim=32).to(dev)
    d_img = PatchDiscriminator('C4-64-2,C4-128-2,C4-256-2', padding='valid').to(dev)
    d_obj = AcCropDiscriminator(vocab, 'C4-64-2,C4-128-2,C4-256-2', 'batch', 'leakyrelu-0.2', 32, 'valid').to(dev)
  step = TrainStep(m, d_obj, d_img, weights='kcc', fused_adam='flat')
  batch = [t.to(dev) for t in synth_batch(seed=0, **cfg)]
  step.step(batch)
  step.step(batch)
  spy = Spy()
  with spy:
    step.step(batch)
  torch.cuda.synchronize()
  total = sum(spy.cnt.values())
  print('%d non-view ATen dispatches in one eager step' % total)
  for (name, site), c in spy.cnt.most_common(70):
    print('%5d  %-42s %s' % (c, name, site))


if __name__ == '__main__':
  main()
