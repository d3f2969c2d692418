"""Per-layer attribution of the TF32 deviation of the generator forward (CPU, oracle only).

The oracle's conv2d / linear calls are numbered in call order; each experiment rounds the
operands of a chosen SET of calls to nearest TF32 (fp32 accumulate) and reports
max|out - fp32| / max|fp32| of the image, boxes, masks and relationship scores.  Used to pick
which layers the selective 3xTF32 mode has to compensate.

  python tools/tf32_attribution.py [--size 128] [--n 2]
"""
import argparse
import contextlib
import io
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sg2im_oracle as orc            # noqa: E402
from sg2im_b200.synth import make_vocab, synth_batch  # noqa: E402


def rn(t):
  u = t.contiguous().view(torch.int32)
  return ((u + 0x1000) & ~0x1fff).view(torch.float32)


class Selective(object):
  def __init__(self, which):
    self.which, self.i, self.log = which, 0, []

  def __getattr__(self, k):
    return getattr(F, k)

  def _hit(self, kind, x, w):
    i = self.i
    self.i += 1
    self.log.append((i, kind, tuple(x.shape), tuple(w.shape)))
    return self.which is True or (self.which and i in self.which)

  def conv2d(self, x, w, b=None, **kw):
    if self._hit('conv', x, w):
      x, w = rn(x), rn(w)
    return F.conv2d(x, w, b, **kw)

  def linear(self, x, w, b=None):
    if self._hit('lin', x, w):
      x, w = rn(x), rn(w)
    return F.linear(x, w, b)


def rel(a, b):
  return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp(min=1e-30))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--size', type=int, default=128)
  ap.add_argument('--n', type=int, default=2)
  ap.add_argument('--seed', type=int, default=0)
  a = ap.parse_args()
  torch.manual_seed(a.seed)
  sys.path.insert(0, '/root/reference')
  from sg2im.model import Sg2ImModel
  vocab = make_vocab(179, 46)
  with contextlib.redirect_stdout(io.StringIO()):
    m = Sg2ImModel(vocab, image_size=(a.size, a.size), embedding_dim=128, gconv_dim=128,
                   gconv_hidden_dim=512, gconv_num_layers=5, mask_size=16, layout_noise_dim=32,
                   refinement_dims=(1024, 512, 256, 128, 64))
  sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
  imgs, objs, boxes, triples, o2i, _ = synth_batch(N=a.n, objs_per_img=9, rels_per_img=5,
                                                   image_size=(a.size, a.size), num_objs=179,
                                                   num_preds=46, seed=1)
  noise = torch.randn(a.n, 32, a.size, a.size)

  def run(which):
    f = Selective(which)
    saved = orc.F
    orc.F = f
    try:
      with torch.no_grad():
        out = orc.generator_forward({k: v.clone() for k, v in sd.items()}, (a.size, a.size), objs,
                                    triples, o2i, boxes_gt=boxes, noise=noise, training=True,
                                    num_imgs=a.n)
    finally:
      orc.F = saved
    return out, f.log

  ref, log = run(None)
  for e in log:
    print(e)
  L = len(log)
  full, _ = run(True)
  print('ALL tf32:', [('%.2e' % rel(x, y)) for x, y in zip(full, ref)])
  print('-- only call i in TF32 (image, boxes, masks, rel):')
  single = []
  for i in range(L):
    o, _ = run({i})
    r = [rel(x, y) for x, y in zip(o, ref)]
    single.append(r[0])
    print(i, log[i][1], log[i][2], log[i][3], ' '.join('%.2e' % v for v in r))
  order = sorted(range(L), key=lambda i: -single[i])
  print('-- cumulative: the k worst calls compensated (left in fp32), the rest TF32')
  for k in range(0, L + 1, 1):
    keep = set(range(L)) - set(order[:k])
    o, _ = run(keep if keep else None)
    r = [rel(x, y) for x, y in zip(o, ref)]
    print(k, 'compensated', sorted(order[:k]), ' '.join('%.2e' % v for v in r))
    if max(r) < 2e-4:
      break


if __name__ == '__main__':
  main()
