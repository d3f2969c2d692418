"""Stand-alone launcher of the HBM-bound kernels at their benchmark sizes, for ncu captures
(profiles/) and a quick events-based GB/s line per kernel.  Not part of the library.

  python tools/prof_hbm.py [layout|graph|crop|bn|all] [small]
  ncu --set full --clock-control none --import-source on -k regex:layout_fwd -c 1 \
      -o gpurun_out/prof_layout_fwd python tools/prof_hbm.py layout

Sizes: layout / crop / bn at VG-128 (N = 32, O = 320, D = 128, mask 16, noise 32); the scene-graph
gather / ordered pooling at the dense configuration (T = 4096 triples, O = 2112 objects, hidden 512 —
at VG-128 it moves 2.5 MB and sits under the launch-latency floor).  `small` shrinks
everything (smoke test of this script on the emulated device).
Algorithmic bytes are the ones ops.py hands to bench.py's `hbm_kernels` table (ops._call_b)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from sg2im_b200 import ops  # noqa: E402


def _dev():
  return torch.device(os.environ.get('SG2IM_PROF_DEVICE', 'cuda:0'))


def _timed(label, fn, reps=5):
  """Run fn `reps` times under ops.PROFILE_HBM; print per-entry-point GB/s of the last repetition."""
  fn()                                                       # warm-up (allocator, caches)
  if _dev().type == 'cuda':
    torch.cuda.synchronize()
  entries = None
  for _ in range(reps):
    entries = []
    ops.PROFILE_HBM = entries
    try:
      fn()
    finally:
      ops.PROFILE_HBM = None
  if _dev().type == 'cuda':
    torch.cuda.synchronize()
  for name, nbytes, a, b in entries:
    ms = a.elapsed_time(b)
    print('%-10s %-28s %9.2f MB  %8.1f us  %8.1f GB/s' % (label, name.replace('sg2im_', ''), nbytes / 1e6,
                                                        ms * 1e3, nbytes / max(ms, 1e-9) / 1e6))


def layout(small):
  N, per, D, M, H, W, nc = (2, 3, 16, 8, 16, 16, 4) if small else (32, 10, 128, 16, 128, 128, 32)
  O = N * per
  g = torch.Generator().manual_seed(0)
  vecs = torch.randn(O, D, generator=g).to(_dev()).requires_grad_(True)
  xy = torch.rand(O, 2, generator=g) * 0.6
  boxes = torch.cat([xy, xy + 0.15 + 0.25 * torch.rand(O, 2, generator=g)], 1)
  boxes[per - 1::per] = torch.tensor([0.0, 0.0, 1.0, 1.0])           # the __image__ objects
  boxes = boxes.to(_dev())
  masks = torch.rand(O, M, M, generator=g).to(_dev()).requires_grad_(True)
  o2i = torch.arange(N).repeat_interleave(per).to(_dev())
  noise = torch.randn(N, nc, H, W, generator=g).to(_dev())
  dout = torch.randn(N, H, W, D + nc, generator=g).to(_dev())

  def run():
    vecs.grad = masks.grad = None
    out = ops.Layout.apply(vecs, boxes, masks, o2i, N, H, W, noise, False)
    out.backward(dout)
  _timed('layout', run)


def graph(small):
  T, O, Hd, D = (40, 24, 32, 16) if small else (4096, 2112, 512, 128)
  g = torch.Generator().manual_seed(1)
  obj = torch.randn(O, D, generator=g).to(_dev()).requires_grad_(True)
  pred = torch.randn(T, D, generator=g).to(_dev()).requires_grad_(True)
  edges = torch.randint(0, O, (T, 2), generator=g).to(_dev())
  new_t = torch.randn(T, 2 * Hd + D, generator=g).to(_dev()).requires_grad_(True)
  csr = ops.csr_build(edges, 2, O)

  def run():
    obj.grad = pred.grad = new_t.grad = None
    cur = ops.TripleGather.apply(obj, pred, edges, csr)
    pooled, new_p = ops.GraphPool.apply(new_t, edges, csr, Hd, D, O, True)
    (cur.sum() + pooled.sum() + new_p.sum()).backward()
  _timed('graph', run)


def crop(small):
  N, per, H, W, HH = (2, 3, 16, 16, 8) if small else (32, 10, 128, 128, 32)
  B = N * per
  g = torch.Generator().manual_seed(2)
  feats = torch.randn(N, H, W, 3, generator=g).to(_dev()).requires_grad_(True)
  xy = torch.rand(B, 2, generator=g) * 0.6
  boxes = torch.cat([xy, xy + 0.15 + 0.25 * torch.rand(B, 2, generator=g)], 1).to(_dev())
  idx = torch.arange(N).repeat_interleave(per).to(_dev())
  dout = torch.randn(B, HH, HH, 3, generator=g).to(_dev())

  def run():
    feats.grad = None
    ops.Crop.apply(feats, boxes, idx, HH, HH, False).backward(dout)
  _timed('crop', run)


def bn(small):
  """BatchNorm + LeakyReLU (+ nearest x2 upsample into the next stage's buffer) forward and
  backward on the largest CRN activations: 64 channels at 128x128, 128 channels at 64x64 -> x2."""
  import torch.nn as nn
  cases = [(2, 8, 8, 8, 1), (2, 4, 4, 8, 2)] if small else [(32, 128, 128, 64, 1), (32, 64, 64, 128, 2)]
  for N, H, W, C, up in cases:
    g = torch.Generator().manual_seed(3)
    x = torch.randn(N, H, W, C, generator=g).to(_dev()).requires_grad_(True)
    mod = nn.BatchNorm2d(C).to(_dev()).train()
    dout = torch.randn(N, H * up, W * up, C, generator=g).to(_dev())

    def run():
      x.grad = None
      ops.bn_act(x, mod, 0.2, up).backward(dout)
    _timed('bn up=%d' % up, run)


def main():
  what = sys.argv[1] if len(sys.argv) > 1 else 'all'
  small = len(sys.argv) > 2 and sys.argv[2] == 'small'
  for name, fn in (('layout', layout), ('graph', graph), ('crop', crop), ('bn', bn)):
    if what in (name, 'all'):
      fn(small)


if __name__ == '__main__':
  main()
