"""How much of a graph-replayed training step is the GPU idle BETWEEN kernels?  Replays the benchmarked
step (VG-128, 32 images, bf16x3, in-place weights, flat Adam) under torch.profiler (CUPTI kernel
records), then reports per step: span from the first kernel's start to the last kernel's end, the
union of the kernel intervals (busy), their difference (idle), the number of kernels, the histogram
of the idle intervals and which kernels they follow.  Numbers come from a run under the profiler:
they explain the bench value, they are not bench values.
Usage (GPU): python tools/graph_gaps.py > gpurun_out/graph_gaps.txt"""
import collections
import contextlib
import io
import json
import os
import sys
import tempfile

import torch
from torch.profiler import profile, ProfilerActivity

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sg2im_b200 import ops  # noqa: E402
from sg2im_b200.model import Sg2ImModel  # noqa: E402
from sg2im_b200.discriminators import PatchDiscriminator, AcCropDiscriminator  # noqa: E402
from sg2im_b200.synth import CONFIGS, make_vocab, synth_batch  # noqa: E402
from sg2im_b200.train_step import TrainStep  # noqa: E402

STEPS = 3


def short(name):
  name = name.replace('void ', '').replace('(anonymous namespace)::', '').replace('<unnamed>::', '')
  return name.split('(')[0][:70]


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
  for _ in range(6):                                     # eager, capture, replays
    step.step(batch)
  torch.cuda.synchronize()
  with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
      step.step(batch)
      torch.cuda.synchronize()
  path = os.path.join(tempfile.mkdtemp(), 'trace.json')
  prof.export_chrome_trace(path)
  ev = [e for e in json.load(open(path))['traceEvents']
        if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset') and 'dur' in e]
  ev.sort(key=lambda e: e['ts'])
  print('%d GPU activities in %d replayed steps (%d kernels)' % (
      len(ev), STEPS, sum(e['cat'] == 'kernel' for e in ev)))
  # steps are separated by the host synchronize: split at idle intervals > 200 us
  groups, cur = [], [ev[0]]
  for a, b in zip(ev, ev[1:]):
    if b['ts'] - (a['ts'] + a['dur']) > 200.0:
      groups.append(cur)
      cur = []
    cur.append(b)
  groups.append(cur)
  after = collections.Counter()
  after_n = collections.Counter()
  hist = collections.Counter()
  for g in groups:
    span = g[-1]['ts'] + g[-1]['dur'] - g[0]['ts']
    busy, end = 0.0, g[0]['ts']
    gaps = []
    for prev, e in zip([None] + g[:-1], g):
      s, t = e['ts'], e['ts'] + e['dur']
      if s > end:
        if prev is not None:
          gaps.append((s - end, short(prev['name'])))
        busy += t - s
      elif t > end:
        busy += t - end
      end = max(end, t)
    print('step: %4d activities  span %9.1f us  busy %9.1f us  idle %7.1f us (%.1f %%)  %d idle intervals' % (
        len(g), span, busy, span - busy, 100.0 * (span - busy) / span, len(gaps)))
    for d, n in gaps:
      after[n] += d
      after_n[n] += 1
      hist[min(int(d), 10)] += 1
  last = max(groups, key=len)
  tot = collections.Counter()
  cnt = collections.Counter()
  for e in last:
    tot[short(e['name'])] += e['dur']
    cnt[short(e['name'])] += 1
  print('kernel time inside the longest replayed step (CUPTI durations, %d activities, %.1f us busy):' % (
      len(last), sum(tot.values())))
  for n, d in tot.most_common(40):
    print('  %8.1f us  %4d x  %7.2f us each  %s' % (d, cnt[n], d / cnt[n], n))
  print('idle-interval histogram (us, floor; 10 = 10 or more), all steps:', dict(sorted(hist.items())))
  print('idle time by preceding kernel (all steps):')
  for n, d in after.most_common(25):
    print('  %8.1f us  %4d x  %.2f us each  after %s' % (d, after_n[n], d / after_n[n], n))


if __name__ == '__main__':
  main()
