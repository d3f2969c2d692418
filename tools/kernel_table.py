"""Per-kernel CUDA time of one eager training iteration (torch.profiler/CUPTI):
cheaper than an ncu launch list and warm-cache.  Not part of the library."""
import contextlib, io, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from sg2im_b200 import ops
from sg2im_b200.model import Sg2ImModel
from sg2im_b200.discriminators import PatchDiscriminator, AcCropDiscriminator
from sg2im_b200.synth import make_vocab, synth_batch, CONFIGS
from sg2im_b200.train_step import TrainStep
import bench

cfg = dict(CONFIGS['vg128'])
dev = torch.device('cuda:0')
ops.set_conv_math('tf32')
vocab = make_vocab(cfg['num_objs'], cfg['num_preds'])
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
  model = Sg2ImModel(vocab, **bench.model_kwargs(cfg)).to(dev)
  d_img = PatchDiscriminator(bench.D_ARCH, padding='valid').to(dev)
  d_obj = AcCropDiscriminator(vocab, bench.D_ARCH, 'batch', 'leakyrelu-0.2', 32, 'valid').to(dev)
step = TrainStep(model, d_obj, d_img)
batch = [t.to(dev) for t in synth_batch(seed=0, **cfg)]
for _ in range(3):
  step.step(batch)
torch.cuda.synchronize()
steps = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  for _ in range(steps):
    step.step(batch)
  torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
  if ev.device_type.name == 'CUDA':
    name = ev.name.replace('(anonymous namespace)::', '').replace('void ', '')
    name = name.split('(')[0][:78]
    agg[name][0] += 1
    agg[name][1] += ev.device_time if hasattr(ev, 'device_time') else ev.cuda_time
tot = sum(v[1] for v in agg.values())
print('total CUDA kernel time per step: %.3f ms, launches per step %d' % (tot / steps / 1e3, sum(v[0] for v in agg.values()) / steps))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
  print('%-80s %5.0f %8.3f ms %5.1f%%' % (k, v[0] / steps, v[1] / steps / 1e3, 100 * v[1] / tot))
