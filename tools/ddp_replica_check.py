"""torchrun helper (N >= 2 GPUs): every rank builds the three networks from a DIFFERENT seed, wraps
them in TrainStep (which must broadcast rank 0's parameters and buffers), runs three iterations on
different shards and checks that all replicas hold bit-identical parameters afterwards."""
import contextlib
import io
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sg2im_b200 import ops  # noqa: E402
from sg2im_b200.model import Sg2ImModel  # noqa: E402
from sg2im_b200.discriminators import PatchDiscriminator, AcCropDiscriminator  # noqa: E402
from sg2im_b200.synth import make_vocab, synth_batch  # noqa: E402
from sg2im_b200.train_step import TrainStep  # noqa: E402


def main():
  rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  dist.init_process_group('nccl', device_id=dev)
  ops.set_conv_math('bf16x3')
  vocab = make_vocab(9, 5)
  torch.manual_seed(100 + rank)                       # deliberately different replicas
  kw = dict(image_size=(32, 32), embedding_dim=32, gconv_dim=32, gconv_hidden_dim=64, gconv_num_layers=3,
            refinement_dims=(64, 32), mask_size=8, layout_noise_dim=8)
  with contextlib.redirect_stdout(io.StringIO()):
    m = Sg2ImModel(vocab, **kw).to(dev)
    d_img = PatchDiscriminator('C4-16-2,C4-32-2', padding='valid').to(dev)
    d_obj = AcCropDiscriminator(vocab, 'C4-16-2,C4-32-2', 'batch', 'leakyrelu-0.2', 16, 'valid').to(dev)
  step = TrainStep(m, d_obj, d_img, weights='kcc', fused_adam='flat')
  for it in range(3):
    batch = [t.to(dev) for t in synth_batch(N=4, objs_per_img=3, rels_per_img=2, image_size=(32, 32),
                                            num_objs=9, num_preds=5, seed=1000 * rank + it)]
    step.step(batch)
  flat = torch.cat([p.detach().reshape(-1) for net in (m, d_obj, d_img) for p in net.parameters()])
  gathered = [torch.zeros_like(flat) for _ in range(dist.get_world_size())]
  dist.all_gather(gathered, flat)
  same = all(torch.equal(gathered[0], g) for g in gathered)
  if rank == 0:
    print('replicas identical after 3 iterations from different seeds:', same)
  dist.barrier()
  torch.cuda.synchronize()
  os._exit(0 if same else 3)


if __name__ == '__main__':
  main()
