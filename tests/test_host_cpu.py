"""CPU-only tests of the host side: the C-ABI library loads and exports every
symbol include/sg2im_b200.h declares, the module mirror has the reference's
state_dict keys, the factories parse architectures, the synthetic batches
have the collate format, and the product refuses to run without CUDA."""
import contextlib
import io
import os
import re

import pytest
import torch

from conftest import ROOT, load_golden


def _quiet():
  return contextlib.redirect_stdout(io.StringIO())


def test_abi_exports_every_declared_symbol():
  from sg2im_b200 import _lib
  hdr = open(os.path.join(ROOT, 'include', 'sg2im_b200.h')).read()
  declared = set(re.findall(r'\b(sg2im_[a-z0-9_]+)\s*\(', hdr))
  declared.discard('sg2im_stream_t')
  assert len(declared) >= 20
  lib = _lib.load()
  for name in sorted(declared):
    assert hasattr(lib, name), 'missing export %s' % name
  assert set(_lib.SIGNATURES) | {'sg2im_last_error_string'} == declared
  assert lib.sg2im_abi_version() == 1
  assert isinstance(lib.sg2im_last_error_string(), bytes)


def test_invalid_argument_is_an_error_not_a_crash():
  from sg2im_b200 import _lib
  with pytest.raises(RuntimeError) as e:
    _lib.call('sg2im_csr_build', None, 4, 2, 3, 5, None, None, None)     # nroles=3
  assert 'sg2im_csr_build' in str(e.value)


def test_no_cpu_fallback():
  from sg2im_b200 import ops
  with pytest.raises(RuntimeError):
    ops.linear(torch.zeros(2, 4), torch.zeros(3, 4), None)


def test_state_dict_keys_match_golden_reference_keys():
  from sg2im_b200.model import Sg2ImModel
  from sg2im_b200.discriminators import PatchDiscriminator, AcCropDiscriminator
  g = load_golden('train_step.pt')
  with _quiet():
    m = Sg2ImModel(vocab=g['vocab'], **g['kwargs'])
    d_img = PatchDiscriminator(arch=g['arch'], padding='valid')
    d_obj = AcCropDiscriminator(vocab=g['vocab'], arch=g['arch'], normalization='batch',
                                activation='leakyrelu-0.2', padding='valid', object_size=16)
  for net, sd in ((m, g['sd_g']), (d_img, g['sd_img']), (d_obj, g['sd_obj'])):
    mine = net.state_dict()
    assert list(mine.keys()) == list(sd.keys())
    for k in sd:
      assert mine[k].shape == sd[k].shape, k
    net.load_state_dict(sd)


def test_default_state_dict_key_list():
  """Key list spelled out in SURVEY.md §8(b)."""
  from sg2im_b200.model import Sg2ImModel
  from sg2im_b200.synth import make_vocab
  with _quiet():
    m = Sg2ImModel(make_vocab(10, 5), image_size=(64, 64), embedding_dim=8, gconv_dim=8,
                   gconv_hidden_dim=16, mask_size=16, layout_noise_dim=4,
                   refinement_dims=(16, 8))
  keys = set(m.state_dict().keys())
  for k in ['obj_embeddings.weight', 'pred_embeddings.weight', 'gconv.net1.0.weight',
            'gconv.net2.2.bias', 'gconv_net.gconvs.3.net1.2.weight', 'box_net.2.bias',
            'mask_net.1.running_mean', 'mask_net.13.num_batches_tracked', 'mask_net.14.weight',
            'mask_net.16.bias', 'rel_aux_net.0.weight',
            'refinement_net.refinement_modules.0.net.0.weight',
            'refinement_net.refinement_modules.1.net.4.running_var',
            'refinement_net.output_conv.2.weight']:
    assert k in keys, k
  assert m.state_dict()['refinement_net.refinement_modules.0.net.0.weight'].shape == (16, 13, 3, 3)


def test_activation_quirk_and_factories():
  from sg2im_b200.layers import get_activation, build_cnn, build_mlp, get_normalization_2d
  assert get_activation('relu').negative_slope == 0.01            # layers.py:39 quirk
  assert get_activation('leakyrelu-0.2').negative_slope == 0.2
  assert get_activation('leakyrelu').negative_slope == 0.01
  assert get_normalization_2d(4, 'none') is None
  with pytest.raises(ValueError):
    get_normalization_2d(4, 'bogus')
  with _quiet():
    cnn, C = build_cnn('I5,C4-64-2,C4-128-2,C4-256-2', padding='valid')
  assert C == 256
  kinds = [type(l).__name__ for l in cnn]
  assert kinds == ['Conv2d', 'BatchNorm2d', 'LeakyReLU', 'Conv2d', 'BatchNorm2d', 'LeakyReLU', 'Conv2d']
  assert cnn[0].in_channels == 5 and cnn[0].stride == (2, 2) and cnn[0].padding == (0, 0)
  with pytest.raises(ValueError):
    with _quiet():
      build_cnn('X3')
  mlp = build_mlp([8, 16, 4])
  assert [type(l).__name__ for l in mlp] == ['Linear', 'ReLU', 'Linear', 'ReLU']
  assert [type(l).__name__ for l in build_mlp([8, 4], final_nonlinearity=False)] == ['Linear']


def test_encode_scene_graphs_matches_reference_encoding():
  import copy
  from sg2im_b200.model import Sg2ImModel
  g = load_golden('sheep.pt')
  with _quiet():
    m = Sg2ImModel(vocab=g['vocab'], **g['kwargs'])
  enc = m.encode_scene_graphs(copy.deepcopy(g['scene_graphs']))
  for a, b in zip(enc, g['encoded']):
    assert torch.equal(a, b)
  with pytest.raises(ValueError):
    m.encode_scene_graphs({'objects': ['unicorn'], 'relationships': []})


def test_synthetic_batch_format():
  from sg2im_b200.synth import synth_config
  (imgs, objs, boxes, triples, o2i, t2i), cfg = synth_config('vg128')
  assert imgs.shape == (32, 3, 128, 128) and objs.shape == (320,) and triples.shape == (448, 3)
  assert boxes.shape == (320, 4) and (boxes[:, 2:] > boxes[:, :2]).all()
  assert torch.equal(o2i, torch.sort(o2i).values)
  assert (objs.view(32, 10)[:, -1] == 0).all()                 # __image__ last per image
  assert (o2i[triples[:, 0]] == o2i[triples[:, 2]]).all()      # triples never cross images
  assert (o2i[triples[:, 0]] == t2i).all()
  batch, cfg = synth_config('coco64')
  assert len(batch) == 7 and batch[3].shape == (224, 16, 16) and batch[3].dtype == torch.int64
  assert batch[4].shape == (384, 3)


def test_losses_match_oracle():
  from sg2im_b200 import losses
  from oracle import sg2im_oracle as orc
  torch.manual_seed(0)
  r, f = torch.randn(5, 1), torch.randn(5, 1)
  assert torch.allclose(losses.gan_g_loss(f), orc.gan_g_loss(f))
  assert torch.allclose(losses.gan_d_loss(r, f), orc.gan_d_loss(r, f))
  with pytest.raises(ValueError):
    losses.get_gan_losses('nope')


def test_bench_reference_arm_contract():
  """`bench.py --impl reference` (the CPU port of the reference step on the host
  cores) prints one JSON line with the contract's keys; tiny workload so it runs
  in seconds.  Under torchrun only rank 0 prints (covered by the same code path:
  RANK != 0 returns immediately)."""
  import json
  import subprocess
  import sys
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                        '--workload', 'tiny32', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
  assert out.returncode == 0, out.stderr[-2000:]
  line = json.loads([l for l in out.stdout.splitlines() if l.startswith('{')][-1])
  assert line['impl'] == 'reference' and line['unit'] == 'images/s' and line['value'] > 0
  assert line['higher_is_better'] is True and line['vs_baseline'] is None
  assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
  assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['value'] == line['value']
  env = dict(os.environ, RANK='1', WORLD_SIZE='2')
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                        '--workload', 'tiny32', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
  assert out.returncode == 0 and out.stdout.strip() == ''


def test_tensor_core_eligibility_rules():
  """sg2im_conv_tc_supported / sg2im_conv_wgrad_tc_supported are pure host
  functions: which shapes of the benchmark take the tcgen05 kernels."""
  from sg2im_b200 import _lib
  lib = _lib.load()
  fwd = lib.sg2im_conv_tc_supported
  wg = lib.sg2im_conv_wgrad_tc_supported
  # args: N, Hin, Win, Cin, x_cstride, KH, KW, S, P, Hout, Wout, Cout, y_cstride, y_coff
  assert fwd(32, 128, 128, 288, 288, 3, 3, 1, 1, 128, 128, 64, 64, 0) == 1      # CRN stage-4 conv1
  assert fwd(32, 8, 8, 160, 160, 3, 3, 1, 1, 8, 8, 1024, 1024, 0) == 1          # stage 0 (zero channel dropped)
  assert fwd(32, 8, 8, 161, 161, 3, 3, 1, 1, 8, 8, 1024, 1024, 0) == 0          # 161 channels: 644-byte pixel stride
  assert fwd(448, 1, 1, 384, 384, 1, 1, 1, 0, 1, 1, 512, 512, 0) == 1           # gconv Linear as 1x1 conv
  assert fwd(32, 32, 32, 256, 256, 2, 2, 1, 0, 30, 30, 128, 128, 0) == 1        # s2d discriminator conv, cropped output
  assert fwd(32, 64, 64, 12, 12, 2, 2, 1, 0, 63, 63, 64, 64, 0) == 1            # first D layer on the s2d image
  assert fwd(32, 128, 128, 64, 64, 1, 1, 1, 0, 128, 128, 3, 3, 0) == 0          # RGB head: Cout = 3
  assert fwd(32, 128, 128, 3, 3, 4, 4, 2, 0, 63, 63, 64, 64, 0) == 0            # stride 2 itself is not taken
  assert fwd(32, 64, 64, 128, 128, 3, 3, 1, 1, 64, 64, 128, 288, 160) == 1      # write into a channel slice
  assert fwd(32, 64, 64, 128, 128, 3, 3, 1, 1, 64, 64, 128, 288, 162) == 0      # misaligned slice offset
  # args: N, Hin, Win, Cin, x_cstride, KH, KW, S, P, Hout, Wout, Cout
  assert wg(32, 128, 128, 288, 288, 3, 3, 1, 1, 128, 128, 64) == 1
  assert wg(448, 1, 1, 384, 384, 1, 1, 1, 0, 1, 1, 512) == 1                    # 448 rows: multiple of 32
  assert wg(70, 1, 1, 512, 512, 1, 1, 1, 0, 1, 1, 1152) == 0                    # 70 rows: FFMA kernel
  assert wg(32, 32, 32, 256, 256, 2, 2, 1, 0, 30, 30, 128) == 1                 # s2d conv, ragged 30x30 output
  assert wg(32, 128, 128, 64, 64, 1, 1, 1, 0, 128, 128, 3) == 0                 # Cout = 3 -> skinny kernel
  assert wg(32, 16, 16, 64, 64, 5, 5, 1, 2, 16, 16, 64) == 0                    # K > 3


def test_unsupported_tensor_core_call_is_refused_without_a_device():
  from sg2im_b200 import _lib
  import ctypes
  buf = (ctypes.c_float * 64)()
  ptr = ctypes.addressof(buf)
  with pytest.raises(RuntimeError) as e:
    _lib.call('sg2im_conv_tc', ptr, 161, 1, 4, 4, 161, ptr, None, 3, 3, 1, 4, 4, 64, 0, 0.0, ptr, 64, 0,
              None, 0, 1, None)
  assert 'unsupported shape' in str(e.value)


def test_product_never_imports_the_oracle():
  """oracle/ is test infrastructure: no module of the shipped package may import it
  (statically: no import statement; dynamically: importing every product module
  leaves `oracle` out of sys.modules)."""
  import ast
  import glob
  import os
  import subprocess
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  files = sorted(glob.glob(os.path.join(root, 'sg2im_b200', '*.py')))
  assert len(files) > 10
  for path in files:
    tree = ast.parse(open(path).read(), path)
    for node in ast.walk(tree):
      names = []
      if isinstance(node, ast.Import):
        names = [a.name for a in node.names]
      elif isinstance(node, ast.ImportFrom):
        names = [node.module or '']
      for n in names:
        assert n.split('.')[0] != 'oracle', '%s imports %s' % (path, n)
  mods = [os.path.splitext(os.path.basename(f))[0] for f in files if not f.endswith('__init__.py')]
  code = ('import sys, importlib\n'
          'for m in %r: importlib.import_module("sg2im_b200." + m)\n'
          'bad = [m for m in sys.modules if m == "oracle" or m.startswith("oracle.")]\n'
          'assert not bad, bad\n' % (mods,))
  subprocess.check_call([sys.executable, '-c', code], cwd=root)
