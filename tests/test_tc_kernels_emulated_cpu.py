"""The tensor-core kernel SOURCES executed on the CPU against a functional model of TMA,
mbarriers, tcgen05 / TMEM (tests/emul/tc_emul.h).

`conv_tc_kernel`, `conv_tc_halo_kernel` and `conv_wgrad_tc_kernel` — packed and in-place
(weight-gradient layout) weights — are proven on the B200 in their kind::tf32 form
(tests/test_gpu_ops.py, tests/test_gpu_next_rows.py); under the model they must reproduce exact
convolutions, which pins the model's reading of swizzles, shared-memory descriptors (K-major and
MN-major, row-shifted starts), TMEM addressing and the producer / MMA / epilogue barrier
protocol.  The same sources are then run in their kind::f16 form ('bf16x3' / 'bf16': converter
warps split every landed tile into bf16 hi / mid halves in place): the split itself, the
hi / mid descriptor offsets, the converter <-> MMA barrier protocol (also under the model's
adversarial asynchronous schedules), and the accuracy claim — arbitrary fp32 operands agree with
an fp64 convolution to ~2^-16, bf16-exact operands to fp32 rounding.  No timing.
"""
import ctypes
import glob
import os
import shutil
import subprocess

import pytest
import torch
import torch.nn.functional as F

from conftest import ROOT, rel_err

CUDA_INC = next((d for d in (os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'include'),
                             '/usr/local/cuda/include') if os.path.exists(os.path.join(d, 'cuda.h'))), None)
pytestmark = pytest.mark.skipif(shutil.which('g++') is None or CUDA_INC is None,
                                reason='needs g++ (C++20) and the CUDA headers (cuda.h)')

TF32, BF16X3, BF16 = 0, 1, 2          # SG2IM_MATH_* of include/sg2im_b200.h
MATHS = [TF32, BF16X3, BF16]
MATH_IDS = ['tf32', 'bf16x3', 'bf16']


@pytest.fixture(scope='module')
def lib(tmp_path_factory):
  from sg2im_b200._lib import SIGNATURES
  out = tmp_path_factory.mktemp('emultc') / 'libemul_tc.so'
  src = sorted(glob.glob(os.path.join(ROOT, 'tests', 'emul', 'emul*.cpp')))
  subprocess.check_call(['g++'] + os.environ.get('SG2IM_EMUL_CXXFLAGS', '').split() + ['-std=c++20', '-O1', '-pthread', '-shared', '-fPIC', '-Wno-psabi', '-U_FORTIFY_SOURCE', '-DSG2IM_EMUL',
                         '-I', os.path.join(ROOT, 'tests', 'emul'), '-I', os.path.join(ROOT, 'include'),
                         '-I', os.path.join(ROOT, 'sg2im_b200', 'csrc'), '-I', CUDA_INC] + src +
                        ['-o', str(out)])
  L = ctypes.CDLL(str(out))
  for name, sig in SIGNATURES.items():
    if hasattr(L, name):
      getattr(L, name).argtypes = sig
  L.emul_last_error.restype = ctypes.c_char_p
  return L


def _p(t):
  return None if t is None else t.data_ptr()


def _exact(t, math):
  """Values the tensor core consumes exactly in this arithmetic: TF32 (13 low mantissa bits
  clear) or bf16 (16 low bits clear); 'bf16x3' takes arbitrary fp32 values."""
  if math == BF16X3:
    return t
  mask = ~0x1fff if math == TF32 else ~0xffff
  return (t.view(torch.int32) & mask).view(torch.float32)


def _tol(math):
  # bf16x3: operands carry 2^-17 relative error each, the dropped mid*mid term 2^-18
  return 3e-5 if math == BF16X3 else 2e-6


@pytest.fixture
def env():
  keys = ('SG2IM_NO_HALO', 'SG2IM_TC_BN', 'SG2IM_HALO_BN', 'SG2IM_EMUL_ASYNC_SLOW3D', 'SG2IM_EMUL_SLOW_EPILOGUE',
          'SG2IM_EMUL_SMS', 'SG2IM_EMUL_SLOW_PIPE')
  def set_(**kw):
    for k in keys:
      os.environ.pop(k, None)
    for k, v in kw.items():
      os.environ[k] = str(v)
  yield set_
  for k in keys:
    os.environ.pop(k, None)


FWD_CASES = [  # N, H, W, Ci, Co, K, P, env
    (1, 8, 8, 32, 64, 3, 1, {}),                               # per-tap kernel (H < 16)
    (2, 16, 16, 32, 64, 3, 1, {}),                             # halo kernel
    (2, 16, 16, 32, 64, 3, 1, {'SG2IM_NO_HALO': 1}),           # same shape, per-tap kernel
    (1, 16, 8, 72, 36, 3, 1, {}),                              # ragged channels (Cin, Cout % 32 != 0)
    (2, 20, 12, 32, 64, 3, 1, {}),                             # ragged spatial tiles
    (4, 1, 1, 64, 128, 1, 0, {}),                              # Linear
    (1, 8, 8, 64, 256, 3, 1, {'SG2IM_TC_BN': 256}),            # N tile 256
    (1, 8, 8, 64, 256, 3, 1, {'SG2IM_TC_BN': 128}),
    (2, 15, 15, 48, 32, 2, 0, {}),                             # 2x2 taps of the space-to-depth route
    (3, 16, 16, 96, 128, 3, 1, {'SG2IM_HALO_BN': 128}),        # halo kernel, Cout tile 128 (bf16 arithmetic)
    (2, 32, 8, 40, 160, 3, 1, {'SG2IM_HALO_BN': 128}),         # ... ragged second Cout tile, ragged channels
    (5, 17, 9, 32, 256, 2, 0, {'SG2IM_EMUL_SMS': 2, 'SG2IM_HALO_BN': 128})]   # ... weight tiles refilled in flight


@pytest.mark.parametrize('math', MATHS, ids=MATH_IDS)
@pytest.mark.parametrize('N,H,W,Ci,Co,K,P,e', FWD_CASES)
def test_forward_kernels(lib, env, N, H, W, Ci, Co, K, P, e, math):
  env(**e)
  g = torch.Generator().manual_seed(Ci + Co)
  x = _exact(torch.randn(N, H, W, Ci, generator=g), math)
  w = _exact(torch.randn(Co, Ci, K, K, generator=g) * 0.1, math)
  b = torch.randn(Co, generator=g)
  wt = w.permute(2, 3, 0, 1).reshape(K * K, Co, Ci).contiguous()
  Ho, Wo = H + 2 * P - K + 1, W + 2 * P - K + 1
  extra = 8
  y = torch.full((N, Ho, Wo, Co + extra), 7.0)
  stats = torch.zeros(2 * Co, dtype=torch.float64)
  assert lib.sg2im_conv_tc(_p(x), Ci, N, H, W, Ci, _p(wt), _p(b), K, K, P, Ho, Wo, Co, 0, 0.0, _p(y),
                           Co + extra, extra, _p(stats), 0, math, None) == 0, lib.emul_last_error()
  ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), b.double(), padding=P).permute(0, 2, 3, 1)
  assert rel_err(y[..., extra:], ref) < _tol(math)
  assert bool((y[..., :extra] == 7.0).all())               # channel slice of a wider buffer
  # fused BatchNorm statistics of the epilogue
  assert torch.allclose(stats[:Co], ref.sum((0, 1, 2)), rtol=1e-4, atol=1e-3)
  assert torch.allclose(stats[Co:], (ref ** 2).sum((0, 1, 2)), rtol=1e-4, atol=1e-3)
  # fused LeakyReLU (+ RN-TF32 output rounding, a 'tf32' contract the bf16 modes ignore)
  y2 = torch.empty(N, Ho, Wo, Co)
  assert lib.sg2im_conv_tc(_p(x), Ci, N, H, W, Ci, _p(wt), _p(b), K, K, P, Ho, Wo, Co, 1, 0.2, _p(y2),
                           Co, 0, None, 1, math, None) == 0
  if math == TF32:
    assert int((y2.view(torch.int32) & 0x1fff).abs().max()) == 0
    assert rel_err(y2, F.leaky_relu(ref, 0.2)) < 2.0 ** -10
  else:
    assert rel_err(y2, F.leaky_relu(ref, 0.2)) < _tol(math)


def test_bf16x3_beats_tf32_by_orders_of_magnitude_on_arbitrary_operands(lib, env):
  """The point of the mode: on operands that are NOT pre-rounded, kind::tf32 (which truncates) is
  ~1e-3 off an fp64 convolution, plain bf16 ~1e-2, the three-product bf16 form ~1e-5."""
  env()
  g = torch.Generator().manual_seed(5)
  N, H, W, Ci, Co = 2, 16, 16, 96, 64
  x = torch.randn(N, H, W, Ci, generator=g)
  w = torch.randn(Co, Ci, 3, 3, generator=g) * 0.1
  wt = w.permute(2, 3, 0, 1).reshape(9, Co, Ci).contiguous()
  ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), padding=1).permute(0, 2, 3, 1)
  err = {}
  for math in MATHS:
    y = torch.empty(N, H, W, Co)
    assert lib.sg2im_conv_tc(_p(x), Ci, N, H, W, Ci, _p(wt), None, 3, 3, 1, H, W, Co, 0, 0.0, _p(y), Co, 0,
                             None, 0, math, None) == 0, lib.emul_last_error()
    err[math] = rel_err(y, ref)
  assert err[BF16X3] < 2e-5 and err[TF32] > 20 * err[BF16X3] and err[BF16] > 100 * err[BF16X3]


WG_CASES = [  # N, H, W, Ci, Co, K
    (2, 8, 8, 32, 64, 3), (2, 16, 16, 160, 128, 3), (1, 8, 8, 96, 256, 3), (32, 1, 1, 128, 128, 1),
    (2, 16, 24, 64, 192, 3), (3, 9, 11, 32, 64, 3), (2, 15, 15, 64, 64, 2), (2, 8, 8, 12, 32, 2),
    (8, 64, 32, 160, 64, 3)]            # more work items than SMs: the persistent loop re-enters,
                                        # pipeline slots and barrier phases wrap many times


def _wgrad(lib, N, H, W, Ci, Co, K, math):
  g = torch.Generator().manual_seed(Ci + Co)
  P = 1 if K == 3 else 0
  x = _exact(torch.randn(N, H, W, Ci, generator=g), math)
  Ho, Wo = H + 2 * P - K + 1, W + 2 * P - K + 1
  dy = _exact(torch.randn(N, Ho, Wo, Co, generator=g), math)
  dw = torch.zeros(K * K * Ci, Co)
  assert lib.sg2im_conv_wgrad_tc(_p(x), Ci, N, H, W, Ci, _p(dy), K, K, P, Ho, Wo, Co, _p(dw), math, 0, None) == 0, \
      lib.emul_last_error()
  ref = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2).double(), (Co, Ci, K, K),
                                    dy.permute(0, 3, 1, 2).double(), padding=P)
  return dw, ref.permute(2, 3, 1, 0).reshape(K * K * Ci, Co).float()


@pytest.mark.parametrize('math', MATHS, ids=MATH_IDS)
@pytest.mark.parametrize('N,H,W,Ci,Co,K', WG_CASES)
def test_weight_gradient_kernel(lib, env, N, H, W, Ci, Co, K, math):
  env(SG2IM_EMUL_SMS=8) if N * H * W > 4096 else env()
  dw, ref = _wgrad(lib, N, H, W, Ci, Co, K, math)
  assert rel_err(dw, ref) < _tol(math)


KCC_CASES = [  # N, H, W, Ci, Co, K, P, Ci_full, env
    (1, 8, 8, 32, 64, 3, 1, 32, {}),                           # per-tap kernel
    (2, 16, 16, 32, 64, 3, 1, 32, {}),                         # halo kernel
    (2, 16, 16, 64, 96, 3, 1, 64, {'SG2IM_NO_HALO': 1}),       # ragged N tile
    (1, 16, 8, 72, 36, 3, 1, 80, {}),                          # channel prefix of wider weights, ragged
    (4, 1, 1, 64, 128, 1, 0, 64, {}),                          # Linear
    (1, 8, 8, 64, 256, 3, 1, 64, {'SG2IM_TC_BN': 256}),
    (1, 8, 8, 64, 256, 3, 1, 64, {'SG2IM_TC_BN': 128}),
    (2, 15, 15, 48, 32, 2, 0, 48, {}),
    (3, 16, 24, 64, 128, 3, 1, 64, {'SG2IM_HALO_BN': 128}),    # halo kernel, Cout tile 128
    (2, 16, 16, 96, 192, 3, 1, 104, {'SG2IM_EMUL_SMS': 2, 'SG2IM_HALO_BN': 128})]   # ... two Cout tiles, channel prefix


@pytest.mark.parametrize('math', MATHS, ids=MATH_IDS)
@pytest.mark.parametrize('N,H,W,Ci,Co,K,P,Cf,e', KCC_CASES)
def test_convolution_straight_from_the_weight_gradient_layout(lib, env, N, H, W, Ci, Co, K, P, Cf, e, math):
  """sg2im_conv_tc_kcc (WMODE 1: forward with an MN-major B operand — in the bf16 arithmetic the
  converters fold pairs of 32-co atoms into 64-co hi / mid atoms; WMODE 2: data gradient with the
  tap flip in the TMA coordinate) — weights in [tap][Cin][Cout], no pack pass."""
  env(**e)
  g = torch.Generator().manual_seed(Ci * 3 + Co)
  T = K * K
  x = _exact(torch.randn(N, H, W, Ci, generator=g), math)
  w_full = _exact(torch.randn(Co, Cf, K, K, generator=g) * 0.1, math)   # OIHW with Cf >= Ci input channels
  w = w_full[:, :Ci]
  b = torch.randn(Co, generator=g)
  kcc = w_full.permute(2, 3, 1, 0).reshape(T, Cf, Co).contiguous()    # [tap][ci][co]: the wgrad layout
  Ho, Wo = H + 2 * P - K + 1, W + 2 * P - K + 1
  y = torch.full((N, Ho, Wo, Co + 4), 7.0)
  assert lib.sg2im_conv_tc_kcc(_p(x), Ci, N, H, W, Ci, _p(kcc), Cf, 0, _p(b), K, K, P, Ho, Wo, Co, 1, 0.2,
                               _p(y), Co + 4, 4, None, 0, math, None) == 0, lib.emul_last_error()
  xr = x.double().permute(0, 3, 1, 2).clone().requires_grad_(True)
  pre = F.conv2d(xr, w.double(), b.double(), padding=P)
  assert rel_err(y[..., 4:], F.leaky_relu(pre, 0.2).permute(0, 2, 3, 1)) < _tol(math)
  assert bool((y[..., :4] == 7.0).all())
  # data gradient from the very same weight buffer
  gy = _exact(torch.randn(N, Ho, Wo, Co, generator=g), math)
  pre.backward(gy.double().permute(0, 3, 1, 2))
  dx = torch.empty(N, H, W, Ci)
  assert lib.sg2im_conv_tc_kcc(_p(gy), Co, N, Ho, Wo, Co, _p(kcc), Cf, 1, None, K, K, K - 1 - P, H, W, Ci, 0,
                               0.0, _p(dx), Ci, 0, None, 0, math, None) == 0, lib.emul_last_error()
  assert rel_err(dx, xr.grad.permute(0, 2, 3, 1)) < _tol(math)


ADVERSARIAL = [{'SG2IM_EMUL_ASYNC_SLOW3D': 60}, {'SG2IM_EMUL_SLOW_EPILOGUE': 40}, {'SG2IM_EMUL_SLOW_PIPE': 25}]


@pytest.mark.parametrize('sched', ADVERSARIAL, ids=['slow-weights', 'slow-epilogue', 'slow-pipe'])
def test_converter_protocol_under_adversarial_schedules(lib, env, sched):
  """bf16x3 kernels with TWO CTAs doing all the work (SG2IM_EMUL_SMS: pipeline slots, weight sets,
  accumulator sets and barrier phases are reused many times) while the asynchronous model makes
  loads, the tensor pipe or the epilogue lag: a converter that reads a tile before it landed, an
  MMA that reads it before it was split, or a refill before the MMAs retired gives wrong numbers."""
  g = torch.Generator().manual_seed(3)
  # per-tap kernel, packed weights
  env(SG2IM_EMUL_SMS=2, SG2IM_TC_BN=64, **sched)
  N, H, W, Ci, Co = 12, 4, 4, 64, 128
  x = torch.randn(N, H, W, Ci, generator=g)
  w = torch.randn(Co, Ci, 3, 3, generator=g) * 0.1
  wt = w.permute(2, 3, 0, 1).reshape(9, Co, Ci).contiguous()
  y = torch.empty(N, H, W, Co)
  assert lib.sg2im_conv_tc(_p(x), Ci, N, H, W, Ci, _p(wt), None, 3, 3, 1, H, W, Co, 0, 0.0, _p(y), Co, 0,
                           None, 0, BF16X3, None) == 0, lib.emul_last_error()
  ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), None, padding=1).permute(0, 2, 3, 1)
  assert rel_err(y, ref) < 3e-5
  # halo kernel, in-place weights (MN-major B): 6 work items on 2 CTAs
  env(SG2IM_EMUL_SMS=2, **sched)
  N, H, W, Ci, Co = 3, 32, 16, 64, 128
  x = torch.randn(N, H, W, Ci, generator=g)
  w = torch.randn(Co, Ci, 3, 3, generator=g) * 0.1
  kw = w.permute(2, 3, 1, 0).reshape(9, Ci, Co).contiguous()
  y = torch.empty(N, H, W, Co)
  assert lib.sg2im_conv_tc_kcc(_p(x), Ci, N, H, W, Ci, _p(kw), Ci, 0, None, 3, 3, 1, H, W, Co, 0, 0.0,
                               _p(y), Co, 0, None, 0, BF16X3, None) == 0, lib.emul_last_error()
  ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), None, padding=1).permute(0, 2, 3, 1)
  assert rel_err(y, ref) < 3e-5
  # halo kernel with the 128-wide Cout tile and ONE weight set refilled tap by tap: 2 CTAs, 8 items
  env(SG2IM_EMUL_SMS=2, SG2IM_HALO_BN=128, **sched)
  N, H, W, Ci, Co = 4, 32, 16, 96, 256
  x = torch.randn(N, H, W, Ci, generator=g)
  w = torch.randn(Co, Ci, 3, 3, generator=g) * 0.1
  wt = w.permute(2, 3, 0, 1).reshape(9, Co, Ci).contiguous()
  y = torch.empty(N, H, W, Co)
  assert lib.sg2im_conv_tc(_p(x), Ci, N, H, W, Ci, _p(wt), None, 3, 3, 1, H, W, Co, 0, 0.0, _p(y), Co, 0,
                           None, 0, BF16X3, None) == 0, lib.emul_last_error()
  ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), None, padding=1).permute(0, 2, 3, 1)
  assert rel_err(y, ref) < 3e-5
  # weight gradient
  env(SG2IM_EMUL_SMS=2, **sched)
  dw, ref = _wgrad(lib, 4, 32, 16, 96, 64, 3, BF16X3)
  assert rel_err(dw, ref) < 3e-5


def test_operand_split_is_exact_to_bf16_pairs(lib):
  """hi + mid reproduces x to 2^-17 relative: run a 1x1 'convolution' with an identity weight —
  bf16x3 then returns hi*1 + mid*1 (+ hi*0) = the split operand itself."""
  g = torch.Generator().manual_seed(9)
  C = 64
  x = torch.randn(128, 1, 1, C, generator=g) * torch.logspace(-6, 6, 128).view(128, 1, 1, 1)
  wt = torch.eye(C).view(1, C, C).contiguous()
  y = torch.empty(128, 1, 1, C)
  assert lib.sg2im_conv_tc(_p(x), C, 128, 1, 1, C, _p(wt), None, 1, 1, 0, 1, 1, C, 0, 0.0, _p(y), C, 0,
                           None, 0, BF16X3, None) == 0, lib.emul_last_error()
  rel = ((y - x).abs() / x.abs().clamp(min=1e-30)).max()
  assert float(rel) < 2.0 ** -16
  y1 = torch.empty(128, 1, 1, C)
  assert lib.sg2im_conv_tc(_p(x), C, 128, 1, 1, C, _p(wt), None, 1, 1, 0, 1, 1, C, 0, 0.0, _p(y1), C, 0,
                           None, 0, BF16, None) == 0
  assert torch.equal(y1, x.bfloat16().float())             # plain bf16: round-to-nearest-even hi only


def _retile_4x4(kcc16, C, Co):
  """[ky*4+kx][c][co] of a 4x4 stride-2 filter -> [(ty,tx)][(py,px,c)][co] of the equivalent 2x2
  stride-1 filter on the space-to-depth input (ky = 2 ty + py, kx = 2 tx + px)."""
  return (kcc16.reshape(2, 2, 2, 2, C, Co).permute(0, 2, 1, 3, 4, 5).reshape(4, 4 * C, Co).contiguous())


@pytest.mark.parametrize('C,Co', [(3, 64), (64, 128), (40, 32)])
def test_stride2_filter_in_place(lib, env, C, Co):
  """The discriminators' 4x4 stride-2 filters stay in their own [16][C][Cout] order: the split
  kernel writes the re-tiled 2x2 copies straight from it (table field 8 = C) and the weight
  gradient kernel maps its rows back (s2d_channels = C) — both equal the plain kernels applied to
  an explicitly re-tiled filter, bit for bit / up to the order of the atomics."""
  env()
  g = torch.Generator().manual_seed(C + Co)
  w16 = torch.randn(16, C, Co, generator=g) * 0.1
  w2 = _retile_4x4(w16, C, Co)
  Ci = 4 * C
  cip, cop = (Ci + 31) // 32 * 32, (Co + 31) // 32 * 32
  tiles = 4 * (cip // 32) * (cop // 32)
  out = []
  for src, flag in ((w16, C), (w2, 0)):
    fwd, dgr = torch.full((4, Co, cip), 7.0), torch.full((4, Ci, cop), 7.0)
    table = torch.tensor([[src.data_ptr(), fwd.data_ptr(), dgr.data_ptr(), 4, Ci, Co, 0, flag]],
                         dtype=torch.int64)
    assert lib.sg2im_split_weights(_p(table), 1, tiles, None) == 0, lib.emul_last_error()
    out.append((fwd, dgr))
  assert torch.equal(out[0][0].view(torch.int32), out[1][0].view(torch.int32))
  assert torch.equal(out[0][1].view(torch.int32), out[1][1].view(torch.int32))
  # weight gradient: rows land in the 4x4 filter's order
  N, H, W = 2, 9, 7                                          # space-to-depth extent; 2x2 'valid' conv
  x = torch.randn(N, H, W, Ci, generator=g)
  dy = torch.randn(N, H - 1, W - 1, Co, generator=g)
  for math in (BF16X3, TF32):
    d16, d2 = torch.full((16, C, Co), 0.5), torch.full((4, Ci, Co), 0.5)     # ADDS into the slot
    assert lib.sg2im_conv_wgrad_tc(_p(x), Ci, N, H, W, Ci, _p(dy), 2, 2, 0, H - 1, W - 1, Co, _p(d16), math, C,
                                   None) == 0, lib.emul_last_error()
    assert lib.sg2im_conv_wgrad_tc(_p(x), Ci, N, H, W, Ci, _p(dy), 2, 2, 0, H - 1, W - 1, Co, _p(d2), math, 0,
                                   None) == 0, lib.emul_last_error()
    assert rel_err(_retile_4x4(d16, C, Co), d2) < 1e-6
  assert lib.sg2im_conv_wgrad_tc(_p(x), Ci, N, H, W, Ci, _p(dy), 2, 2, 0, H - 1, W - 1, Co, _p(d16), TF32, C + 1,
                                 None) != 0                   # Cin != 4 * s2d_channels


@pytest.mark.parametrize('N,H,W,Ci,Co,K,P,Cf,e', KCC_CASES)
def test_presplit_weight_operands(lib, env, N, H, W, Ci, Co, K, P, Cf, e):
  """sg2im_split_weights (all weights of a network in one launch -> bf16 hi / mid operand copies)
  + sg2im_conv_tc_presplit (the converters then split only the activation tiles): forward and data
  gradient, bit-identical to the in-kernel split of the same weights (same values, same products,
  same accumulation order)."""
  env(**e)
  g = torch.Generator().manual_seed(Ci * 3 + Co)
  T = K * K
  x = torch.randn(N, H, W, Ci, generator=g)
  w_full = torch.randn(Co, Cf, K, K, generator=g) * 0.1
  b = torch.randn(Co, generator=g)
  kcc = w_full.permute(2, 3, 1, 0).reshape(T, Cf, Co).contiguous()
  other = torch.randn(3, 40, 36, generator=g)                  # a second entry of the table
  cip, cop = (Cf + 31) // 32 * 32, (Co + 31) // 32 * 32
  fwd, dgr = torch.full((T, Co, cip), 7.0), torch.full((T, Cf, cop), 7.0)
  f2, d2 = torch.zeros(3, 36, 64), torch.zeros(3, 40, 64)
  t0 = 3 * 2 * 2
  table = torch.tensor([[other.data_ptr(), f2.data_ptr(), d2.data_ptr(), 3, 40, 36, 0, 0],
                        [kcc.data_ptr(), fwd.data_ptr(), dgr.data_ptr(), T, Cf, Co, t0, 0]], dtype=torch.int64)
  total = t0 + T * (cip // 32) * (cop // 32)
  assert lib.sg2im_split_weights(_p(table), 2, total, None) == 0, lib.emul_last_error()
  # hi + mid reproduces the weights to 2^-17; pad channels are zero
  def join(t):                                               # (.., blocks*32 floats) -> hi + mid per channel
    u = t.contiguous().view(torch.int32).reshape(*t.shape[:-1], t.shape[-1] // 32, 2, 16)
    lo16 = (u << 16).view(torch.float32)
    hi16 = (u & ~0xffff).view(torch.float32)
    vals = torch.stack([lo16, hi16], -1).reshape(*t.shape[:-1], t.shape[-1] // 32, 2, 32)
    return (vals[..., 0, :] + vals[..., 1, :]).reshape(*t.shape[:-1], -1)
  wf = join(fwd)                                             # (T, Co, cip)
  ref_f = kcc.permute(0, 2, 1)                               # (T, Co, Cf)
  assert float((wf[..., :Cf] - ref_f).abs().max()) <= 2.0 ** -16 * float(ref_f.abs().max())
  assert float(wf[..., Cf:].abs().max()) == 0 if cip > Cf else True
  wd = join(dgr)                                             # (T flipped, Cf, cop)
  assert float((wd[..., :Co] - kcc.flip(0)).abs().max()) <= 2.0 ** -16 * float(kcc.abs().max())
  # forward / data gradient through the pre-split operands == the in-kernel split, bit for bit
  Ho, Wo = H + 2 * P - K + 1, W + 2 * P - K + 1
  for math in (BF16X3, BF16):
    y0, y1 = torch.empty(N, Ho, Wo, Co), torch.empty(N, Ho, Wo, Co)
    assert lib.sg2im_conv_tc_kcc(_p(x), Ci, N, H, W, Ci, _p(kcc), Cf, 0, _p(b), K, K, P, Ho, Wo, Co, 1, 0.2,
                                 _p(y0), Co, 0, None, 0, math, None) == 0, lib.emul_last_error()
    assert lib.sg2im_conv_tc_presplit(_p(x), Ci, N, H, W, Ci, _p(fwd), cip, Co, _p(b), K, K, P, Ho, Wo, Co, 1,
                                      0.2, _p(y1), Co, 0, None, math, None) == 0, lib.emul_last_error()
    assert torch.equal(y0, y1)
    gy = torch.randn(N, Ho, Wo, Co, generator=g)
    d0, d1 = torch.empty(N, H, W, Ci), torch.empty(N, H, W, Ci)
    assert lib.sg2im_conv_tc_kcc(_p(gy), Co, N, Ho, Wo, Co, _p(kcc), Cf, 1, None, K, K, K - 1 - P, H, W, Ci, 0,
                                 0.0, _p(d0), Ci, 0, None, 0, math, None) == 0, lib.emul_last_error()
    assert lib.sg2im_conv_tc_presplit(_p(gy), Co, N, Ho, Wo, Co, _p(dgr), cop, Cf, None, K, K, K - 1 - P, H, W,
                                      Ci, 0, 0.0, _p(d1), Ci, 0, None, math, None) == 0, lib.emul_last_error()
    assert torch.equal(d0, d1)
