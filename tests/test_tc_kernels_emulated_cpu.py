"""The tensor-core kernel SOURCES executed on the CPU against a functional model of TMA,
mbarriers, tcgen05 / TMEM and thread-block clusters (tests/emul/tc_emul.h).

Calibration first: `conv_tc_kernel`, `conv_tc_halo_kernel` and `conv_wgrad_tc_kernel` are proven
on the B200 (tests/test_gpu_ops.py); under the model they must reproduce exact convolutions —
that pins the model's reading of swizzles, shared-memory descriptors (K-major and MN-major,
row-shifted starts), TMEM addressing and the producer / MMA / epilogue barrier protocol.  The
calibrated model then runs `conv_wgrad_tc_mc_kernel` (cluster of 2 / 4 CTAs, multicast TMA,
multicast tcgen05.commit), which has not run on hardware yet: a protocol slip would deadlock
(bounded waits trap) or corrupt dW here.  No timing, no asynchrony, no memory-ordering claims.
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
  L.emul_cluster_blocks_run.restype = ctypes.c_ulonglong
  return L


def _p(t):
  return None if t is None else t.data_ptr()


def _tf32(t):
  """Values the tensor core consumes exactly (13 low mantissa bits clear)."""
  return (t.view(torch.int32) & ~0x1fff).view(torch.float32)


@pytest.fixture
def env():
  keys = ('SG2IM_NO_HALO', 'SG2IM_TC_BN', 'SG2IM_WGRAD_MC', 'SG2IM_CONV_MC', 'SG2IM_HALO_SMALL', 'SG2IM_HALO_PAIR',
          'SG2IM_EMUL_ASYNC_SLOW3D', 'SG2IM_EMUL_SLOW_EPILOGUE', 'SG2IM_EMUL_SMS', 'SG2IM_EMUL_SLOW_PIPE')
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
    (2, 15, 15, 48, 32, 2, 0, {})]                             # 2x2 taps of the space-to-depth route


@pytest.mark.parametrize('N,H,W,Ci,Co,K,P,e', FWD_CASES)
def test_forward_kernels_calibrate_the_model(lib, env, N, H, W, Ci, Co, K, P, e):
  env(**e)
  g = torch.Generator().manual_seed(Ci + Co)
  x = _tf32(torch.randn(N, H, W, Ci, generator=g))
  w = _tf32(torch.randn(Co, Ci, K, K, generator=g) * 0.1)
  b = torch.randn(Co, generator=g)
  wt = w.permute(2, 3, 0, 1).reshape(K * K, Co, Ci).contiguous()
  Ho, Wo = H + 2 * P - K + 1, W + 2 * P - K + 1
  extra = 8
  y = torch.full((N, Ho, Wo, Co + extra), 7.0)
  stats = torch.zeros(2 * Co, dtype=torch.float64)
  assert lib.sg2im_conv_tc(_p(x), Ci, N, H, W, Ci, _p(wt), _p(b), K, K, P, Ho, Wo, Co, 0, 0.0, _p(y),
                           Co + extra, extra, _p(stats), 0, None) == 0, lib.emul_last_error()
  ref = F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=P).permute(0, 2, 3, 1)
  assert rel_err(y[..., extra:], ref) < 2e-6
  assert bool((y[..., :extra] == 7.0).all())               # channel slice of a wider buffer
  # fused BatchNorm statistics of the epilogue
  assert torch.allclose(stats[:Co], ref.double().sum((0, 1, 2)), rtol=1e-5, atol=1e-4)
  assert torch.allclose(stats[Co:], (ref.double() ** 2).sum((0, 1, 2)), rtol=1e-5, atol=1e-4)
  # fused LeakyReLU + RN-TF32 output rounding
  y2 = torch.empty(N, Ho, Wo, Co)
  assert lib.sg2im_conv_tc(_p(x), Ci, N, H, W, Ci, _p(wt), _p(b), K, K, P, Ho, Wo, Co, 1, 0.2, _p(y2),
                           Co, 0, None, 1, None) == 0
  assert int((y2.view(torch.int32) & 0x1fff).abs().max()) == 0
  assert rel_err(y2, F.leaky_relu(ref, 0.2)) < 2.0 ** -10


WG_CASES = [  # N, H, W, Ci, Co, K
    (2, 8, 8, 32, 64, 3), (2, 16, 16, 160, 128, 3), (1, 8, 8, 96, 256, 3), (32, 1, 1, 128, 128, 1),
    (2, 16, 24, 64, 192, 3), (3, 9, 11, 32, 64, 3), (2, 15, 15, 64, 64, 2),
    (8, 64, 32, 160, 64, 3)]            # more work items than clusters: the persistent loop re-enters,
                                        # pipeline slots and barrier phases wrap many times


def _wgrad(lib, N, H, W, Ci, Co, K):
  g = torch.Generator().manual_seed(Ci + Co)
  P = 1 if K == 3 else 0
  x = _tf32(torch.randn(N, H, W, Ci, generator=g))
  Ho, Wo = H + 2 * P - K + 1, W + 2 * P - K + 1
  dy = _tf32(torch.randn(N, Ho, Wo, Co, generator=g))
  dw = torch.zeros(K * K * Ci, Co)
  assert lib.sg2im_conv_wgrad_tc(_p(x), Ci, N, H, W, Ci, _p(dy), K, K, P, Ho, Wo, Co, _p(dw), None) == 0, \
      lib.emul_last_error()
  ref = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2).double(), (Co, Ci, K, K),
                                    dy.permute(0, 3, 1, 2).double(), padding=P)
  return dw, ref.permute(2, 3, 1, 0).reshape(K * K * Ci, Co).float()


@pytest.mark.parametrize('N,H,W,Ci,Co,K', WG_CASES)
def test_weight_gradient_kernel_calibrates_the_model(lib, env, N, H, W, Ci, Co, K):
  env()
  dw, ref = _wgrad(lib, N, H, W, Ci, Co, K)
  assert rel_err(dw, ref) < 2e-6


@pytest.mark.parametrize('N,H,W,Ci,Co,K', WG_CASES)
def test_cluster_multicast_weight_gradient_kernel(lib, env, N, H, W, Ci, Co, K):
  """conv_wgrad_tc_mc_kernel (not yet run on hardware) under the calibrated model."""
  env(SG2IM_WGRAD_MC=1)
  c0 = lib.emul_cluster_blocks_run()
  dw, ref = _wgrad(lib, N, H, W, Ci, Co, K)
  assert rel_err(dw, ref) < 2e-6
  members = -(-Co // 64) * (2 if K == 3 else 1)            # co tiles x tap passes at N = 64
  if members % 2 == 0:
    assert lib.emul_cluster_blocks_run() > c0              # the cluster kernel really ran
  else:
    assert lib.emul_cluster_blocks_run() == c0             # odd: plain kernel (documented fallback)


KCC_CASES = [  # N, H, W, Ci, Co, K, P, Ci_full, env
    (1, 8, 8, 32, 64, 3, 1, 32, {}),                           # per-tap kernel
    (2, 16, 16, 32, 64, 3, 1, 32, {}),                         # halo kernel
    (2, 16, 16, 64, 96, 3, 1, 64, {'SG2IM_NO_HALO': 1}),       # ragged N tile
    (1, 16, 8, 72, 36, 3, 1, 80, {}),                          # channel prefix of wider weights, ragged
    (4, 1, 1, 64, 128, 1, 0, 64, {}),                          # Linear
    (1, 8, 8, 64, 256, 3, 1, 64, {'SG2IM_TC_BN': 256}),
    (1, 8, 8, 64, 256, 3, 1, 64, {'SG2IM_TC_BN': 128}),
    (2, 15, 15, 48, 32, 2, 0, 48, {})]


@pytest.mark.parametrize('N,H,W,Ci,Co,K,P,Cf,e', KCC_CASES)
def test_convolution_straight_from_the_weight_gradient_layout(lib, env, N, H, W, Ci, Co, K, P, Cf, e):
  """sg2im_conv_tc_kcc (WMODE 1: forward with an MN-major B operand; WMODE 2: data gradient with
  the tap flip in the TMA coordinate) — weights in [tap][Cin][Cout], no pack pass.  Not yet run on
  hardware; executed here under the model calibrated by the tests above."""
  env(**e)
  g = torch.Generator().manual_seed(Ci * 3 + Co)
  T = K * K
  x = _tf32(torch.randn(N, H, W, Ci, generator=g))
  w_full = _tf32(torch.randn(Co, Cf, K, K, generator=g) * 0.1)       # OIHW with Cf >= Ci input channels
  w = w_full[:, :Ci]
  b = torch.randn(Co, generator=g)
  kcc = w_full.permute(2, 3, 1, 0).reshape(T, Cf, Co).contiguous()    # [tap][ci][co]: the wgrad layout
  Ho, Wo = H + 2 * P - K + 1, W + 2 * P - K + 1
  y = torch.full((N, Ho, Wo, Co + 4), 7.0)
  lib.sg2im_conv_tc_kcc.argtypes = __import__('sg2im_b200._lib', fromlist=['x']).SIGNATURES['sg2im_conv_tc_kcc']
  assert lib.sg2im_conv_tc_kcc(_p(x), Ci, N, H, W, Ci, _p(kcc), Cf, 0, _p(b), K, K, P, Ho, Wo, Co, 1, 0.2,
                               _p(y), Co + 4, 4, None, 0, None) == 0, lib.emul_last_error()
  xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
  pre = F.conv2d(xr, w, b, padding=P)
  assert rel_err(y[..., 4:], F.leaky_relu(pre, 0.2).permute(0, 2, 3, 1)) < 2e-6
  assert bool((y[..., :4] == 7.0).all())
  # data gradient from the very same weight buffer
  gy = _tf32(torch.randn(N, Ho, Wo, Co, generator=g))
  pre.backward(gy.permute(0, 3, 1, 2))
  dx = torch.empty(N, H, W, Ci)
  assert lib.sg2im_conv_tc_kcc(_p(gy), Co, N, Ho, Wo, Co, _p(kcc), Cf, 1, None, K, K, K - 1 - P, H, W, Ci, 0,
                               0.0, _p(dx), Ci, 0, None, 0, None) == 0, lib.emul_last_error()
  assert rel_err(dx, xr.grad.permute(0, 2, 3, 1)) < 2e-6


MC_FWD_CASES = [  # N, H, W, Ci, Co, K, P, env  (per-tap kernel, even number of Cout tiles)
    (2, 8, 8, 64, 256, 3, 1, {'SG2IM_TC_BN': 64}),              # 4 Cout tiles -> clusters of 4
    (2, 8, 8, 64, 256, 3, 1, {'SG2IM_TC_BN': 128}),             # 2 Cout tiles -> clusters of 2
    (3, 8, 8, 96, 512, 3, 1, {'SG2IM_TC_BN': 256}),             # ragged image count, N tile 256
    (8, 1, 1, 128, 384, 1, 0, {'SG2IM_TC_BN': 64}),             # Linear, 6 tiles -> clusters of 2
    (2, 16, 16, 32, 128, 3, 1, {'SG2IM_NO_HALO': 1, 'SG2IM_TC_BN': 64}),
    (40, 4, 4, 64, 128, 3, 1, {'SG2IM_TC_BN': 64})]             # many pixel tiles: persistent loop, phase wraps


@pytest.mark.parametrize('N,H,W,Ci,Co,K,P,e', MC_FWD_CASES)
@pytest.mark.parametrize('kcc', [False, True])
def test_cluster_multicast_forward_kernel(lib, env, N, H, W, Ci, Co, K, P, e, kcc):
  """conv_tc_mc_kernel (SG2IM_CONV_MC=1; not yet run on hardware): rank 0 multicasts the activation
  tile to the CS CTAs that own consecutive Cout tiles.  Packed weights and, with kcc, the in-place
  weight-gradient layout (forward MN-major B and data gradient with flipped taps)."""
  env(SG2IM_CONV_MC=1, **e)
  g = torch.Generator().manual_seed(Ci + Co + N)
  T = K * K
  x = _tf32(torch.randn(N, H, W, Ci, generator=g))
  w = _tf32(torch.randn(Co, Ci, K, K, generator=g) * 0.1)
  b = torch.randn(Co, generator=g)
  Ho, Wo = H + 2 * P - K + 1, W + 2 * P - K + 1
  y = torch.empty(N, Ho, Wo, Co)
  c0 = lib.emul_cluster_blocks_run()
  if kcc:
    kw = w.permute(2, 3, 1, 0).reshape(T, Ci, Co).contiguous()
    lib.sg2im_conv_tc_kcc.argtypes = __import__('sg2im_b200._lib', fromlist=['x']).SIGNATURES['sg2im_conv_tc_kcc']
    rc = lib.sg2im_conv_tc_kcc(_p(x), Ci, N, H, W, Ci, _p(kw), Ci, 0, _p(b), K, K, P, Ho, Wo, Co, 0, 0.0,
                               _p(y), Co, 0, None, 0, None)
  else:
    wt = w.permute(2, 3, 0, 1).reshape(T, Co, Ci).contiguous()
    rc = lib.sg2im_conv_tc(_p(x), Ci, N, H, W, Ci, _p(wt), _p(b), K, K, P, Ho, Wo, Co, 0, 0.0, _p(y), Co,
                           0, None, 0, None)
  assert rc == 0, lib.emul_last_error()
  assert lib.emul_cluster_blocks_run() > c0                # the cluster kernel really ran
  xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
  ref = F.conv2d(xr, w, b, padding=P)
  assert rel_err(y, ref.permute(0, 2, 3, 1)) < 2e-6
  if kcc and Ci % 64 == 0:                                 # data gradient: Cout_dgrad = Ci tiles must pair up
    gy = _tf32(torch.randn(N, Ho, Wo, Co, generator=g))
    ref.backward(gy.permute(0, 3, 1, 2))
    dx = torch.empty(N, H, W, Ci)
    assert lib.sg2im_conv_tc_kcc(_p(gy), Co, N, Ho, Wo, Co, _p(kw), Ci, 1, None, K, K, K - 1 - P, H, W, Ci, 0,
                                 0.0, _p(dx), Ci, 0, None, 0, None) == 0, lib.emul_last_error()
    assert rel_err(dx, xr.grad.permute(0, 2, 3, 1)) < 2e-6


SMALL_CASES = [  # N, H, W, Ci, Co, K, P
    (4, 8, 8, 64, 64, 3, 1), (5, 8, 8, 96, 160, 3, 1),       # odd image count: last tile half empty
    (2, 8, 8, 40, 36, 3, 1), (6, 6, 7, 32, 64, 3, 1),        # ragged channels; Hout = 6 rows, W = 7
    (3, 8, 20, 32, 64, 3, 1), (4, 9, 9, 32, 32, 2, 0),       # wide rows (3 column tiles); 2x2 taps -> 8x8 out
    (2, 10, 10, 64, 64, 3, 0)]                               # valid conv: 8x8 output from 10x10


@pytest.mark.parametrize('N,H,W,Ci,Co,K,P', SMALL_CASES)
@pytest.mark.parametrize('kcc', [False, True])
def test_small_image_halo_kernel(lib, env, N, H, W, Ci, Co, K, P, kcc):
  """conv_tc_halo_small_kernel (SG2IM_HALO_SMALL=1; not yet run on hardware): two images per tile
  through a (C, W, N, H)-ordered tensor map, interleaved halo rows, uniform descriptor stride."""
  env(SG2IM_HALO_SMALL=1)
  g = torch.Generator().manual_seed(Ci + Co + N + H)
  T = K * K
  x = _tf32(torch.randn(N, H, W, Ci, generator=g))
  w = _tf32(torch.randn(Co, Ci, K, K, generator=g) * 0.1)
  b = torch.randn(Co, generator=g)
  Ho, Wo = H + 2 * P - K + 1, W + 2 * P - K + 1
  assert 4 < Ho <= 8
  y = torch.full((N, Ho, Wo, Co + 4), 7.0)
  stats = torch.zeros(2 * Co, dtype=torch.float64)
  if kcc:
    kw = w.permute(2, 3, 1, 0).reshape(T, Ci, Co).contiguous()
    lib.sg2im_conv_tc_kcc.argtypes = __import__('sg2im_b200._lib', fromlist=['x']).SIGNATURES['sg2im_conv_tc_kcc']
    rc = lib.sg2im_conv_tc_kcc(_p(x), Ci, N, H, W, Ci, _p(kw), Ci, 0, _p(b), K, K, P, Ho, Wo, Co, 0, 0.0,
                               _p(y), Co + 4, 4, _p(stats), 0, None)
  else:
    wt = w.permute(2, 3, 0, 1).reshape(T, Co, Ci).contiguous()
    rc = lib.sg2im_conv_tc(_p(x), Ci, N, H, W, Ci, _p(wt), _p(b), K, K, P, Ho, Wo, Co, 0, 0.0, _p(y),
                           Co + 4, 4, _p(stats), 0, None)
  assert rc == 0, lib.emul_last_error()
  xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
  ref = F.conv2d(xr, w, b, padding=P)
  assert rel_err(y[..., 4:], ref.permute(0, 2, 3, 1)) < 2e-6
  assert bool((y[..., :4] == 7.0).all())
  assert torch.allclose(stats[:Co], ref.double().sum((0, 2, 3)), rtol=1e-5, atol=1e-4)
  # the same shape on the default (per-tap) kernels: same products, different summation order
  env()
  y0 = torch.empty(N, Ho, Wo, Co)
  wt = w.permute(2, 3, 0, 1).reshape(T, Co, Ci).contiguous()
  assert lib.sg2im_conv_tc(_p(x), Ci, N, H, W, Ci, _p(wt), _p(b), K, K, P, Ho, Wo, Co, 0, 0.0, _p(y0), Co, 0,
                           None, 0, None) == 0
  assert rel_err(y[..., 4:], y0) < 2e-6
  if kcc and K - 1 - P >= 0 and 4 < H <= 8:                 # data gradient also lands on 8-row maps
    env(SG2IM_HALO_SMALL=1)
    gy = _tf32(torch.randn(N, Ho, Wo, Co, generator=g))
    ref.backward(gy.permute(0, 3, 1, 2))
    dx = torch.empty(N, H, W, Ci)
    assert lib.sg2im_conv_tc_kcc(_p(gy), Co, N, Ho, Wo, Co, _p(kw), Ci, 1, None, K, K, K - 1 - P, H, W, Ci, 0,
                                 0.0, _p(dx), Ci, 0, None, 0, None) == 0, lib.emul_last_error()
    assert rel_err(dx, xr.grad.permute(0, 2, 3, 1)) < 2e-6


PAIR_CASES = [  # N, H, W, Ci, Co, K, P
    (1, 16, 16, 32, 64, 3, 1),      # 2 pixel tiles: one (half-padded) group per... one pair, rank 1 all padding
    (2, 32, 16, 64, 64, 3, 1),      # 8 tiles = 2 groups = exactly one pair
    (3, 16, 24, 40, 96, 3, 1),      # 9 tiles -> 3 groups (odd): last pair half padding; ragged channels, 2 Cout tiles
    (5, 48, 8, 96, 128, 3, 1),      # several units per cluster (persistent loop), 2 weight sets in flight
    (2, 17, 9, 32, 64, 2, 0),       # 2x2 taps, partial edge tiles
]


# adversarial schedules of the asynchronous model (tests/emul/tc_emul.h): the peer's weight loads land
# late / the peer's epilogue lags, so a missing cross-CTA wait shows up as wrong numbers
# (SG2IM_EMUL_SMS=2: one cluster runs every work unit, so accumulator sets and weight sets are reused)
SCHEDULES = [{}, {'SG2IM_EMUL_ASYNC_SLOW3D': 60, 'SG2IM_EMUL_SMS': 2}, {'SG2IM_EMUL_SLOW_EPILOGUE': 40, 'SG2IM_EMUL_SMS': 2}]


@pytest.mark.parametrize('N,H,W,Ci,Co,K,P', PAIR_CASES)
@pytest.mark.parametrize('kcc', [False, True])
@pytest.mark.parametrize('sched', SCHEDULES, ids=['plain', 'slow-weights', 'slow-epilogue'])
def test_cta_pair_halo_kernel(lib, env, N, H, W, Ci, Co, K, P, kcc, sched):
  """conv_tc_halo_pair_kernel (SG2IM_HALO_PAIR=1; not yet run on hardware): two CTAs as one M = 256
  tile under the ASSUMED cta_group::2 semantics of the model (tools/umma_2cta_probe.cu pins them on
  hardware): own pixel tiles and accumulators per CTA, half of every weight tile each, relay /
  multicast-commit / remote-release protocol.  Same products in the same order as the single-CTA
  halo kernel => identical bits."""
  g = torch.Generator().manual_seed(Ci + Co + N + H)
  T = K * K
  x = _tf32(torch.randn(N, H, W, Ci, generator=g))
  w = _tf32(torch.randn(Co, Ci, K, K, generator=g) * 0.1)
  b = torch.randn(Co, generator=g)
  Ho, Wo = H + 2 * P - K + 1, W + 2 * P - K + 1
  wt = w.permute(2, 3, 0, 1).reshape(T, Co, Ci).contiguous()
  kw = w.permute(2, 3, 1, 0).reshape(T, Ci, Co).contiguous()
  lib.sg2im_conv_tc_kcc.argtypes = __import__('sg2im_b200._lib', fromlist=['x']).SIGNATURES['sg2im_conv_tc_kcc']

  def run(pair):
    env(**(dict(sched, SG2IM_HALO_PAIR=1) if pair else {}))
    y = torch.full((N, Ho, Wo, Co + 4), 7.0)
    stats = torch.zeros(2 * Co, dtype=torch.float64)
    if kcc:
      rc = lib.sg2im_conv_tc_kcc(_p(x), Ci, N, H, W, Ci, _p(kw), Ci, 0, _p(b), K, K, P, Ho, Wo, Co, 1, 0.2,
                                 _p(y), Co + 4, 4, None, 1, None)      # fused LeakyReLU, RN-TF32 outputs
    else:
      rc = lib.sg2im_conv_tc(_p(x), Ci, N, H, W, Ci, _p(wt), _p(b), K, K, P, Ho, Wo, Co, 0, 0.0, _p(y),
                             Co + 4, 4, _p(stats), 0, None)      # fused BatchNorm statistics
    assert rc == 0, lib.emul_last_error()
    return y, stats

  y, stats = run(True)
  y0, stats0 = run(False)
  assert torch.equal(y, y0)
  assert torch.allclose(stats, stats0, rtol=1e-6, atol=1e-5)      # smem / global atomics in another order
  ref = F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=P)
  if kcc:
    ref = F.leaky_relu(ref, 0.2)
  assert rel_err(y[..., 4:], ref.permute(0, 2, 3, 1)) < (2e-3 if kcc else 2e-6)   # kcc case writes RN-TF32 outputs
  assert bool((y[..., :4] == 7.0).all())
  if kcc and K - 1 - P >= 0:
    gy = _tf32(torch.randn(N, Ho, Wo, Co, generator=g))
    outs = []
    for pair in (True, False):
      env(**(dict(sched, SG2IM_HALO_PAIR=1) if pair else {}))
      dx = torch.empty(N, H, W, Ci)
      assert lib.sg2im_conv_tc_kcc(_p(gy), Co, N, Ho, Wo, Co, _p(kw), Ci, 1, None, K, K, K - 1 - P, H, W, Ci, 0,
                                   0.0, _p(dx), Ci, 0, None, 0, None) == 0, lib.emul_last_error()
      outs.append(dx)
    assert torch.equal(outs[0], outs[1])
    xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
    F.conv2d(xr, w, None, padding=P).backward(gy.permute(0, 3, 1, 2))
    assert rel_err(outs[0], xr.grad.permute(0, 2, 3, 1)) < 2e-6


ADVERSARIAL = [{'SG2IM_EMUL_ASYNC_SLOW3D': 60}, {'SG2IM_EMUL_SLOW_EPILOGUE': 40}, {'SG2IM_EMUL_SLOW_PIPE': 25}]


@pytest.mark.parametrize('sched', ADVERSARIAL, ids=['slow-weights', 'slow-epilogue', 'slow-peer-pipe'])
def test_cluster_kernels_under_adversarial_schedules(lib, env, sched):
  """The cluster-multicast kernels and the CTA-pair kernel once more with ONE cluster doing all the
  work (SG2IM_EMUL_SMS: pipeline slots, accumulator sets and barrier phases are reused many times)
  while the asynchronous model makes the peer CTAs' loads, tensor pipes or epilogues lag."""
  # cluster weight gradient: clusters of 2 (co tiles x tap passes), 4 SMs -> 2 clusters
  env(SG2IM_WGRAD_MC=1, SG2IM_EMUL_SMS=4, **sched)
  dw, ref = _wgrad(lib, 4, 32, 16, 96, 64, 3)
  assert rel_err(dw, ref) < 2e-6
  # cluster forward: 4 Cout tiles of 64 -> one cluster of 4 runs every pixel tile
  env(SG2IM_CONV_MC=1, SG2IM_TC_BN=64, SG2IM_EMUL_SMS=4, **sched)
  g = torch.Generator().manual_seed(3)
  N, H, W, Ci, Co = 20, 4, 4, 64, 256
  x = _tf32(torch.randn(N, H, W, Ci, generator=g))
  w = _tf32(torch.randn(Co, Ci, 3, 3, generator=g) * 0.1)
  wt = w.permute(2, 3, 0, 1).reshape(9, Co, Ci).contiguous()
  y = torch.empty(N, H, W, Co)
  assert lib.sg2im_conv_tc(_p(x), Ci, N, H, W, Ci, _p(wt), None, 3, 3, 1, H, W, Co, 0, 0.0, _p(y), Co, 0,
                           None, 0, None) == 0, lib.emul_last_error()
  assert rel_err(y, F.conv2d(x.permute(0, 3, 1, 2), w, None, padding=1).permute(0, 2, 3, 1)) < 2e-6
  # CTA pair: one pair runs 6 work units
  env(SG2IM_HALO_PAIR=1, SG2IM_EMUL_SMS=2, **sched)
  N, H, W, Ci, Co = 6, 32, 16, 64, 128
  x = _tf32(torch.randn(N, H, W, Ci, generator=g))
  w = _tf32(torch.randn(Co, Ci, 3, 3, generator=g) * 0.1)
  wt = w.permute(2, 3, 0, 1).reshape(9, Co, Ci).contiguous()
  y = torch.empty(N, H, W, Co)
  assert lib.sg2im_conv_tc(_p(x), Ci, N, H, W, Ci, _p(wt), None, 3, 3, 1, H, W, Co, 0, 0.0, _p(y), Co, 0,
                           None, 0, None) == 0, lib.emul_last_error()
  assert rel_err(y, F.conv2d(x.permute(0, 3, 1, 2), w, None, padding=1).permute(0, 2, 3, 1)) < 2e-6
