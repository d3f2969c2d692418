"""The non-tensor-core kernel SOURCES executed on the CPU.

`tests/emul/cuda_emul.h` is a host-side SIMT emulation of the CUDA builtins (one OS
thread per CUDA thread of a block, real barriers for __syncthreads / shuffles /
ballots, atomics under a lock).  With -DSG2IM_EMUL the very files nvcc compiles
(`csrc/layout.cu`, `norm_act_v2.cu`, `adam.cu`, `deprocess.cu` — kernels AND their host
launchers: grid sizing, tiling, dispatch on SG2IM_*_V2) build with g++ and run here
against the oracle / torch.  This pins index arithmetic, tiling, masking and reduction
logic of kernels that have not run on hardware yet, and cross-checks
the emulator itself on the hardware-validated first-generation layout kernels.
It does not model the tensor-core / TMA kernels, memory ordering or timing.
"""
import ctypes
import os
import shutil
import subprocess

import pytest
import torch
import torch.nn.functional as F

from conftest import ROOT, load_golden, rel_err

pytestmark = pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++ (C++20)')

_i64, _f32, _int, _ptr = ctypes.c_int64, ctypes.c_float, ctypes.c_int, ctypes.c_void_p


@pytest.fixture(scope='module')
def lib(tmp_path_factory):
  out = tmp_path_factory.mktemp('emul') / 'libemul.so'
  src = sorted(str(p) for p in (__import__('pathlib').Path(ROOT) / 'tests' / 'emul').glob('emul_*.cpp'))
  cmd = ['g++', '-std=c++20', '-O1', '-pthread', '-shared', '-fPIC', '-Wno-psabi', '-U_FORTIFY_SOURCE', '-DSG2IM_EMUL',
         '-I', os.path.join(ROOT, 'tests', 'emul'), '-I', os.path.join(ROOT, 'include'),
         '-I', os.path.join(ROOT, 'sg2im_b200', 'csrc')] + src + ['-o', str(out)]
  # e.g. SG2IM_EMUL_CXXFLAGS='-g -fsanitize=address,alignment,bounds' with LD_PRELOAD=libasan.so
  # (memcheck of the kernel sources) or '-g -fsanitize=thread' with LD_PRELOAD=libtsan.so (racecheck)
  cmd[1:1] = os.environ.get('SG2IM_EMUL_CXXFLAGS', '').split()
  subprocess.check_call(cmd)
  L = ctypes.CDLL(str(out))
  L.sg2im_layout_fwd.argtypes = [_ptr, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _int,
                                 _ptr, _i64, _i64, _i64, _i64, _i64, _ptr, _i64, _int, _ptr]
  L.sg2im_layout_bwd.argtypes = [_ptr, _i64, _ptr, _ptr, _ptr, _i64, _ptr, _i64, _i64, _i64, _i64, _i64,
                                 _int, _ptr, _ptr, _ptr]
  L.emul_bn_bwd_reduce_v2.argtypes = [_ptr, _i64, _i64, _ptr, _i64, _i64, _i64, _i64, _ptr, _ptr, _ptr,
                                      _f32, _int, _ptr]
  L.emul_bn_bwd_apply_v2.argtypes = [_ptr, _i64, _i64, _ptr, _i64, _i64, _i64, _i64, _ptr, _ptr, _ptr,
                                     _f32, _int, _ptr, _ptr]
  L.emul_scale_act_fwd_v2.argtypes = [_ptr, _i64, _i64, _i64, _i64, _ptr, _ptr, _f32, _int, _ptr, _i64,
                                      _i64, _int]
  L.emul_colsum_small.argtypes = [_ptr, _i64, _i64, _ptr]
  L.sg2im_adam_flat.argtypes = [_ptr, _ptr, _ptr, _ptr, _i64, _f32, _f32, _f32, _f32, _f32, _ptr, _ptr,
                                _ptr, _f32, _ptr]
  L.sg2im_deprocess.argtypes = [_ptr, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr, _int,
                                _ptr, _ptr, _i64, _i64, _i64, _i64, _ptr]
  L.emul_last_error.restype = ctypes.c_char_p
  L.emul_blocks_run.restype = ctypes.c_ulonglong
  return L


def _p(t):
  return None if t is None else t.data_ptr()


def _ok(lib, rc):
  assert rc == 0, lib.emul_last_error()


@pytest.fixture
def v2_env():
  keys = ('SG2IM_LAYOUT_V2',)
  yield lambda on: [os.environ.__setitem__(k, '1' if on else '0') for k in keys]
  for k in keys:
    os.environ.pop(k, None)


def _scene(O, N, D, M, seed, crowded=False):
  g = torch.Generator().manual_seed(seed)
  vecs = torch.randn(O, D, generator=g)
  xy = torch.rand(O, 2, generator=g) * 0.6
  boxes = torch.cat([xy, xy + torch.rand(O, 2, generator=g) * 0.35 + 0.1], 1)
  boxes[-1] = torch.tensor([0., 0., 1., 1.])
  if crowded:                                             # > 16 objects in image 0: several passes
    o2i = torch.cat([torch.zeros(O - 2, dtype=torch.int64), torch.full((2,), N - 1, dtype=torch.int64)])
  else:
    o2i = torch.sort(torch.randint(0, N, (O,), generator=g)).values
  masks = torch.rand(O, M, M, generator=g) if M else None
  counts = torch.bincount(o2i, minlength=N)
  row_ptr = torch.cat([torch.zeros(1, dtype=torch.int64), counts.cumsum(0)]).to(torch.int32)
  entries = (torch.arange(O, dtype=torch.int32) * 2).contiguous()    # objects are grouped by image
  return vecs, boxes, masks, o2i, row_ptr, entries


LAYOUT_CASES = [  # O, N, D, M, H, W, noise channels, crowded
    (6, 2, 8, 5, 8, 16, 0, False), (9, 3, 16, 8, 12, 40, 5, False), (20, 2, 8, 4, 6, 8, 3, True),
    (5, 2, 12, 0, 8, 8, 0, False), (4, 1, 8, 20, 8, 36, 33, False),
    (5, 2, 6, 4, 6, 10, 2, False)]            # D % 4 != 0: the scalar fallback kernel


@pytest.mark.parametrize('case', LAYOUT_CASES)
@pytest.mark.parametrize('gen', [1, 2])
def test_layout_forward_sources_on_cpu(lib, v2_env, case, gen):
  """csrc/layout.cu forward: first generation (validated on the B200 — here it validates the
  emulator) and second generation (band-resident, not yet run on hardware) vs the oracle."""
  from oracle import sg2im_oracle as orc
  O, N, D, M, H, W, nc, crowded = case
  vecs, boxes, masks, o2i, row_ptr, entries = _scene(O, N, D, M, seed=O + W, crowded=crowded)
  noise = torch.randn(N, nc, H, W, generator=torch.Generator().manual_seed(1)) if nc else None
  ref = (orc.masks_to_layout(vecs, boxes, masks, o2i, H, W, N) if M
         else orc.boxes_to_layout(vecs, boxes, o2i, H, W, N))
  if nc:
    ref = torch.cat([ref, noise], dim=1)
  ref = ref.permute(0, 2, 3, 1)
  # trailing channels of a wider buffer stay untouched; pixel stride a multiple of 4 floats so
  # that D % 4 == 0 cases take the float4 kernels
  extra = 4 + (-(D + nc)) % 4
  out = torch.full((N, H, W, D + nc + extra), 7.0)
  ns = noise.stride() if nc else (0, 0, 0, 0)
  v2_env(gen == 2)
  b0 = lib.emul_blocks_run()
  _ok(lib, lib.sg2im_layout_fwd(_p(vecs), _p(boxes), _p(masks), M, _p(row_ptr), _p(entries), N, O, D,
                                H, W, 0, _p(noise), nc, ns[0], ns[1], ns[2], ns[3], _p(out),
                                out.size(3), 0, None))
  if D % 4 == 0:                                          # the generation asked for really ran
    segs = -(-W // 32)
    assert lib.emul_blocks_run() - b0 == (N * -(-H // 4) if gen == 2 else N * H * segs)
  assert rel_err(out[..., :D + nc], ref) < 1e-5
  assert bool((out[..., D + nc:] == 7.0).all())
  # TF32-rounded variant: every value has its 13 low mantissa bits clear and is within 2^-11 relative
  out_r = torch.zeros(N, H, W, D + nc + extra)
  _ok(lib, lib.sg2im_layout_fwd(_p(vecs), _p(boxes), _p(masks), M, _p(row_ptr), _p(entries), N, O, D,
                                H, W, 0, _p(noise), nc, ns[0], ns[1], ns[2], ns[3], _p(out_r),
                                out_r.size(3), 1, None))
  out_r = out_r[..., :D + nc].contiguous()
  assert int((out_r.view(torch.int32) & 0x1fff).abs().max()) == 0
  assert float((out_r - ref).abs().max()) <= float(ref.abs().max()) * 2.0 ** -11 + 1e-6


@pytest.mark.parametrize('case', LAYOUT_CASES[:4])
@pytest.mark.parametrize('gen', [1, 2])
def test_layout_backward_sources_on_cpu(lib, v2_env, case, gen):
  from oracle import sg2im_oracle as orc
  O, N, D, M, H, W, nc, crowded = case
  vecs, boxes, masks, o2i, _, _ = _scene(O, N, D, M, seed=O + W, crowded=crowded)
  vr = vecs.clone().requires_grad_(True)
  mr = masks.clone().requires_grad_(True) if M else None
  ref = (orc.masks_to_layout(vr, boxes, mr, o2i, H, W, N) if M
         else orc.boxes_to_layout(vr, boxes, o2i, H, W, N))
  gy = torch.randn(N, H, W, D + 4, generator=torch.Generator().manual_seed(2))   # wider buffer: dcs > D
  ref.backward(gy[..., :D].permute(0, 3, 1, 2))
  dvecs = torch.zeros(O, D)
  dmasks = torch.zeros(O, M, M) if M else None
  v2_env(gen == 2)
  b0 = lib.emul_blocks_run()
  _ok(lib, lib.sg2im_layout_bwd(_p(gy), gy.size(3), _p(vecs), _p(boxes), _p(masks), M, _p(o2i), N, O, D,
                                H, W, 0, _p(dvecs), _p(dmasks), None))
  assert lib.emul_blocks_run() - b0 == O * -(-H // 8) * (-(-W // 32) if gen == 2 else 1)
  assert rel_err(dvecs, vr.grad) < 1e-4
  if M:
    assert rel_err(dmasks, mr.grad) < 1e-4


def _bn_case(N, H, W, C, up, extra, seed):
  g = torch.Generator().manual_seed(seed)
  x = torch.randn(N, H, W, C, generator=g)
  gamma = torch.linspace(0.5, 1.5, C)
  beta = torch.linspace(-0.2, 0.3, C)
  gy = torch.randn(N, H * up, W * up, C + extra, generator=g)
  mean = x.reshape(-1, C).mean(0)
  var = x.reshape(-1, C).var(0, unbiased=False)
  invstd = torch.rsqrt(var + 1e-5)
  scale = (gamma * invstd).contiguous()
  shift = (beta - mean * gamma * invstd).contiguous()
  save = torch.cat([mean, invstd]).contiguous()
  return x, gamma, beta, gy, scale, shift, save


BN_CASES = [(2, 8, 8, 16, 1, 0), (2, 4, 6, 12, 2, 8), (1, 16, 16, 64, 2, 32), (3, 5, 7, 4, 1, 4),
            (2, 8, 8, 132, 1, 0)]


@pytest.mark.parametrize('N,H,W,C,up,extra', BN_CASES)
def test_bn_backward_v2_source_on_cpu(lib, N, H, W, C, up, extra):
  """csrc/norm_act_v2.cu reduce + apply (and their launch geometry) vs torch autograd of
  BatchNorm(train) -> LeakyReLU -> nearest x`up` -> channel slice of a wider buffer."""
  x, gamma, beta, gy, scale, shift, save = _bn_case(N, H, W, C, up, extra, seed=C + up)
  xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
  gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
  y = F.leaky_relu(F.batch_norm(xr, None, None, gr, br, True, 0.1, 1e-5), 0.2)
  if up > 1:
    y = F.interpolate(y, scale_factor=up, mode='nearest')
  y.backward(gy[..., extra:].permute(0, 3, 1, 2))
  sums = torch.zeros(2 * C, dtype=torch.float64)
  _ok(lib, lib.emul_bn_bwd_reduce_v2(_p(gy), gy.size(3), extra, _p(x), N, H, W, C, _p(scale), _p(shift),
                                     _p(save), 0.2, up, _p(sums)))
  # sums = (d loss / d beta, d loss / d gamma)
  assert rel_err(sums[:C].float(), br.grad) < 1e-5
  assert rel_err(sums[C:].float(), gr.grad) < 1e-5
  dx = torch.full((N, H, W, C), float('nan'))
  _ok(lib, lib.emul_bn_bwd_apply_v2(_p(gy), gy.size(3), extra, _p(x), N, H, W, C, _p(scale), _p(shift),
                                    _p(save), 0.2, up, _p(sums), _p(dx)))
  assert bool(torch.isfinite(dx).all())                    # every element written
  assert rel_err(dx, xr.grad.permute(0, 2, 3, 1)) < 1e-4


@pytest.mark.parametrize('N,H,W,C,up,extra', BN_CASES)
@pytest.mark.parametrize('affine', [True, False])
def test_scale_act_forward_v2_source_on_cpu(lib, N, H, W, C, up, extra, affine):
  x, _, _, _, scale, shift, _ = _bn_case(N, H, W, C, up, extra, seed=C)
  pre = x * scale + shift if affine else x
  ref = F.leaky_relu(pre, 0.2)
  if up > 1:
    ref = ref.repeat_interleave(up, dim=1).repeat_interleave(up, dim=2)
  y = torch.full((N, H * up, W * up, C + extra), 7.0)
  _ok(lib, lib.emul_scale_act_fwd_v2(_p(x), N, H, W, C, _p(scale) if affine else None,
                                     _p(shift) if affine else None, 0.2, up, _p(y), y.size(3), extra, 0))
  assert rel_err(y[..., extra:], ref) < 1e-6
  assert bool((y[..., :extra] == 7.0).all())


@pytest.mark.parametrize('M,C', [(448, 100), (1, 5), (37, 179), (300, 32)])
def test_colsum_small_source_on_cpu(lib, M, C):
  x = torch.randn(M, C, generator=torch.Generator().manual_seed(M))
  out = torch.full((C,), float('nan'))
  _ok(lib, lib.emul_colsum_small(_p(x), M, C, _p(out)))
  assert torch.allclose(out, x.double().sum(0).float(), rtol=1e-6, atol=1e-5)


def test_adam_flat_source_on_cpu(lib):
  """csrc/adam.cu vs torch.optim.Adam: several steps, a length that is not a multiple of 4
  (scalar tail), weight decay, and the found_inf skip (no update, no step increment)."""
  for n, wd in ((1027, 0.0), (4096, 0.0), (515, 0.01)):
    g = torch.Generator().manual_seed(n)
    p0 = torch.randn(n, generator=g)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-2, weight_decay=wd)
    p, m, v = p0.clone(), torch.zeros(n), torch.zeros(n)
    step = torch.zeros(())
    for it in range(5):
      grad = torch.randn(n, generator=g)
      ref.grad = grad.clone()
      opt.step()
      shadow = torch.full((n,), float('nan'))
      _ok(lib, lib.sg2im_adam_flat(_p(p), _p(grad), _p(m), _p(v), n, 1e-2, 0.9, 0.999, 1e-8, wd,
                                   _p(step), None, _p(shadow), 1.0, None))
      assert torch.allclose(p, ref.detach(), rtol=1e-5, atol=1e-7), (n, it)
      # the RN-TF32 shadow: low 13 mantissa bits clear, within half a TF32 ulp of the master
      assert int((shadow.view(torch.int32) & 0x1fff).abs().max()) == 0
      assert bool(((shadow - p).abs() <= p.abs() * 2.0 ** -11 + 1e-30).all())
    assert float(step) == 5
    before = p.clone()
    inf = torch.ones(())
    _ok(lib, lib.sg2im_adam_flat(_p(p), _p(grad), _p(m), _p(v), n, 1e-2, 0.9, 0.999, 1e-8, wd,
                                 _p(step), _p(inf), None, 1.0, None))
    assert torch.equal(p, before) and float(step) == 5
    # gradient scale (1 / world of a SUM all-reduce): same update as Adam on the scaled gradient
    p2, m2, v2, s2 = p.clone(), m.clone(), v.clone(), step.clone()
    _ok(lib, lib.sg2im_adam_flat(_p(p), _p(grad), _p(m), _p(v), n, 1e-2, 0.9, 0.999, 1e-8, wd,
                                 _p(step), None, None, 0.25, None))
    g4 = (grad * 0.25).contiguous()
    _ok(lib, lib.sg2im_adam_flat(_p(p2), _p(g4), _p(m2), _p(v2), n, 1e-2, 0.9, 0.999, 1e-8, wd,
                                 _p(s2), None, None, 1.0, None))
    assert torch.equal(p, p2) and torch.equal(m, m2) and torch.equal(v, v2)


def test_deprocess_source_on_cpu_bit_exact(lib):
  """csrc/deprocess.cu against the bytes of the reference's imagenet_deprocess_batch."""
  g = load_golden('aux.pt')['deprocess']
  x = g['imgs']
  N, C, H, W = x.shape
  inv_std = torch.tensor([1.0 / s for s in (0.229, 0.224, 0.225)], dtype=torch.float32)
  neg_mean = torch.tensor([-m for m in (0.485, 0.456, 0.406)], dtype=torch.float32)
  views = {'nchw': x, 'nhwc-backed': x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)}
  for name, src in views.items():
    for rescale, want in ((1, g['rescaled']), (0, g['plain'])):
      out = torch.zeros(N, C, H, W, dtype=torch.uint8)
      mm = torch.zeros(2 * N, dtype=torch.int32)
      sn, sc, sh, sw = src.stride()
      on, oc, oh, ow = out.stride()
      _ok(lib, lib.sg2im_deprocess(_p(src), sn, sc, sh, sw, N, C, H, W, _p(inv_std), _p(neg_mean),
                                   rescale, _p(mm), _p(out), on, oc, oh, ow, None))
      assert torch.equal(out, want), (name, rescale)
  # channels-last output
  out = torch.zeros(N, H, W, C, dtype=torch.uint8)
  mm = torch.zeros(2 * N, dtype=torch.int32)
  sn, sc, sh, sw = x.stride()
  on, oh, ow, oc = out.stride()
  _ok(lib, lib.sg2im_deprocess(_p(x), sn, sc, sh, sw, N, C, H, W, _p(inv_std), _p(neg_mean), 1,
                               _p(mm), _p(out), on, oc, oh, ow, None))
  assert torch.equal(out.permute(0, 3, 1, 2), g['rescaled'])


# ---------------------------------------------------------------------------
# First-generation kernels (graph, crop, pack, norm/act, exact-fp32 convolutions): these HAVE
# run on the B200 (tests/test_gpu_*.py); executing their sources here puts the same parity
# checks into the CPU suite and validates the emulator against hardware-proven code.
# ---------------------------------------------------------------------------

@pytest.fixture(scope='module')
def api(lib):
  from sg2im_b200._lib import SIGNATURES
  for name, sig in SIGNATURES.items():
    if hasattr(lib, name):
      getattr(lib, name).argtypes = sig
  return lib


@pytest.mark.parametrize('T,O,H,D,avg', [(10, 7, 16, 8, True), (448, 320, 32, 8, True),
                                         (33, 5, 6, 3, False), (64, 9, 12, 4, True)])
def test_graph_kernels_source_on_cpu_bit_exact(api, T, O, H, D, avg):
  """csrc/graph.cu: CSR build (ballot / popc compaction) + ordered segment sum must equal the
  CPU scatter_add of the reference's pooling BIT FOR BIT (sg2im/graph.py:85-114)."""
  from oracle import sg2im_oracle as orc
  g = torch.Generator().manual_seed(T)
  edges = torch.randint(0, O, (T, 2), generator=g)
  edges[:, 0] = torch.where(edges[:, 0] == O - 1, torch.zeros(()).long(), edges[:, 0])   # leave a row unused
  new_t = torch.randn(T, 2 * H + D, generator=g)
  row_ptr = torch.zeros(O + 1, dtype=torch.int32)
  entries = torch.zeros(2 * T, dtype=torch.int32)
  _ok(api, api.sg2im_csr_build(_p(edges), T, 2, 2, O, _p(row_ptr), _p(entries), None))
  counts = torch.bincount(edges.reshape(-1), minlength=O)
  assert torch.equal(row_ptr[1:].long() - row_ptr[:-1].long(), counts)
  pooled = torch.full((O, H), float('nan'))
  _ok(api, api.sg2im_segment_sum(_p(new_t), new_t.size(1), 0, H + D, H, _p(row_ptr), _p(entries), O,
                                 int(avg), _p(pooled), None))
  ref = orc.graph_pool(new_t, edges, O, H, D, 'avg' if avg else 'sum')
  assert torch.equal(pooled, ref)
  # forward gather cat([obj[s], pred, obj[o]])
  obj, pred = torch.randn(O, H, generator=g), torch.randn(T, D, generator=g)
  out = torch.empty(T, 2 * H + D)
  _ok(api, api.sg2im_triple_gather(_p(obj), _p(pred), _p(edges), T, H, D, None, _p(out), None))
  assert torch.equal(out, torch.cat([obj[edges[:, 0]], pred, obj[edges[:, 1]]], dim=1))


def test_crop_kernels_source_on_cpu(api):
  from oracle import sg2im_oracle as orc
  g = load_golden('crop.pt')
  feats, boxes, idx = g['feats'], g['boxes'], g['bbox_to_feats']
  N, C, H, W = feats.shape
  B, HH = boxes.size(0), 8
  nhwc = feats.permute(0, 2, 3, 1)                       # strided view of the NCHW fixture
  sn, sh, sw, sc = nhwc.stride()
  out = torch.empty(B, HH, HH, C)
  _ok(api, api.sg2im_crop_fwd(_p(feats), sn, sh, sw, sc, N, H, W, C, _p(boxes), _p(idx), B, HH, HH, 0,
                              _p(out), None))
  assert rel_err(out.permute(0, 3, 1, 2), g['crops']) < 1e-5
  fr = feats.clone().requires_grad_(True)
  ref = orc.crop_bbox_batch(fr, boxes, idx, HH)
  gy = torch.randn(B, HH, HH, C, generator=torch.Generator().manual_seed(1))
  ref.backward(gy.permute(0, 3, 1, 2))
  dfeats = torch.zeros(N, H, W, C)
  _ok(api, api.sg2im_crop_bwd(_p(gy), _p(boxes), _p(idx), N, H, W, C, B, HH, HH, 0, _p(dfeats), None))
  assert rel_err(dfeats.permute(0, 3, 1, 2), fr.grad) < 1e-5


@pytest.mark.parametrize('Co,Ci,cu,K', [(40, 36, 36, 3), (33, 64, 32, 3), (8, 12, 12, 1), (16, 8, 8, 2),
                                        (5, 7, 7, 5)])
def test_pack_kernels_source_on_cpu(api, Co, Ci, cu, K):
  T = K * K
  w = torch.randn(Co, Ci, K, K, generator=torch.Generator().manual_seed(Co))
  fwd, dgr = torch.empty(T, Co, cu), torch.empty(T, cu, Co)
  _ok(api, api.sg2im_pack_weights(_p(w), Co, Ci, cu, T, _p(fwd), _p(dgr), 0, None))
  ws = w[:, :cu].reshape(Co, cu, T)
  assert torch.equal(fwd, ws.permute(2, 0, 1))
  assert torch.equal(dgr, ws.flip(2).permute(2, 1, 0))
  # round-to-nearest TF32 variant, one layout at a time (what ops._pack does)
  fr = torch.empty(T, Co, cu)
  _ok(api, api.sg2im_pack_weights(_p(w), Co, Ci, cu, T, _p(fr), None, 1, None))
  assert int((fr.view(torch.int32) & 0x1fff).abs().max()) == 0
  assert float((fr - fwd).abs().max()) <= float(fwd.abs().max()) * 2.0 ** -11
  # unpack: wgrad layout -> OIHW, with and without accumulation, partial channel use
  dw = torch.randn(T, cu, Co, generator=torch.Generator().manual_seed(1))
  grad = torch.zeros(Co, Ci, K, K)
  _ok(api, api.sg2im_unpack_wgrad(_p(dw), Co, Ci, cu, T, _p(grad), 0, None))
  want = torch.zeros(Co, Ci, T)
  want[:, :cu] = dw.permute(2, 1, 0)
  assert torch.equal(grad.view(Co, Ci, T), want)
  _ok(api, api.sg2im_unpack_wgrad(_p(dw), Co, Ci, cu, T, _p(grad), 1, None))
  assert torch.equal(grad.view(Co, Ci, T), 2 * want)


@pytest.mark.parametrize('N,H,W,C,up,extra', [(2, 8, 8, 16, 1, 0), (2, 4, 6, 12, 2, 8), (3, 5, 7, 5, 1, 3)])
def test_norm_act_kernels_source_on_cpu(api, N, H, W, C, up, extra):
  """csrc/norm_act.cu: batch statistics, finalize (running stats), the fused
  BN+LeakyReLU+upsample+slice forward and its backward (generic kernels)."""
  g = torch.Generator().manual_seed(C)
  x = torch.randn(N, H, W, C, generator=g)
  gamma, beta = torch.linspace(0.5, 1.5, C), torch.linspace(-0.2, 0.3, C)
  M = N * H * W
  sums = torch.zeros(2 * C, dtype=torch.float64)
  _ok(api, api.sg2im_bn_stats(_p(x), M, C, _p(sums), None))
  assert torch.allclose(sums[:C], x.double().reshape(M, C).sum(0), rtol=1e-6, atol=1e-6)
  rm, rv = torch.zeros(C), torch.ones(C)
  scale, shift, save = torch.empty(C), torch.empty(C), torch.empty(2 * C)
  nbt = torch.tensor(4, dtype=torch.int64)
  _ok(api, api.sg2im_bn_finalize(_p(sums), M, 1, C, _p(gamma), _p(beta), 1e-5, 0.1, 1, _p(rm), _p(rv),
                                 _p(scale), _p(shift), _p(save), _p(nbt), None))
  assert int(nbt) == 5                                    # nn.BatchNorm's counter advances in-kernel
  bn = torch.nn.BatchNorm2d(C)
  with torch.no_grad():
    bn.weight.copy_(gamma); bn.bias.copy_(beta)
  xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
  y = F.leaky_relu(bn(xr), 0.2)
  if up > 1:
    y = F.interpolate(y, scale_factor=up, mode='nearest')
  assert rel_err(rm, bn.running_mean) < 1e-5 and rel_err(rv, bn.running_var) < 1e-5
  out = torch.full((N, H * up, W * up, C + extra), 7.0)
  _ok(api, api.sg2im_scale_act_fwd(_p(x), N, H, W, C, _p(scale), _p(shift), 0.2, up, _p(out),
                                   out.size(3), extra, 0, None))
  assert rel_err(out[..., extra:], y.permute(0, 2, 3, 1)) < 1e-5
  assert bool((out[..., :extra] == 7.0).all())
  gy = torch.randn(N, H * up, W * up, C + extra, generator=g)
  y.backward(gy[..., extra:].permute(0, 3, 1, 2))
  bs = torch.zeros(2 * C, dtype=torch.float64)
  _ok(api, api.sg2im_scale_act_bwd_reduce(_p(gy), gy.size(3), extra, _p(x), N, H, W, C, _p(scale),
                                          _p(shift), _p(save), 0.2, up, _p(bs), None))
  dx, dg, db = torch.empty(N, H, W, C), torch.empty(C), torch.empty(C)
  _ok(api, api.sg2im_scale_act_bwd_apply(_p(gy), gy.size(3), extra, _p(x), N, H, W, C, _p(scale),
                                         _p(shift), _p(save), 0.2, up, 1, _p(bs), _p(dx), _p(dg), _p(db),
                                         None))
  assert rel_err(dx, xr.grad.permute(0, 2, 3, 1)) < 1e-4
  assert rel_err(dg, bn.weight.grad) < 1e-5 and rel_err(db, bn.bias.grad) < 1e-5


def test_pool_s2d_act_colsum_sources_on_cpu(api):
  g = torch.Generator().manual_seed(4)
  x = torch.randn(2, 8, 6, 12, generator=g)
  out = torch.full((2, 4, 3, 16), 7.0)
  _ok(api, api.sg2im_avgpool2_fwd(_p(x), 12, 4, 2, 8, 6, 8, _p(out), 16, 4, None))   # channels 4..11 -> 4..11
  ref = F.avg_pool2d(x[..., 4:12].permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
  assert rel_err(out[..., 4:12], ref) < 1e-6 and bool((out[..., :4] == 7.0).all())
  dfine = torch.ones(2, 8, 6, 12)
  dco = torch.randn(2, 4, 3, 16, generator=g)
  _ok(api, api.sg2im_avgpool2_bwd(_p(dco), 16, 4, 2, 8, 6, 8, _p(dfine), 12, 4, 1, None))
  up = (dco[..., 4:12] * 0.25).repeat_interleave(2, 1).repeat_interleave(2, 2)
  assert rel_err(dfine[..., 4:12], 1 + up) < 1e-6 and bool((dfine[..., :4] == 1).all())
  # space-to-depth (odd sizes are zero-padded) and back
  xs = torch.randn(2, 7, 5, 4, generator=g)
  s2d = torch.empty(2, 4, 3, 16)
  sn, sh, sw, sc = xs.stride()
  _ok(api, api.sg2im_s2d_fwd(_p(xs), sn, sh, sw, sc, 2, 7, 5, 4, _p(s2d), None))
  pad = F.pad(xs, (0, 0, 0, 1, 0, 1))
  want = pad.view(2, 4, 2, 3, 2, 4).permute(0, 1, 3, 2, 4, 5).reshape(2, 4, 3, 16)
  assert torch.equal(s2d, want)
  back = torch.empty(2, 7, 5, 4)
  _ok(api, api.sg2im_s2d_bwd(_p(s2d), 2, 7, 5, 4, _p(back), None))
  assert torch.equal(back, xs)
  # activation backward and bias-gradient column sum
  yv, dy = torch.randn(50, 7, generator=g), torch.randn(50, 7, generator=g)
  dxv = torch.empty(50, 7)
  _ok(api, api.sg2im_act_bwd(_p(dy), _p(yv), 0.2, 350, _p(dxv), None))
  assert torch.equal(dxv, torch.where(yv > 0, dy, dy * 0.2))
  cs, scratch = torch.empty(7), torch.empty(7, dtype=torch.float64)
  _ok(api, api.sg2im_colsum(_p(dy), 50, 7, _p(cs), _p(scratch), None))
  assert torch.allclose(cs, dy.double().sum(0).float(), rtol=1e-6, atol=1e-6)


CONV_CASES = [  # N, H, W, Ci, Co, K, S, P, act
    (2, 6, 6, 5, 7, 3, 1, 1, 1), (1, 9, 8, 4, 6, 4, 2, 0, 0), (3, 1, 1, 20, 9, 1, 1, 0, 1),
    (1, 10, 10, 3, 130, 3, 1, 1, 0), (2, 5, 5, 8, 3, 1, 1, 0, 0),
    (37, 1, 1, 260, 3, 1, 1, 0, 1), (9, 1, 1, 1024, 1, 1, 1, 0, 0)]     # warp-per-row skinny kernel (heads)


@pytest.mark.parametrize('N,H,W,Ci,Co,K,S,P,act', CONV_CASES)
def test_exact_fp32_conv_sources_on_cpu(api, N, H, W, Ci, Co, K, S, P, act):
  """csrc/conv_simt.cu (forward, data gradient, weight gradient, skinny variants) vs torch."""
  g = torch.Generator().manual_seed(Ci * Co)
  x = torch.randn(N, Ci, H, W, generator=g)             # NCHW storage, addressed through strides
  w = torch.randn(Co, Ci, K, K, generator=g) * 0.2
  b = torch.randn(Co, generator=g)
  xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
  pre = F.conv2d(xr, wr, b, stride=S, padding=P)
  ref = F.leaky_relu(pre, 0.2) if act else pre
  Ho, Wo = ref.shape[2], ref.shape[3]
  nhwc = x.permute(0, 2, 3, 1)
  sn, sh, sw, sc = nhwc.stride()
  if H == 1 and W == 1:
    sh = sw = sn                                        # rows of a matrix, as ops.linear passes them
  wf = w.permute(2, 3, 1, 0).reshape(K * K * Ci, Co).contiguous()
  y = torch.full((N, Ho, Wo, Co + 2), 7.0)
  _ok(api, api.sg2im_conv_igemm(0, _p(x), sn, sh, sw, sc, N, H, W, Ci, _p(wf), _p(b), K, K, S, P, Ho, Wo,
                                Co, act, 0.2, _p(y), Co + 2, 2, None))
  assert rel_err(y[..., 2:], ref.permute(0, 2, 3, 1)) < 1e-5
  assert bool((y[..., :2] == 7.0).all())
  gy = torch.randn(N, Ho, Wo, Co, generator=g)
  pre.backward(gy.permute(0, 3, 1, 2))
  wd = w.permute(2, 3, 0, 1).reshape(K * K * Co, Ci).contiguous()
  dx = torch.empty(N, H, W, Ci)
  gs = gy.stride()
  _ok(api, api.sg2im_conv_igemm(1, _p(gy), gs[0], gs[1], gs[2], gs[3], N, Ho, Wo, Co, _p(wd), None, K, K,
                                S, P, H, W, Ci, 0, 0.0, _p(dx), Ci, 0, None))
  assert rel_err(dx, xr.grad.permute(0, 2, 3, 1)) < 1e-5
  dw = torch.zeros(K * K * Ci, Co)
  _ok(api, api.sg2im_conv_wgrad(_p(x), sn, sh, sw, sc, N, H, W, Ci, _p(gy), K, K, S, P, Ho, Wo, Co,
                                _p(dw), None))
  assert rel_err(dw, wr.grad.permute(2, 3, 1, 0).reshape(K * K * Ci, Co)) < 1e-5


@pytest.mark.parametrize('M,C', [(50, 8), (1000, 64), (37, 132), (4096, 4)])
def test_act_backward_with_bias_gradient_source_on_cpu(api, M, C):
  """csrc/norm_act_v2.cu act_bwd_colsum: dx = dy * leaky'(y) and db += column sums in one pass."""
  g = torch.Generator().manual_seed(M + C)
  dy, y = torch.randn(M, C, generator=g), torch.randn(M, C, generator=g)
  dx = torch.full((M, C), float('nan'))
  db = torch.full((C,), 2.0)                              # accumulated INTO
  _ok(api, api.sg2im_act_bwd_colsum(_p(dy), _p(y), 0.2, M, C, _p(dx), _p(db), None))
  want = torch.where(y > 0, dy, dy * 0.2)
  assert torch.equal(dx, want)
  assert torch.allclose(db - 2.0, want.double().sum(0).float(), rtol=1e-5, atol=1e-4)


# ---- csrc/pool.cu, csrc/split.cu: direct C-ABI calls (so tools/emul_sanitize.sh covers them)

@pytest.mark.parametrize('N,H,W,C,f,mode', [(2, 8, 8, 8, 2, 1), (3, 13, 9, 6, 3, 1), (2, 7, 10, 5, 2, 0),
                                            (1, 4, 4, 4, 4, 0), (1, 9, 9, 12, 3, 1)])
def test_pool2d_kernels(lib, N, H, W, C, f, mode):
  lib.sg2im_pool2d_fwd.argtypes = [_ptr, _i64, _i64, _i64, _i64, _int, _int, _ptr, _ptr]
  lib.sg2im_pool2d_bwd.argtypes = [_ptr, _ptr, _i64, _i64, _i64, _i64, _int, _int, _ptr, _ptr]
  g = torch.Generator().manual_seed(H * 31 + C)
  x = (torch.randn(N, C, H, W, generator=g) * 2).round() / 2          # exact ties inside windows
  xr = x.clone().requires_grad_(True)
  want = (F.avg_pool2d if mode == 0 else F.max_pool2d)(xr, f, f)
  dy = torch.randn(want.shape, generator=g)
  want.backward(dy)
  h = x.permute(0, 2, 3, 1).contiguous()
  y = torch.full((N, H // f, W // f, C), float('nan'))
  _ok(lib, lib.sg2im_pool2d_fwd(_p(h), N, H, W, C, f, mode, _p(y), None))
  dx = torch.zeros(N, H, W, C)
  dyh = dy.permute(0, 2, 3, 1).contiguous()            # keep alive across the call
  _ok(lib, lib.sg2im_pool2d_bwd(_p(dyh), _p(h), N, H, W, C, f, mode, _p(dx), None))
  if mode == 1:
    assert torch.equal(y.permute(0, 3, 1, 2), want.detach())
    assert torch.equal(dx.permute(0, 3, 1, 2), xr.grad)
  else:
    assert rel_err(y.permute(0, 3, 1, 2), want.detach()) < 1e-6
    assert rel_err(dx.permute(0, 3, 1, 2), xr.grad) < 1e-6
  assert lib.sg2im_pool2d_fwd(_p(h), N, H, W, C, H + 1, mode, _p(y), None) != 0     # empty output refused


@pytest.mark.parametrize('with_masks,align', [(True, 0), (False, 0), (True, 1)])
def test_layout_bwd_boxes_kernel(lib, with_masks, align):
  """csrc/layout_boxes.cu vs torch's grid_sample grid gradient (direct C-ABI call)."""
  from oracle import sg2im_oracle as orc
  lib.sg2im_layout_bwd_boxes.argtypes = [_ptr, _i64, _ptr, _ptr, _ptr, _i64, _ptr, _i64, _i64, _i64,
                                         _i64, _i64, _int, _ptr, _ptr]
  g = torch.Generator().manual_seed(5 + align)
  O, N, D, M, H, W, CS = 6, 2, 20, 6, 14, 11, 27
  vecs = torch.randn(O, D, generator=g)
  xy = torch.rand(O, 2, generator=g) * 0.5
  boxes = torch.cat([xy, xy + 0.2 + 0.3 * torch.rand(O, 2, generator=g)], 1)
  masks = torch.rand(O, M, M, generator=g) if with_masks else None
  o2i = torch.tensor([0, 0, 0, 1, 1, 1])
  br = boxes.clone().requires_grad_(True)
  if with_masks:
    want = orc.masks_to_layout(vecs, br, masks, o2i, H, W, N, align_corners=bool(align))
  else:
    want = orc.boxes_to_layout(vecs, br, o2i, H, W, N, align_corners=bool(align))
  dout = torch.randn(N, H, W, CS, generator=g)                  # gradient slice of a wider buffer
  (want * dout[..., :D].permute(0, 3, 1, 2)).sum().backward()
  db = torch.full((O, 4), float('nan'))
  _ok(lib, lib.sg2im_layout_bwd_boxes(_p(dout), CS, _p(vecs), _p(boxes), _p(masks), M, _p(o2i), N, O, D,
                                     H, W, align, _p(db), None))
  assert rel_err(db, br.grad) < 2e-4, (db, br.grad)


@pytest.mark.parametrize('n,target', [(320, 1.0), (7, 0.0), (5000, 1.0), (1, 0.0)])
def test_fused_bce_with_logits_mean_source_on_cpu(api, n, target):
  """csrc/loss.cu vs the reference's composition (sg2im/losses.py:39-57) and its autograd."""
  from sg2im_b200.losses import bce_loss
  g = torch.Generator().manual_seed(n)
  x = (torch.randn(n, generator=g) * 4).requires_grad_(True)
  ref = bce_loss(x, torch.full_like(x, target))
  ref.backward(torch.tensor(0.37))
  out, scratch = torch.empty(()), torch.zeros(1, dtype=torch.float64)
  _ok(api, api.sg2im_bce_logits_mean_fwd(_p(x), n, target, _p(scratch), _p(out), None))
  assert abs(float(out) - float(ref)) <= 2e-6 * max(1.0, abs(float(ref)))
  gout, dx = torch.tensor(0.37), torch.empty(n)
  _ok(api, api.sg2im_bce_logits_mean_bwd(_p(x), n, target, _p(gout), _p(dx), None))
  assert rel_err(dx, x.grad) < 1e-5


@pytest.mark.parametrize('target', [0.0, 1.0])
def test_fused_bce_extreme_logits(api, target):
  """Saturated scores (a discriminator that has won): the reference's formula is stable there
  (max(x,0) - x t + log1p(exp(-|x|))), so is the kernel; gradients are exactly sigmoid(x) - t."""
  from sg2im_b200.losses import bce_loss
  x = torch.tensor([-1e4, -88.0, -30.0, -1e-8, 0.0, 1e-8, 30.0, 88.0, 1e4, 3.0e38, -3.0e38]).requires_grad_(True)
  n = x.numel()
  ref = bce_loss(x, torch.full_like(x, target))
  ref.backward()
  out, scratch = torch.empty(()), torch.zeros(1, dtype=torch.float64)
  _ok(api, api.sg2im_bce_logits_mean_fwd(_p(x), n, target, _p(scratch), _p(out), None))
  assert torch.isfinite(out) and abs(float(out) - float(ref)) <= 1e-6 * abs(float(ref))
  gout, dx = torch.tensor(1.0), torch.empty(n)
  _ok(api, api.sg2im_bce_logits_mean_bwd(_p(x), n, target, _p(gout), _p(dx), None))
  assert bool(torch.isfinite(dx).all())
  assert float((dx - x.grad).abs().max()) <= 1e-7

