"""Image de-normalisation for sampling / validation — reference surface
``sg2im/data/utils.py:19-67`` (``imagenet_deprocess_batch``) on one device
kernel family (csrc/deprocess.cu): the float image never leaves HBM, the host
receives uint8."""
import torch

from . import _lib, ops

IMAGENET_MEAN = [0.485, 0.456, 0.406]
IMAGENET_STD = [0.229, 0.224, 0.225]

_consts = {}


def _norm_consts(device):
  key = str(device)
  if key not in _consts:
    # the two T.Normalize stages of utils.py:35-38: divide by 1/std, subtract -mean
    inv_std = torch.tensor([1.0 / s for s in IMAGENET_STD], dtype=torch.float32, device=device)
    neg_mean = torch.tensor([-m for m in IMAGENET_MEAN], dtype=torch.float32, device=device)
    _consts[key] = (inv_std, neg_mean)
  return _consts[key]


def imagenet_deprocess_batch(imgs, rescale=True, device_out=False, channels_last=False):
  """imgs: (N, 3, H, W) float32 CUDA tensor of ImageNet-normalised images (any
  strides: the generator's NCHW view of its NHWC buffer is read in place).
  Returns uint8 (N, 3, H, W) — or (N, H, W, 3) with ``channels_last`` — on the
  CPU like the reference, or on the device with ``device_out``.  Bytes are
  identical to the reference's (same fp32 operation order)."""
  ops._chk(imgs, name='imgs')
  if imgs.dim() != 4 or imgs.size(1) != 3:
    raise ValueError('imagenet_deprocess_batch: expected (N, 3, H, W), got %s' % (tuple(imgs.shape),))
  imgs = imgs.detach()
  N, C, H, W = imgs.shape
  inv_std, neg_mean = _norm_consts(imgs.device)
  if channels_last:
    out = torch.empty(N, H, W, C, dtype=torch.uint8, device=imgs.device)
    on, oh, ow, oc = out.stride()
  else:
    out = torch.empty(N, C, H, W, dtype=torch.uint8, device=imgs.device)
    on, oc, oh, ow = out.stride()
  scratch = torch.empty(max(2 * N, 1), dtype=torch.int32, device=imgs.device) if rescale else None
  sn, sc, sh, sw = imgs.stride()
  _lib.call('sg2im_deprocess', imgs.data_ptr(), sn, sc, sh, sw, N, C, H, W, inv_std.data_ptr(),
            neg_mean.data_ptr(), int(bool(rescale)), ops._p(scratch), out.data_ptr(), on, oc, oh, ow,
            ops._stream())
  ops._count(3 if rescale else 1)
  return out if device_out else out.cpu()
