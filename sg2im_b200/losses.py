"""GAN losses with the reference's surface (sg2im/losses.py).  Scalar reductions over discriminator
scores.  The 'gan' losses — bce_loss against all-ones / all-zeros — run as ONE fused kernel forward
and one backward on CUDA tensors (ops.BCELogitsMean: the reference's ~20 elementwise ATen ops per
evaluation, six evaluations per step, were the largest group of tiny launches left); everything else,
and CPU tensors, use the reference's torch composition."""
import torch


def get_gan_losses(gan_type):
  """sg2im/losses.py:21-36."""
  if gan_type == 'gan':
    return gan_g_loss, gan_d_loss
  elif gan_type == 'wgan':
    return wgan_g_loss, wgan_d_loss
  elif gan_type == 'lsgan':
    return lsgan_g_loss, lsgan_d_loss
  else:
    raise ValueError('Unrecognized GAN type "%s"' % gan_type)


def bce_loss(input, target):
  """Numerically stable BCE-with-logits, mean reduced (sg2im/losses.py:39-57)."""
  neg_abs = -input.abs()
  loss = input.clamp(min=0) - input * target + (1 + neg_abs.exp()).log()
  return loss.mean()


def _flat(x):
  return x.reshape(-1) if x.dim() > 1 else x


def _bce_const(s, target):
  """bce_loss(s, full_like(s, target)), target 0 or 1."""
  if s.is_cuda and s.dtype == torch.float32:
    from . import ops
    return ops.BCELogitsMean.apply(s, float(target))
  return bce_loss(s, torch.full_like(s, float(target)))


def gan_g_loss(scores_fake):
  return _bce_const(_flat(scores_fake), 1.0)


def gan_d_loss(scores_real, scores_fake):
  assert scores_real.size() == scores_fake.size()
  r, f = _flat(scores_real), _flat(scores_fake)
  return _bce_const(r, 1.0) + _bce_const(f, 0.0)


def wgan_g_loss(scores_fake):
  return -scores_fake.mean()


def wgan_d_loss(scores_real, scores_fake):
  return scores_fake.mean() - scores_real.mean()


def lsgan_g_loss(scores_fake):
  s = _flat(scores_fake)
  return torch.nn.functional.mse_loss(s.sigmoid(), torch.ones_like(s))


def lsgan_d_loss(scores_real, scores_fake):
  assert scores_real.size() == scores_fake.size()
  r, f = _flat(scores_real), _flat(scores_fake)
  mse = torch.nn.functional.mse_loss
  return mse(r.sigmoid(), torch.ones_like(r)) + mse(f.sigmoid(), torch.zeros_like(f))
