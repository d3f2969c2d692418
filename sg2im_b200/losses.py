"""GAN losses with the reference's surface (sg2im/losses.py).  Scalar
reductions over discriminator scores: these stay PyTorch."""
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


def gan_g_loss(scores_fake):
  s = _flat(scores_fake)
  return bce_loss(s, torch.ones_like(s))


def gan_d_loss(scores_real, scores_fake):
  assert scores_real.size() == scores_fake.size()
  r, f = _flat(scores_real), _flat(scores_fake)
  return bce_loss(r, torch.ones_like(r)) + bce_loss(f, torch.zeros_like(f))


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
