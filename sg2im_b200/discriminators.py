"""Image and object discriminators — reference surface of
sg2im/discriminators.py (same constructor arguments, module tree and
state_dict keys) on the sm_100a conv / crop kernels."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .bilinear import crop_bbox_batch
from .layers import GlobalAvgPool, Flatten, get_activation, build_cnn, Conv2d, Linear


class PatchDiscriminator(nn.Module):
  """sg2im/discriminators.py:25-45.  As in the reference, ``classifier`` is
  constructed (it lives in the state_dict) but never applied by forward."""

  def __init__(self, arch, normalization='batch', activation='leakyrelu-0.2',
               padding='same', pooling='avg', input_size=(128, 128),
               layout_dim=0):
    super(PatchDiscriminator, self).__init__()
    input_dim = 3 + layout_dim
    arch = 'I%d,%s' % (input_dim, arch)
    cnn_kwargs = {
      'arch': arch,
      'normalization': normalization,
      'activation': activation,
      'pooling': pooling,
      'padding': padding,
    }
    self.cnn, output_dim = build_cnn(**cnn_kwargs)
    self.classifier = Conv2d(output_dim, 1, kernel_size=1, stride=1)

  def forward(self, x, layout=None):
    if layout is not None:
      x = torch.cat([x, layout], dim=1)
    return self.cnn(x)


class AcDiscriminator(nn.Module):
  """sg2im/discriminators.py:48-75."""

  def __init__(self, vocab, arch, normalization='none', activation='relu',
               padding='same', pooling='avg'):
    super(AcDiscriminator, self).__init__()
    self.vocab = vocab
    cnn_kwargs = {
      'arch': arch,
      'normalization': normalization,
      'activation': activation,
      'pooling': pooling,
      'padding': padding,
    }
    cnn, D = build_cnn(**cnn_kwargs)
    self.cnn = nn.Sequential(cnn, GlobalAvgPool(), Linear(D, 1024))
    num_objects = len(vocab['object_idx_to_name'])
    self.real_classifier = Linear(1024, 1)
    self.obj_classifier = Linear(1024, num_objects)

  def forward(self, x, y):
    if x.dim() == 3:
      x = x[:, None]
    vecs = self.cnn(x)
    real_scores = self.real_classifier(vecs)
    obj_scores = self.obj_classifier(vecs)
    ac_loss = F.cross_entropy(obj_scores, y)
    return real_scores, ac_loss


class AcCropDiscriminator(nn.Module):
  """sg2im/discriminators.py:78-90."""

  def __init__(self, vocab, arch, normalization='none', activation='relu',
               object_size=64, padding='same', pooling='avg'):
    super(AcCropDiscriminator, self).__init__()
    self.vocab = vocab
    self.discriminator = AcDiscriminator(vocab, arch, normalization,
                                         activation, padding, pooling)
    self.object_size = object_size

  def forward(self, imgs, objs, boxes, obj_to_img):
    crops = crop_bbox_batch(imgs, boxes, obj_to_img, self.object_size)
    real_scores, ac_loss = self.discriminator(crops, objs)
    return real_scores, ac_loss
