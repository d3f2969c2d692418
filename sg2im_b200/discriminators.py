"""Discriminators of the sg2im training step on the sm_100a kernels.

Public surface (constructor arguments, forward signatures, module tree and
therefore ``state_dict`` keys) is the reference's ``sg2im/discriminators.py``:

  PatchDiscriminator(arch, ...)(x, layout=None)            -> feature map (N, C, h, w)
  AcDiscriminator(vocab, arch, ...)(x, y)                  -> (real_scores, ac_loss)
  AcCropDiscriminator(vocab, arch, ..., object_size)(imgs, objs, boxes, obj_to_img)

The conv stacks come from ``layers.build_cnn`` (tcgen05 convolutions via the
space-to-depth route for the 4x4 stride-2 layers, fused BatchNorm statistics /
LeakyReLU), the per-object crops from the single-launch crop kernel.
"""
import torch
from torch import nn
from torch.nn.functional import cross_entropy

from . import layers
from .bilinear import crop_bbox_batch

_HEAD_WIDTH = 1024          # width of the object discriminator's embedding (discriminators.py:62)


def _conv_stack(arch, normalization, activation, padding, pooling):
  """-> (FusedSequential, output channels); thin wrapper so both discriminators
  build their trunk the same way."""
  return layers.build_cnn(arch=arch, normalization=normalization, activation=activation,
                          padding=padding, pooling=pooling)


class PatchDiscriminator(nn.Module):
  """Image discriminator: a conv trunk whose whole output map is scored by the GAN
  loss.  Like the reference (discriminators.py:40-45) a 1x1 ``classifier`` is
  created — it is part of every checkpoint — but never applied."""

  def __init__(self, arch, normalization='batch', activation='leakyrelu-0.2',
               padding='same', pooling='avg', input_size=(128, 128), layout_dim=0):
    super().__init__()
    trunk_arch = 'I%d,%s' % (3 + layout_dim, arch)           # RGB (+ optional layout channels)
    self.cnn, width = _conv_stack(trunk_arch, normalization, activation, padding, pooling)
    self.classifier = layers.Conv2d(width, 1, kernel_size=1, stride=1)

  def forward(self, x, layout=None):
    inp = x if layout is None else torch.cat([x, layout], dim=1)
    return self.cnn(inp)


class AcDiscriminator(nn.Module):
  """Auxiliary-classifier discriminator on object crops: trunk -> global average
  pool -> Linear(D, 1024) -> {real/fake logit, object-class logits}."""

  def __init__(self, vocab, arch, normalization='none', activation='relu',
               padding='same', pooling='avg'):
    super().__init__()
    self.vocab = vocab
    trunk, width = _conv_stack(arch, normalization, activation, padding, pooling)
    # indices 0 / 2 of this Sequential carry parameters: keys cnn.0.*, cnn.2.*
    self.cnn = nn.Sequential(trunk, layers.GlobalAvgPool(), layers.Linear(width, _HEAD_WIDTH))
    self.real_classifier = layers.Linear(_HEAD_WIDTH, 1)
    self.obj_classifier = layers.Linear(_HEAD_WIDTH, len(vocab['object_idx_to_name']))

  def forward(self, x, y):
    crops = x.unsqueeze(1) if x.dim() == 3 else x            # single-channel crops without a C axis
    embedding = self.cnn(crops)
    class_logits = self.obj_classifier(embedding)
    return self.real_classifier(embedding), cross_entropy(class_logits, y)


class AcCropDiscriminator(nn.Module):
  """Crops every object's box out of its image (bilinear, ``object_size`` square)
  and scores the crops with an AcDiscriminator."""

  def __init__(self, vocab, arch, normalization='none', activation='relu',
               object_size=64, padding='same', pooling='avg'):
    super().__init__()
    self.vocab = vocab
    self.object_size = object_size
    self.discriminator = AcDiscriminator(vocab, arch, normalization, activation, padding,
                                         pooling)

  def forward(self, imgs, objs, boxes, obj_to_img):
    object_crops = crop_bbox_batch(imgs, boxes, obj_to_img, self.object_size)
    return self.discriminator(object_crops, objs)
