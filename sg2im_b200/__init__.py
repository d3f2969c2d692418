"""sg2im_b200 — B200-native (sm_100a) implementation of the sg2im generator +
discriminator training step behind the reference's Python surface.

  from sg2im_b200.model import Sg2ImModel
  from sg2im_b200.discriminators import PatchDiscriminator, AcCropDiscriminator
  from sg2im_b200.layers import build_mlp, build_cnn

The compute lives in libsg2im_b200.so (C-ABI: include/sg2im_b200.h), built by
``__graft_entry__.build()``; importing this package does not need a GPU, using
it does.
"""
__version__ = '0.1.0'
