"""Sg2ImModel — reference surface of sg2im/model.py (constructor, forward,
encode_scene_graphs, forward_json, module tree and state_dict keys) with the
hot path on libsg2im_b200.so."""
import torch
import torch.nn as nn

from . import ops
from . import layout as _layout
from .graph import GraphTripleConv, GraphTripleConvNet, GraphCSR
from .crn import RefinementNetwork
from .layers import build_mlp, Conv2d, BatchNorm2d, Linear, Upsample, ReLU, FusedSequential


class Sg2ImModel(nn.Module):
  """Scene graph -> image generator.  Constructor keywords, sub-module names and
  parameter shapes follow sg2im/model.py:30-92 (checkpoints are interchangeable);
  unknown keywords only warn, as there."""

  def __init__(self, vocab, image_size=(64, 64), embedding_dim=64,
               gconv_dim=128, gconv_hidden_dim=512,
               gconv_pooling='avg', gconv_num_layers=5,
               refinement_dims=(1024, 512, 256, 128, 64),
               normalization='batch', activation='leakyrelu-0.2',
               mask_size=None, mlp_normalization='none', layout_noise_dim=0,
               **kwargs):
    super().__init__()
    if kwargs:
      print('WARNING: Model got unexpected kwargs ', kwargs)

    self.vocab, self.image_size, self.layout_noise_dim = vocab, image_size, layout_noise_dim
    n_obj_classes = len(vocab['object_idx_to_name'])
    n_predicates = len(vocab['pred_idx_to_name'])

    # category / predicate embeddings (one spare object row, as in the reference)
    self.obj_embeddings = nn.Embedding(n_obj_classes + 1, embedding_dim)
    self.pred_embeddings = nn.Embedding(n_predicates, embedding_dim)

    # graph convolution: first layer maps embedding_dim -> gconv_dim, the rest keep gconv_dim
    shared = dict(hidden_dim=gconv_hidden_dim, pooling=gconv_pooling,
                  mlp_normalization=mlp_normalization)
    self.gconv = (Linear(embedding_dim, gconv_dim) if gconv_num_layers == 0 else
                  GraphTripleConv(input_dim=embedding_dim, output_dim=gconv_dim, **shared)
                  if gconv_num_layers > 0 else None)
    self.gconv_net = (GraphTripleConvNet(input_dim=gconv_dim, num_layers=gconv_num_layers - 1,
                                         **shared) if gconv_num_layers > 1 else None)

    # per-object heads: box regressor, optional mask generator; auxiliary relation classifier
    self.box_net = build_mlp([gconv_dim, gconv_hidden_dim, 4], batch_norm=mlp_normalization)
    self.mask_net = (self._build_mask_net(n_obj_classes, gconv_dim, mask_size)
                     if mask_size is not None and mask_size > 0 else None)
    self.rel_aux_net = build_mlp([2 * embedding_dim + 8, gconv_hidden_dim, n_predicates],
                                 batch_norm=mlp_normalization)

    # cascaded refinement network over the (layout ++ noise) canvas
    self.refinement_net = RefinementNetwork(
        dims=(gconv_dim + layout_noise_dim,) + tuple(refinement_dims),
        normalization=normalization, activation=activation)

  def _build_mask_net(self, num_objs, dim, mask_size):
    """1x1 -> mask_size x mask_size by repeated [nearest x2, BN, conv3x3, ReLU],
    then a 1x1 conv to one channel (sg2im/model.py:94-106)."""
    if mask_size & (mask_size - 1):
      raise ValueError('Mask size must be a power of 2')
    blocks = []
    for _ in range(mask_size.bit_length() - 1):
      blocks += [Upsample(scale_factor=2, mode='nearest'), BatchNorm2d(dim),
                 Conv2d(dim, dim, kernel_size=3, padding=1), ReLU()]
    blocks.append(Conv2d(dim, 1, kernel_size=1))
    return FusedSequential(*blocks)

  def forward(self, objs, triples, obj_to_img=None,
              boxes_gt=None, masks_gt=None, num_imgs=None, noise=None):
    """sg2im/model.py:108-171.  Extra keyword arguments (defaults preserve the
    reference behaviour): ``num_imgs`` avoids the device->host sync the
    reference needs to size the layout (layout.py:143) — pass imgs.size(0);
    ``noise`` (N, layout_noise_dim, H, W) overrides the torch.randn draw."""
    O, T = objs.size(0), triples.size(0)
    s, p, o = triples.chunk(3, dim=1)
    s, p, o = [x.squeeze(1) for x in [s, p, o]]
    edges = torch.stack([s, o], dim=1)

    if obj_to_img is None:
      obj_to_img = torch.zeros(O, dtype=objs.dtype, device=objs.device)
      if num_imgs is None:
        num_imgs = 1
    N = _layout._num_imgs(obj_to_img, num_imgs)

    obj_vecs = self.obj_embeddings(objs)
    obj_vecs_orig = obj_vecs
    pred_vecs = self.pred_embeddings(p)

    if isinstance(self.gconv, nn.Linear):
      obj_vecs = self.gconv(obj_vecs)
    else:
      csr = GraphCSR(edges, O)
      obj_vecs, pred_vecs = self.gconv(obj_vecs, pred_vecs, edges, csr)
      if self.gconv_net is not None:
        obj_vecs, pred_vecs = self.gconv_net(obj_vecs, pred_vecs, edges, csr)

    boxes_pred = self.box_net(obj_vecs)

    masks_pred = None
    if self.mask_net is not None:
      mask_scores = self.mask_net(obj_vecs.view(O, -1, 1, 1))
      masks_pred = mask_scores.squeeze(1).sigmoid()

    s_boxes, o_boxes = boxes_pred[s], boxes_pred[o]
    s_vecs, o_vecs = obj_vecs_orig[s], obj_vecs_orig[o]
    rel_aux_input = torch.cat([s_boxes, o_boxes, s_vecs, o_vecs], dim=1)
    rel_scores = self.rel_aux_net(rel_aux_input)

    H, W = self.image_size
    layout_boxes = boxes_pred if boxes_gt is None else boxes_gt
    layout_masks = None
    if masks_pred is not None:
      layout_masks = masks_pred if masks_gt is None else masks_gt

    if self.layout_noise_dim > 0 and noise is None:
      noise = torch.randn((N, self.layout_noise_dim, H, W), dtype=obj_vecs.dtype,
                          device=obj_vecs.device)
    elif self.layout_noise_dim == 0:
      noise = None

    # without boxes_gt the image loss reaches box_net through the sampling grid
    # (model.py:151-160): ops.LayoutStack.backward returns the box gradient
    bufs = ops.LayoutStack.apply(obj_vecs, layout_boxes, layout_masks, obj_to_img, N, H, W,
                                 noise, _layout.ALIGN_CORNERS,
                                 tuple(self.refinement_net.stage_extras()))
    img = self.refinement_net.forward_stack(list(bufs)).permute(0, 3, 1, 2)
    return img, boxes_pred, masks_pred, rel_scores

  def encode_scene_graphs(self, scene_graphs):
    """JSON-style scene graph(s) -> (objs, triples, obj_to_img) int64 tensors on the
    model's device (sg2im/model.py:173-227).  Every graph gets a trailing
    ``__image__`` object that every other object is ``__in_image__``; as in the
    reference the input dictionaries are extended in place."""
    graphs = [scene_graphs] if isinstance(scene_graphs, dict) else scene_graphs
    obj_index = self.vocab['object_name_to_idx']
    pred_index = self.vocab['pred_name_to_idx']
    categories, owners, rows = [], [], []
    base = 0
    for img, graph in enumerate(graphs):
      names, rels = graph['objects'], graph['relationships']
      n_real = len(names)
      names.append('__image__')
      rels.extend([k, '__in_image__', n_real] for k in range(n_real))
      for name in names:
        if name not in obj_index:
          raise ValueError('Object "%s" not in vocab' % name)
        categories.append(obj_index[name])
        owners.append(img)
      for subj, pred, obj in rels:
        if pred not in pred_index:
          raise ValueError('Relationship "%s" not in vocab' % pred)
        rows.append([base + subj, pred_index[pred], base + obj])
      base += len(names)
    where = next(self.parameters()).device

    def as_long(values):
      return torch.tensor(values, dtype=torch.int64, device=where)

    return as_long(categories), as_long(rows), as_long(owners)

  def forward_json(self, scene_graphs):
    """sg2im/model.py:229-232."""
    objs, triples, obj_to_img = self.encode_scene_graphs(scene_graphs)
    return self.forward(objs, triples, obj_to_img)
