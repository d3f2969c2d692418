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
  def __init__(self, vocab, image_size=(64, 64), embedding_dim=64,
               gconv_dim=128, gconv_hidden_dim=512,
               gconv_pooling='avg', gconv_num_layers=5,
               refinement_dims=(1024, 512, 256, 128, 64),
               normalization='batch', activation='leakyrelu-0.2',
               mask_size=None, mlp_normalization='none', layout_noise_dim=0,
               **kwargs):
    super(Sg2ImModel, self).__init__()

    if len(kwargs) > 0:
      print('WARNING: Model got unexpected kwargs ', kwargs)     # model.py:41-42

    self.vocab = vocab
    self.image_size = image_size
    self.layout_noise_dim = layout_noise_dim

    num_objs = len(vocab['object_idx_to_name'])
    num_preds = len(vocab['pred_idx_to_name'])
    self.obj_embeddings = nn.Embedding(num_objs + 1, embedding_dim)
    self.pred_embeddings = nn.Embedding(num_preds, embedding_dim)

    if gconv_num_layers == 0:
      self.gconv = Linear(embedding_dim, gconv_dim)
    elif gconv_num_layers > 0:
      self.gconv = GraphTripleConv(input_dim=embedding_dim, output_dim=gconv_dim,
                                   hidden_dim=gconv_hidden_dim, pooling=gconv_pooling,
                                   mlp_normalization=mlp_normalization)

    self.gconv_net = None
    if gconv_num_layers > 1:
      self.gconv_net = GraphTripleConvNet(input_dim=gconv_dim, hidden_dim=gconv_hidden_dim,
                                          pooling=gconv_pooling,
                                          num_layers=gconv_num_layers - 1,
                                          mlp_normalization=mlp_normalization)

    box_net_layers = [gconv_dim, gconv_hidden_dim, 4]
    self.box_net = build_mlp(box_net_layers, batch_norm=mlp_normalization)

    self.mask_net = None
    if mask_size is not None and mask_size > 0:
      self.mask_net = self._build_mask_net(num_objs, gconv_dim, mask_size)

    rel_aux_layers = [2 * embedding_dim + 8, gconv_hidden_dim, num_preds]
    self.rel_aux_net = build_mlp(rel_aux_layers, batch_norm=mlp_normalization)

    self.refinement_net = RefinementNetwork(
        dims=(gconv_dim + layout_noise_dim,) + tuple(refinement_dims),
        normalization=normalization, activation=activation)

  def _build_mask_net(self, num_objs, dim, mask_size):
    """sg2im/model.py:94-106."""
    output_dim = 1
    layers, cur_size = [], 1
    while cur_size < mask_size:
      layers.append(Upsample(scale_factor=2, mode='nearest'))
      layers.append(BatchNorm2d(dim))
      layers.append(Conv2d(dim, dim, kernel_size=3, padding=1))
      layers.append(ReLU())
      cur_size *= 2
    if cur_size != mask_size:
      raise ValueError('Mask size must be a power of 2')
    layers.append(Conv2d(dim, output_dim, kernel_size=1))
    return FusedSequential(*layers)

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

    if boxes_gt is None and torch.is_grad_enabled() and layout_boxes.requires_grad:
      layout_boxes = layout_boxes.detach()      # box gradients: see ops.Layout.backward
    bufs = ops.LayoutStack.apply(obj_vecs, layout_boxes, layout_masks, obj_to_img, N, H, W,
                                 noise, _layout.ALIGN_CORNERS,
                                 tuple(self.refinement_net.stage_extras()))
    img = self.refinement_net.forward_stack(list(bufs)).permute(0, 3, 1, 2)
    return img, boxes_pred, masks_pred, rel_scores

  def encode_scene_graphs(self, scene_graphs):
    """sg2im/model.py:173-227: JSON-style scene graphs -> (objs, triples,
    obj_to_img) LongTensors on the model's device.  Like the reference, the
    input dicts are modified in place (the __image__ object and its
    __in_image__ relationships are appended)."""
    if isinstance(scene_graphs, dict):
      scene_graphs = [scene_graphs]
    objs, triples, obj_to_img = [], [], []
    obj_offset = 0
    for i, sg in enumerate(scene_graphs):
      sg['objects'].append('__image__')
      image_idx = len(sg['objects']) - 1
      for j in range(image_idx):
        sg['relationships'].append([j, '__in_image__', image_idx])
      for obj in sg['objects']:
        obj_idx = self.vocab['object_name_to_idx'].get(obj, None)
        if obj_idx is None:
          raise ValueError('Object "%s" not in vocab' % obj)
        objs.append(obj_idx)
        obj_to_img.append(i)
      for s, p, o in sg['relationships']:
        pred_idx = self.vocab['pred_name_to_idx'].get(p, None)
        if pred_idx is None:
          raise ValueError('Relationship "%s" not in vocab' % p)
        triples.append([s + obj_offset, pred_idx, o + obj_offset])
      obj_offset += len(sg['objects'])
    device = next(self.parameters()).device
    objs = torch.tensor(objs, dtype=torch.int64, device=device)
    triples = torch.tensor(triples, dtype=torch.int64, device=device)
    obj_to_img = torch.tensor(obj_to_img, dtype=torch.int64, device=device)
    return objs, triples, obj_to_img

  def forward_json(self, scene_graphs):
    """sg2im/model.py:229-232."""
    objs, triples, obj_to_img = self.encode_scene_graphs(scene_graphs)
    return self.forward(objs, triples, obj_to_img)
