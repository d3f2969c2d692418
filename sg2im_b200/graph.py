"""Scene-graph convolution with the reference's surface (sg2im/graph.py).

The gather of subject/object rows, the per-triple MLP, the ordered segment
sum that replaces the two ``scatter_add`` calls (bit-exact against the CPU
reference's summation order) and the per-object MLP all run on
libsg2im_b200.so.  The CSR of the graph is built once per batch and shared by
every layer and by the backward pass (edges are identical for all layers)."""
import torch
import torch.nn as nn

from . import ops
from .layers import build_mlp


def _init_weights(module):
  """sg2im/graph.py:26-29."""
  if hasattr(module, 'weight'):
    if isinstance(module, nn.Linear):
      nn.init.kaiming_normal_(module.weight)


class GraphCSR(object):
  """(row_ptr, entries) of object -> (triple, role) incidences, in the order
  subject-role ascending t, then object-role ascending t."""

  def __init__(self, edges, num_objs):
    self.edges = edges.contiguous()
    self.num_objs = num_objs
    self.tables = ops.csr_build(self.edges, 2, num_objs)


class GraphTripleConv(nn.Module):
  """One scene-graph convolution layer (sg2im/graph.py:32-120): per triple an MLP
  over [subject | predicate | object] produces candidate subject / object vectors
  (hidden_dim wide) and the new predicate vector; candidates are pooled per
  object; a second MLP maps the pooled vectors to the new object vectors."""

  def __init__(self, input_dim, output_dim=None, hidden_dim=512,
               pooling='avg', mlp_normalization='none'):
    super().__init__()
    assert pooling in ['sum', 'avg'], 'Invalid pooling "%s"' % pooling
    self.input_dim = input_dim
    self.output_dim = input_dim if output_dim is None else output_dim
    self.hidden_dim, self.pooling = hidden_dim, pooling
    widths = {'net1': [3 * input_dim, hidden_dim, 2 * hidden_dim + self.output_dim],
              'net2': [hidden_dim, hidden_dim, self.output_dim]}
    for name in ('net1', 'net2'):                  # registration order = state_dict order
      mlp = build_mlp(widths[name], batch_norm=mlp_normalization)
      mlp.apply(_init_weights)
      setattr(self, name, mlp)

  def forward(self, obj_vecs, pred_vecs, edges, csr=None):
    """obj_vecs (O, Din), pred_vecs (T, Din), edges int64 (T, 2) ->
    (new_obj_vecs (O, Dout), new_pred_vecs (T, Dout)).  ``csr`` (optional) is
    a GraphCSR reused across layers."""
    O = obj_vecs.size(0)
    H, Dout = self.hidden_dim, self.output_dim
    if csr is None:
      csr = GraphCSR(edges, O)
    cur_t_vecs = ops.TripleGather.apply(obj_vecs, pred_vecs, csr.edges, csr.tables)
    new_t_vecs = self.net1(cur_t_vecs)
    pooled_obj_vecs, new_p_vecs = ops.GraphPool.apply(
        new_t_vecs, csr.edges, csr.tables, H, Dout, O, self.pooling == 'avg')
    new_obj_vecs = self.net2(pooled_obj_vecs)
    return new_obj_vecs, new_p_vecs


class GraphTripleConvNet(nn.Module):
  """A sequence of scene graph convolution layers (sg2im/graph.py:123-144)."""

  def __init__(self, input_dim, num_layers=5, hidden_dim=512, pooling='avg',
               mlp_normalization='none'):
    super(GraphTripleConvNet, self).__init__()
    self.num_layers = num_layers
    self.gconvs = nn.ModuleList()
    gconv_kwargs = {
      'input_dim': input_dim,
      'hidden_dim': hidden_dim,
      'pooling': pooling,
      'mlp_normalization': mlp_normalization,
    }
    for _ in range(self.num_layers):
      self.gconvs.append(GraphTripleConv(**gconv_kwargs))

  def forward(self, obj_vecs, pred_vecs, edges, csr=None):
    if csr is None:
      csr = GraphCSR(edges, obj_vecs.size(0))
    for i in range(self.num_layers):
      obj_vecs, pred_vecs = self.gconvs[i](obj_vecs, pred_vecs, edges, csr)
    return obj_vecs, pred_vecs
