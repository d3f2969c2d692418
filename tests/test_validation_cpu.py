"""The rows next to the training step on CPU: the validation
oracle is pinned against tests/golden/aux.pt (outputs of the unmodified
reference: imagenet_deprocess_batch, jaccard, scripts/train.py check_model), and
the product's host logic (sg2im_b200.validate / metrics) is replayed against
the same fixtures with the CUDA op boundary swapped out (tests/cpu_shim.py)."""
import contextlib
import io

import pytest
import torch

from conftest import load_golden, rel_err
from cpu_shim import cpu_ops
from oracle import validation_oracle as vorc


def _bytes_close(a, b, frac=0.005):
  """uint8 images equal up to 1 LSB on at most `frac` of the bytes (fp32 noise
  of a different-but-equivalent op order can flip a truncation)."""
  assert a.dtype == torch.uint8 and b.dtype == torch.uint8 and a.shape == b.shape
  d = (a.int() - b.int()).abs()
  assert int(d.max()) <= 1, int(d.max())
  assert float((d > 0).float().mean()) <= frac


def test_oracle_deprocess_bit_exact():
  g = load_golden('aux.pt')['deprocess']
  assert torch.equal(vorc.imagenet_deprocess_batch(g['imgs']), g['rescaled'])
  assert torch.equal(vorc.imagenet_deprocess_batch(g['imgs'], rescale=False), g['plain'])
  # strided input (NCHW view of an NHWC buffer, what the generator returns)
  nhwc_view = g['imgs'].permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
  assert torch.equal(vorc.imagenet_deprocess_batch(nhwc_view), g['rescaled'])


def test_jaccard_oracle_and_product():
  from sg2im_b200.metrics import jaccard, intersection
  g = load_golden('aux.pt')['jaccard']
  assert torch.equal(vorc.jaccard(g['pred'], g['gt']), g['value'])
  assert torch.equal(jaccard(g['pred'], g['gt']), g['value'])
  inter = intersection(g['pred'], g['gt'])
  assert inter[1].item() == 0.0 and inter.shape == (9,)


@pytest.mark.parametrize('name', ['check_vg', 'check_coco'])
def test_oracle_check_model_matches_reference(name):
  g = load_golden('aux.pt')[name]
  sd = {k: v.clone() for k, v in g['sd'].items()}
  mean_losses, samples, batch_data, avg_iou = vorc.check_model(
      sd, g['kwargs']['image_size'], g['args'], g['loader'])
  assert set(mean_losses) == set(g['mean_losses'])
  for k, v in g['mean_losses'].items():
    assert abs(mean_losses[k] - v) <= 1e-5 * max(1.0, abs(v)), k
  assert abs(float(avg_iou) - float(g['avg_iou'])) < 1e-6
  assert torch.equal(samples['gt_img'], g['samples']['gt_img'])
  for k in ('gt_box_gt_mask', 'gt_box_pred_mask', 'pred_box_pred_mask'):
    _bytes_close(samples[k], g['samples'][k])
  assert rel_err(batch_data['boxes_pred'], g['batch_data']['boxes_pred']) < 1e-5
  assert rel_err(batch_data['masks_pred'], g['batch_data']['masks_pred']) < 1e-5
  for k, v in g['bn_after'].items():                  # five train-mode forwards moved the statistics
    assert rel_err(sd[k].float(), v.float()) < 1e-5, k


@pytest.mark.parametrize('name', ['check_vg', 'check_coco'])
def test_product_check_model_wiring(name):
  from sg2im_b200.model import Sg2ImModel
  from sg2im_b200.validate import check_model
  g = load_golden('aux.pt')[name]
  with cpu_ops():
    with contextlib.redirect_stdout(io.StringIO()):
      model = Sg2ImModel(vocab=g['vocab'], **g['kwargs'])
    model.load_state_dict(g['sd'])
    model.train()
    mean_losses, samples, batch_data, avg_iou = check_model(
        g['args'], 0, g['loader'], model, deprocess=vorc.imagenet_deprocess_batch)
  assert set(mean_losses) == set(g['mean_losses'])
  for k, v in g['mean_losses'].items():
    assert abs(mean_losses[k] - v) <= 1e-5 * max(1.0, abs(v)), k
  assert abs(float(avg_iou) - float(g['avg_iou'])) < 1e-6
  assert set(samples) == set(g['samples'])
  assert torch.equal(samples['gt_img'], g['samples']['gt_img'])
  for k in ('gt_box_gt_mask', 'gt_box_pred_mask', 'pred_box_pred_mask'):
    _bytes_close(samples[k], g['samples'][k])
  assert set(batch_data) == set(g['batch_data'])
  for k, v in g['batch_data'].items():
    if v is None:
      assert batch_data[k] is None, k
    elif v.dtype == torch.int64:
      assert torch.equal(batch_data[k], v), k
    else:
      assert rel_err(batch_data[k], v) < 1e-5, k
  sd = model.state_dict()
  for k, v in g['bn_after'].items():
    assert rel_err(sd[k].float(), v.float()) < 1e-5, k


def test_check_model_rejects_bad_loaders():
  from sg2im_b200.validate import check_model
  lin = torch.nn.Linear(1, 1)
  with pytest.raises(ValueError):
    check_model({}, 0, [], lin, device='cpu')
  with pytest.raises(ValueError):
    check_model({}, 0, [(torch.zeros(1),) * 5], lin, device='cpu')


def test_oracle_align_corners_true_matches_patched_reference():
  """The torch-0.4 sampling convention (align_corners=True) of the inference
  path: oracle vs the reference's functions run with grid_sample's default
  flipped (tests/golden/make_golden_aux.py §4)."""
  from oracle import sg2im_oracle as orc
  g = load_golden('aux.pt')['align_corners']
  lay, crp = load_golden('layout.pt'), load_golden('crop.pt')
  m = orc.masks_to_layout(lay['rvecs'], lay['rboxes'], lay['rmasks'], lay['robj_to_img'], 24, 40, 3,
                          align_corners=True)
  assert rel_err(m, g['masks']) < 1e-6
  b = orc.boxes_to_layout(lay['vecs'], lay['boxes'], lay['obj_to_img'], 24, 20, 2, align_corners=True)
  assert rel_err(b, g['boxes']) < 1e-6
  c = orc.crop_bbox_batch(crp['feats'], crp['boxes'], crp['bbox_to_feats'], 6, 7, align_corners=True)
  assert rel_err(c, g['crops']) < 1e-6
  # and it is a different function from the default convention
  m0 = orc.masks_to_layout(lay['rvecs'], lay['rboxes'], lay['rmasks'], lay['robj_to_img'], 24, 40, 3)
  assert rel_err(m0, g['masks']) > 1e-3
