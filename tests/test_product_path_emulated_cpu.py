"""The PRODUCT path on CPU: the shipped op layer (sg2im_b200/ops.py: ctypes marshalling, slice /
stride arithmetic, autograd Functions) and nn.Module mirror, running on the kernel sources compiled
for the host (tests/emul_device.py).  The bodies are the `-m gpu` parity tests themselves
(tests/test_gpu_model.py, tests/test_gpu_ops.py) with the device swapped — same fixtures, same
tolerances — restricted to the exact-fp32 configuration (the tcgen05 / TMA kernels have no host
build).  What the hardware run adds on top: the tensor-core path, CUDA graphs, NCCL, timing.
"""
import shutil

import pytest
import torch

import test_gpu_model as G
import test_gpu_ops as O
from emul_device import emulated_device

pytestmark = pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++ (C++20)')


@pytest.fixture
def emul(monkeypatch):
  cpu = lambda: torch.device('cpu')                      # noqa: E731
  monkeypatch.setattr(G, 'dev', cpu)
  monkeypatch.setattr(O, 'dev', cpu)
  with emulated_device() as lib:
    yield lib


def test_generator_vg_coco_eval_through_the_real_op_layer(emul):
  G.test_generator_forward_vg_coco_eval()


def test_forward_json_config1(emul):
  G.test_config1_sheep_forward_json()


def test_discriminators(emul):
  G.test_discriminators_forward()


def test_generator_gradients(emul):
  G.test_generator_gradients_vs_oracle()


def test_two_training_iterations_match_reference(emul):
  G.test_two_training_iterations_match_reference()


def test_gconv_layer_forward_backward(emul):
  O.test_gconv_layer_forward_backward()
  O.test_graph_pool_empty_and_single()


@pytest.mark.parametrize('case', [0, 2, 4])
def test_conv_forward_dgrad_wgrad(emul, case):
  O.test_conv_forward_dgrad_wgrad(*O.CONV_CASES[case])


def test_linear_bn_upsample_slice(emul):
  O.test_linear_relu(7, 24, 46)
  O.test_bn_act_upsample_slice(True, 2, 0.2, 64)
  O.test_bn_act_upsample_slice(False, 1, 0.0, 7)
  O.test_upsample_then_bn_running_var()


def test_layout_and_crop(emul):
  O.test_layout_golden_demo_and_random()
  O.test_layout_backward(12, 3, 16, 8, 24, 20)
  O.test_crop_golden_and_backward()


def test_the_emulated_device_has_no_tensor_core_kernels(emul):
  """A convolution that would take the tcgen05 path fails loudly here: no silent stand-in."""
  from sg2im_b200 import ops
  ops.set_conv_math('tf32')
  try:
    with pytest.raises(AttributeError):
      ops.conv2d(torch.zeros(1, 8, 8, 32), torch.zeros(32, 32, 3, 3), None, 1, 1)
  finally:
    ops.set_conv_math('fp32')


# ---- the staged rows / second-generation kernels of DESIGN.md §7a, same treatment: the opt-in
# GPU tests of tests/test_gpu_next_rows.py run here on the emulated device (exact-fp32 ones).
import test_gpu_next_rows as R  # noqa: E402


@pytest.fixture
def emul_next(monkeypatch, emul):
  monkeypatch.setattr(R, 'dev', lambda: torch.device('cpu'))
  monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
  return emul


def test_denormalisation_and_validation_pass(emul_next):
  R.test_deprocess_bytes_identical_to_reference()
  R.test_check_model_matches_reference('check_vg')
  R.test_check_model_matches_reference('check_coco')


def test_align_corners_true_sampling(emul_next):
  R.test_align_corners_true_layout_and_crop()


def test_eval_bn_folding(emul_next):
  R.test_eval_bn_folding_sheep('fp32')


@pytest.mark.parametrize('case', [(4, 16, 16, 64, 1, 0), (4, 16, 16, 64, 2, 40), (3, 5, 7, 12, 1, 0)])
def test_bn_backward_v2_through_the_op_layer(emul_next, case):
  R.test_bn_backward_v2_matches_v1_and_torch(*case)


def test_layout_v2_through_the_op_layer(emul_next):
  R.test_layout_backward_v2(12, 3, 16, 8, 24, 20, True)
  R.test_layout_forward_v2_bit_identical(45, 2, 128, 16, 32, 32, 40)


def test_colsum_v2_and_flat_adam_through_the_op_layer(emul_next):
  R.test_colsum_small_kernel(37, 179)
  R.test_train_step_with_flat_adam(False)
