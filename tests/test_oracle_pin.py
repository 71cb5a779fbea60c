"""Pin the restated torchvision box ops (oracle/tv_ref.py, oracle/boxops_ref.c) to the installed
torchvision CPU kernels -- the third-party dependency the reference calls."""
import os

import numpy as np
import pytest
import torch
import torchvision
from torchvision.ops import boxes as tvb

from object_detectors_b200 import synthetic as syn
from oracle import cref, tv_ref


@pytest.mark.parametrize("seed,n,k,thr", [(1, 200, 8, 0.5), (2, 900, 20, 0.6), (3, 2000, 30, 0.7),
                                          (4, 1, 0, 0.5), (5, 65, 2, 0.3)])
def test_nms_matches_torchvision(seed, n, k, thr):
    b, s, _ = syn.random_boxes(seed, n, clusters=k)
    want = tvb.nms(torch.from_numpy(b), torch.from_numpy(s), thr).numpy()
    np.testing.assert_array_equal(want, tv_ref.nms(torch.from_numpy(b), torch.from_numpy(s), thr).numpy())
    np.testing.assert_array_equal(want, cref.nms_tv(b, s, thr))


def test_nms_threshold_is_double_and_ties_prefer_low_index():
    # IoU of these two boxes is exactly 0.6f; (double)0.6f > 0.6 so torchvision suppresses.
    b = torch.tensor([[0.0, 0.0, 4.0, 4.0], [0.0, 0.0, 4.0, 2.4]])
    s = torch.tensor([0.5, 0.5])
    want = tvb.nms(b, s, 0.6).numpy()
    np.testing.assert_array_equal(want, tv_ref.nms(b, s, 0.6).numpy())
    np.testing.assert_array_equal(want, cref.nms_tv(b.numpy(), s.numpy(), 0.6))
    assert want.tolist() == [0]
    # zero-area pair: NaN IoU never suppresses
    z = torch.tensor([[1.0, 1.0, 1.0, 1.0], [1.0, 1.0, 1.0, 1.0]])
    assert tvb.nms(z, s, 0.5).tolist() == cref.nms_tv(z.numpy(), s.numpy(), 0.5).tolist() == [0, 1]


@pytest.mark.parametrize("seed,n,c", [(7, 600, 80), (8, 900, 3), (9, 1500, 5)])
def test_batched_nms_both_strategies(seed, n, c):
    b, s, l = syn.random_boxes(seed, n, clusters=15, num_classes=c)
    tb, ts, tl = torch.from_numpy(b), torch.from_numpy(s), torch.from_numpy(l)
    np.testing.assert_array_equal(tvb._batched_nms_vanilla(tb, ts, tl, 0.5).numpy(),
                                  tv_ref.batched_nms_vanilla(tb, ts, tl, 0.5).numpy())
    np.testing.assert_array_equal(tvb._batched_nms_vanilla(tb, ts, tl, 0.5).numpy(),
                                  cref.nms_tv(b, s, 0.5, l))
    np.testing.assert_array_equal(tvb._batched_nms_coordinate_trick(tb, ts, tl, 0.5).numpy(),
                                  tv_ref.batched_nms_coordinate_trick(tb, ts, tl, 0.5).numpy())
    np.testing.assert_array_equal(tvb.batched_nms(tb, ts, tl, 0.5).numpy(),
                                  tv_ref.batched_nms(tb, ts, tl, 0.5).numpy())


def test_box_iou_matches_torchvision():
    b1, _, _ = syn.random_boxes(11, 37, clusters=4)
    b2, _, _ = syn.random_boxes(12, 501, clusters=4)
    want = tvb.box_iou(torch.from_numpy(b1), torch.from_numpy(b2)).numpy()
    np.testing.assert_array_equal(want, tv_ref.box_iou(torch.from_numpy(b1), torch.from_numpy(b2)).numpy())
    np.testing.assert_array_equal(want, cref.box_iou_tv(b1, b2))


def test_versions_recorded():
    # the oracle is pinned against THIS torchvision; a different wheel must re-run the pin
    assert torchvision.__version__.startswith("0.26"), torchvision.__version__


def test_encode_boxes_matches_reference_utils():
    """oracle.tv_ref.encode_boxes == the reference's vendored BoxCoder.encode_single (tvision/_utils.py:80-125,160-166);
    skipped on the GPU box, where /root/reference does not exist (the installed torchvision copy is used there)."""
    import sys
    g = np.random.Generator(np.random.PCG64(8))
    c = g.uniform(50, 500, (200, 2)); s = np.exp(g.uniform(2, 5, (200, 2)))
    prop = torch.from_numpy(np.concatenate([c - s / 2, c + s / 2], 1).astype(np.float32))
    c2 = c + g.normal(0, 5, (200, 2)); s2 = s * np.exp(g.normal(0, 0.2, (200, 2)))
    ref = torch.from_numpy(np.concatenate([c2 - s2 / 2, c2 + s2 / 2], 1).astype(np.float32))
    if os.path.isdir("/root/reference/torchvision_models"):
        sys.path.insert(0, "/root/reference/torchvision_models")
        from tvision import _utils as ref_utils
    else:
        from torchvision.models.detection import _utils as ref_utils
    want = ref_utils.BoxCoder((10.0, 10.0, 5.0, 5.0)).encode_single(ref, prop)
    np.testing.assert_array_equal(tv_ref.encode_boxes(ref, prop, (10.0, 10.0, 5.0, 5.0)).numpy(), want.numpy())


def test_clip_and_small_box_helpers_match_torchvision():
    """clip_boxes_to_image / remove_small_boxes restatements == the installed torchvision functions the reference calls
    (rpn.py:260,263; roi_heads.py:746,767), incl. a leading batch dimension and sizes exactly at the threshold."""
    g = np.random.default_rng(5)
    b = torch.from_numpy((g.standard_normal((2, 500, 4)) * 300 + 200).astype(np.float32))
    assert torch.equal(tv_ref.clip_boxes_to_image(b, (480, 640)), tvb.clip_boxes_to_image(b, (480, 640)))
    f = b[0].clone()
    f[:, 2:] = f[:, :2] + torch.rand(500, 2) * 3
    f[7, 2] = f[7, 0] + 1.0
    for ms in (1e-3, 1e-2, 1.0):
        assert torch.equal(tv_ref.remove_small_boxes(f, ms), tvb.remove_small_boxes(f, ms))
