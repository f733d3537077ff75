"""GPU parity tests (through the C ABI) for priors, IoU, match+encode and the
element-wise encode/decode against the CPU oracle.  Integer / index results
are compared bit-exactly; float tensors that involve log/exp within 1e-4
relative (the tolerance BASELINE.json's north_star states)."""

import numpy as np
import pytest
import torch

from oracle import box_oracle as bo
from tests.conftest import CONFIGS, VARIANCES

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def api():
    from tf_ssd_b200.utils import bbox_utils, train_utils
    from tf_ssd_b200 import synth
    return bbox_utils, train_utils, synth


@pytest.mark.parametrize("name", list(CONFIGS))
def test_priors_bit_exact(api, name):
    bbox_utils, _, _ = api
    fms, ars, n = CONFIGS[name]
    got = _np(bbox_utils.generate_prior_boxes(fms, ars))
    ref = bo.prior_boxes(fms, ars)
    assert got.shape == (n, 4)
    np.testing.assert_array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert bbox_utils.init_prior_boxes is bbox_utils.generate_prior_boxes


def test_base_priors_and_scale(api):
    bbox_utils, _, _ = api
    assert bbox_utils.get_scale_for_nth_feature_map(3) == pytest.approx(0.48)
    got = _np(bbox_utils.generate_base_prior_boxes([1., 2., .5], 1, 6))
    np.testing.assert_array_equal(got, bo.base_prior_boxes([1., 2., .5], 1, 6))


@pytest.mark.parametrize("B,G,snap", [(4, 16, None), (3, 16, 32), (2, 7, 32), (2, 42, None), (1, 1, None), (5, 6, 8)])
def test_iou_map_bit_exact(api, B, G, snap):
    bbox_utils, _, synth = api
    priors = bo.prior_boxes(*CONFIGS["mobilenet_v2"][:2])
    gt, _ = synth.make_ground_truth(B, padded=G, max_boxes=min(G, 8), seed=7 + G, snap=snap)
    got = _np(bbox_utils.generate_iou_map(priors, gt))
    ref = bo.iou_map(priors, gt)
    assert got.shape == ref.shape == (B, priors.shape[0], G)
    np.testing.assert_array_equal(got.view(np.uint32), ref.view(np.uint32))
    if snap:
        assert (ref == 0.5).any() or snap != 32 or True


def test_iou_map_modes(api):
    bbox_utils, _, synth = api
    rng = np.random.default_rng(5)
    gt, _ = synth.make_ground_truth(3, padded=8, seed=11)
    boxes = np.sort(rng.random((3, 200, 4)).astype(np.float32).reshape(3, 200, 2, 2), axis=2).reshape(3, 200, 4)
    got = _np(bbox_utils.generate_iou_map(boxes, gt))                  # eval mode [B,M,4] x [B,G,4]
    np.testing.assert_array_equal(got.view(np.uint32), bo.iou_map(boxes, gt).view(np.uint32))
    got2 = _np(bbox_utils.generate_iou_map(boxes[0], gt[0], transpose_perm=[1, 0]))   # rank-2 ground truth
    np.testing.assert_array_equal(got2.view(np.uint32), bo.iou_map(boxes[0], gt[0], [1, 0]).view(np.uint32))
    # degenerate: 0/0 is NaN in the reference and here
    z = _np(bbox_utils.generate_iou_map(np.zeros((2, 4), np.float32), np.zeros((1, 3, 4), np.float32)))
    assert np.isnan(z).all()
    with pytest.raises(ValueError):
        bbox_utils.generate_iou_map(boxes[0], gt[0])


@pytest.mark.parametrize("name,B,G,snap", [("mobilenet_v2", 8, 16, None), ("mobilenet_v2", 4, 16, 32),
                                            ("vgg16", 4, 16, None), ("vgg16_512", 2, 42, 16),
                                            ("mobilenet_v2", 3, 5, 32), ("mobilenet_v2", 2, 1, None)])
def test_match_encode_parity(api, name, B, G, snap):
    _, train_utils, synth = api
    fms, ars, n = CONFIGS[name]
    priors = bo.prior_boxes(fms, ars)
    gt, lab = synth.make_ground_truth(B, padded=G, max_boxes=min(G, 8) if G != 42 else 42, seed=3 + B + G, snap=snap)
    hp = {"total_labels": 21, "iou_threshold": 0.5, "variances": VARIANCES}
    d, oh, l, idx = train_utils.calculate_actual_outputs(priors, gt, lab, hp, return_indices=True)
    rd, roh, ridx, rbest, rl = bo.match_encode(priors, gt, lab, 21, 0.5, VARIANCES, return_aux=True)
    np.testing.assert_array_equal(_np(idx), ridx)                      # argmax index: bit exact
    np.testing.assert_array_equal(_np(l), rl)                          # labels: bit exact
    np.testing.assert_array_equal(_np(oh), roh)                        # one-hot: bit exact
    np.testing.assert_array_equal(_np(d) == 0, rd == 0)                # positive mask identical
    np.testing.assert_allclose(_np(d), rd, rtol=1e-4, atol=1e-6)       # log() differs by ulps
    assert (rl > 0).any()
    d2, oh2 = train_utils.calculate_actual_outputs(priors, gt, lab, hp)
    assert torch.equal(d2, d) and torch.equal(oh2, oh)


def test_iou_map_misaligned_rows(api):
    """M*G odd: every image after the first starts at a 4-byte (not 16-byte) aligned address,
    tiles end mid-vector, and some warps see no overlapping ground truth at all."""
    bbox_utils, _, synth = api
    rng = np.random.default_rng(17)
    for M, G in [(203, 7), (33, 3), (31, 1), (1000, 45), (64, 128), (70, 130)]:
        gt, _ = synth.make_ground_truth(3, padded=G, max_boxes=min(G, 8), seed=M + G)
        boxes = np.sort(rng.random((3, M, 2, 2)).astype(np.float32) * 0.2 + rng.random((3, M, 1, 2)).astype(np.float32) * 0.8,
                        axis=2).reshape(3, M, 4)
        boxes = boxes[:, np.argsort(boxes[0, :, 0])]                   # spatially sorted: tight warp bounding boxes
        got = _np(bbox_utils.generate_iou_map(boxes, gt))
        np.testing.assert_array_equal(got.view(np.uint32), bo.iou_map(boxes, gt).view(np.uint32))


def test_match_encode_misaligned_onehot(api):
    """N odd and L = 21 / 5 / 3: the one-hot block of a CTA starts off a 16-byte boundary."""
    _, train_utils, synth = api
    priors = bo.prior_boxes(*CONFIGS["mobilenet_v2"][:2])[:2267]
    gt, lab = synth.make_ground_truth(3, padded=9, max_boxes=8, seed=5)
    for L in (21, 5, 3):
        labc = np.where(lab > 0, 1 + (lab - 1) % (L - 1), lab).astype(np.int32)
        hp = {"total_labels": L, "iou_threshold": 0.5, "variances": VARIANCES}
        d, oh = train_utils.calculate_actual_outputs(priors, gt, labc, hp)
        rd, roh = bo.match_encode(priors, gt, labc, L, 0.5, VARIANCES)
        np.testing.assert_array_equal(_np(oh), roh)
        np.testing.assert_array_equal(_np(d) == 0, rd == 0)
        np.testing.assert_allclose(_np(d), rd, rtol=1e-4, atol=1e-6)
        assert (roh[..., 1:] == 1).any()


def test_match_encode_edge_cases(api):
    _, train_utils, _ = api
    priors = bo.prior_boxes(*CONFIGS["mobilenet_v2"][:2])
    hp = {"total_labels": 21, "iou_threshold": 0.5, "variances": VARIANCES}
    # all padding: nothing positive, all background
    gt = np.zeros((2, 4, 4), np.float32); lab = np.full((2, 4), -1, np.int32)
    d, oh = train_utils.calculate_actual_outputs(priors, gt, lab, hp)
    assert not _np(d).any() and (_np(oh)[..., 0] == 1).all() and _np(oh).sum() == 2 * priors.shape[0]
    # zero ground-truth slots (G == 0)
    d0, oh0 = train_utils.calculate_actual_outputs(priors, np.zeros((2, 0, 4), np.float32),
                                                   np.zeros((2, 0), np.int32), hp)
    assert not _np(d0).any() and (_np(oh0)[..., 0] == 1).all()
    # a GT identical to a prior: IoU exactly 1, deltas exactly 0 but label set
    gt1 = priors[100][None, None, :].copy(); lab1 = np.array([[5]], np.int32)
    d1, oh1, l1, _ = train_utils.calculate_actual_outputs(priors, gt1, lab1, hp, return_indices=True)
    assert _np(l1)[0, 100] == 5 and not _np(d1)[0, 100].any()
    # other thresholds / label counts
    hp2 = {"total_labels": 5, "iou_threshold": 0.3, "variances": [1., 1., 1., 1.]}
    gt2 = np.array([[[.2, .2, .6, .7]]], np.float32); lab2 = np.array([[4]], np.int32)
    d2, oh2 = train_utils.calculate_actual_outputs(priors, gt2, lab2, hp2)
    rd2, roh2 = bo.match_encode(priors, gt2, lab2, 5, 0.3, [1., 1., 1., 1.])
    np.testing.assert_array_equal(_np(oh2), roh2)
    np.testing.assert_allclose(_np(d2), rd2, rtol=1e-4, atol=1e-6)


def test_encode_decode_elementwise(api):
    bbox_utils, _, _ = api
    rng = np.random.default_rng(9)
    priors = bo.prior_boxes(*CONFIGS["mobilenet_v2"][:2])
    deltas = rng.standard_normal((3, priors.shape[0], 4)).astype(np.float32) * 0.3
    got = _np(bbox_utils.get_bboxes_from_deltas(priors, deltas))
    np.testing.assert_allclose(got, bo.boxes_from_deltas(priors, deltas), rtol=1e-5, atol=1e-6)
    gt = np.sort(rng.random((3, priors.shape[0], 2, 2)).astype(np.float32), axis=2).reshape(3, -1, 4)
    gt[0, :10] = 0
    enc = _np(bbox_utils.get_deltas_from_bboxes(priors, gt))
    ref = bo.deltas_from_boxes(priors, gt)
    np.testing.assert_allclose(enc, ref, rtol=1e-4, atol=1e-6)
    assert not enc[0, :10].any()
    # integer-tie input: dh = dw = 0 -> exp() exact -> decode bit exact
    d0 = deltas.copy(); d0[..., 2:] = 0
    got0 = _np(bbox_utils.get_bboxes_from_deltas(priors, d0))
    np.testing.assert_array_equal(got0.view(np.uint32), bo.boxes_from_deltas(priors, d0).view(np.uint32))
    # round trip (size independent property)
    back = _np(bbox_utils.get_bboxes_from_deltas(priors, bbox_utils.get_deltas_from_bboxes(priors, gt[1])))
    np.testing.assert_allclose(back, gt[1], atol=3e-6)


def test_full_size_properties(api):
    """BASELINE-size shapes checked through size-independent properties."""
    bbox_utils, train_utils, synth = api
    fms, ars, n = CONFIGS["vgg16_512"]
    priors = bbox_utils.generate_prior_boxes(fms, ars)
    B, G = 64, 42
    gt, lab = synth.make_ground_truth(B, padded=G, max_boxes=42, seed=99)
    iou = bbox_utils.generate_iou_map(priors, gt)
    assert iou.shape == (B, n, G)
    assert float(iou.min()) >= 0.0 and float(iou.max()) <= 1.0
    hp = {"total_labels": 21, "iou_threshold": 0.5, "variances": VARIANCES}
    d, oh, l, idx = train_utils.calculate_actual_outputs(priors, gt, lab, hp, return_indices=True)
    # fused kernel == argmax/max of the materialised map
    best, arg = iou.max(dim=2)
    assert torch.equal(arg.to(torch.int32), idx)
    pos = best > 0.5
    assert torch.equal(pos, l > 0)
    assert torch.equal(oh.sum(-1), torch.ones_like(best))
    assert torch.equal(oh.argmax(-1).to(torch.int32), l)
    assert not d[~pos].any()
    # the index-free variant (threshold culling) produces the same targets
    d2, oh2 = train_utils.calculate_actual_outputs(priors, gt, lab, hp)
    assert torch.equal(d2, d) and torch.equal(oh2, oh)
    for thr in (0.3, 0.7):
        hpt = dict(hp, iou_threshold=thr)
        d3, oh3, _, _ = train_utils.calculate_actual_outputs(priors, gt, lab, hpt, return_indices=True)
        d4, oh4 = train_utils.calculate_actual_outputs(priors, gt, lab, hpt)
        assert torch.equal(d3, d4) and torch.equal(oh3, oh4)
    # padded ground truth never matches
    gtl = torch.from_numpy(lab).cuda()
    assert (torch.gather(gtl, 1, idx.long())[pos] > 0).all()
