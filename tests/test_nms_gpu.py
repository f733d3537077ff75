"""GPU parity tests for softmax, SSDDecoder (decode + NMS) and the general
combined NMS against the CPU oracle.  Selection (which anchors, in which
order, with which class) must be identical; scores are bit exact on the
probability path; boxes within 1e-5 (exp differs by ulps between libms)."""

import numpy as np
import pytest
import torch

from oracle import box_oracle as bo
from tests.conftest import CONFIGS, VARIANCES

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


def _check_decode(priors, deltas, probs, max_total=200, thr=0.5, from_logits_z=None):
    from tf_ssd_b200.models.decoder import SSDDecoder
    dec = SSDDecoder(priors, VARIANCES, max_total_size=max_total, score_threshold=thr)
    if from_logits_z is not None:
        b, l, s = dec.call([deltas, from_logits_z], from_logits=True)
    else:
        b, l, s = dec([deltas, probs])
    rb, rl, rs, rvalid, ri = bo.ssd_decode(priors, VARIANCES, deltas, probs, max_total, thr, return_aux=True)
    np.testing.assert_array_equal(_np(dec.last_valid_detections), rvalid)
    np.testing.assert_array_equal(_np(l), rl)
    if from_logits_z is None:
        np.testing.assert_array_equal(_np(s), rs)
    else:
        np.testing.assert_allclose(_np(s), rs, rtol=1e-5)
    np.testing.assert_allclose(_np(b), rb, rtol=1e-5, atol=1e-6)
    return rvalid


@pytest.mark.parametrize("name,B", [("mobilenet_v2", 8), ("vgg16", 4), ("vgg16_512", 2)])
def test_decoder_parity_random(name, B):
    from tf_ssd_b200 import synth
    priors = bo.prior_boxes(*CONFIGS[name][:2])
    deltas, z = synth.make_head_outputs(B, priors.shape[0], 21, seed=31 + B)
    probs = bo.softmax(z)
    valid = _check_decode(priors, deltas, probs)
    assert valid.min() > 5


def test_decoder_fused_softmax_and_softmax_kernel():
    from tf_ssd_b200 import synth, _ffi
    priors = bo.prior_boxes(*CONFIGS["mobilenet_v2"][:2])
    deltas, z = synth.make_head_outputs(4, priors.shape[0], 21, seed=77)
    zt = _ffi.to_dev(z)
    probs = torch.empty_like(zt)
    _ffi.check(_ffi.lib().ssd_softmax(_ffi.ptr(zt), zt.shape[0] * zt.shape[1], 21, _ffi.ptr(probs), _ffi.stream()))
    np.testing.assert_allclose(_np(probs), bo.softmax(z), rtol=2e-6, atol=1e-9)
    # fused-softmax decoder == decoder fed with the device softmax, bit for bit
    from tf_ssd_b200.models.decoder import SSDDecoder
    dec = SSDDecoder(priors, VARIANCES)
    b1, l1, s1 = dec.call([deltas, zt], from_logits=True)
    b2, l2, s2 = dec([deltas, probs])
    assert torch.equal(b1, b2) and torch.equal(l1, l2) and torch.equal(s1, s2)
    _check_decode(priors, deltas, _np(probs))


@pytest.mark.parametrize("n_cls", [4, 12])       # ~100 candidates per class (candidate-by-candidate loop) / ~33 (bit-mask path)
def test_decoder_integer_ties(n_cls):
    """Equal scores and IoU exactly at 0.5: order must follow the documented
    oracle rule (score desc, class asc, anchor asc); suppression is strict >."""
    rng = np.random.default_rng(12)
    k = 32
    N = 400
    y1 = rng.integers(0, k - 8, N); x1 = rng.integers(0, k - 8, N)
    h = rng.choice([4, 8], N); w = rng.choice([4, 8], N)
    priors = (np.stack([y1, x1, y1 + h, x1 + w], -1) / k).astype(np.float32)
    deltas = np.zeros((3, N, 4), np.float32)                       # exp(0) exact -> boxes exact
    levels = np.array([0.55, 0.6, 0.75, 0.9], np.float32)
    cls = rng.integers(0, n_cls, (3, N))
    sc = levels[rng.integers(0, 4, (3, N))]
    probs = np.zeros((3, N, n_cls), np.float32)
    np.put_along_axis(probs, cls[..., None], sc[..., None], axis=2)
    rest = (1 - sc) / (n_cls - 1)
    probs = np.where(probs == 0, rest[..., None], probs).astype(np.float32)
    from tf_ssd_b200.models.decoder import SSDDecoder
    dec = SSDDecoder(priors, [1., 1., 1., 1.], max_total_size=50)
    b, l, s = dec([deltas, probs])
    rb, rl, rs, rvalid, ri = bo.ssd_decode(priors, [1., 1., 1., 1.], deltas, probs, 50, 0.5, return_aux=True)
    np.testing.assert_array_equal(_np(s), rs)
    np.testing.assert_array_equal(_np(l), rl)
    np.testing.assert_array_equal(_np(b), rb)                      # bit exact boxes -> same anchors in same order
    assert rvalid.max() == 50


def test_decoder_edge_cases():
    priors = bo.prior_boxes(*CONFIGS["mobilenet_v2"][:2])
    N = priors.shape[0]
    from tf_ssd_b200.models.decoder import SSDDecoder
    dec = SSDDecoder(priors, VARIANCES)
    # nothing above threshold / everything background
    probs = np.zeros((2, N, 21), np.float32); probs[..., 0] = 1
    b, l, s = dec([np.zeros((2, N, 4), np.float32), probs])
    assert not _np(b).any() and not _np(l).any() and not _np(s).any()
    assert _np(dec.last_valid_detections).tolist() == [0, 0]
    # every anchor a confident candidate of one class: per-class cap 200 and total cap 200
    probs2 = np.full((1, N, 21), 0.01, np.float32)
    rng = np.random.default_rng(2)
    probs2[0, :, 7] = 0.8 + 0.1 * rng.random(N).astype(np.float32)
    deltas = (0.1 * rng.standard_normal((1, N, 4))).astype(np.float32)
    valid = _check_decode(priors, deltas, probs2)
    assert valid[0] > 20
    # score exactly at the threshold is not a candidate (strict >)
    probs3 = np.zeros((1, N, 21), np.float32); probs3[..., 0] = 0.5; probs3[..., 3] = 0.5
    b3, l3, s3 = dec([np.zeros((1, N, 4), np.float32), probs3])
    assert _np(dec.last_valid_detections)[0] == 0
    # boxes beyond the unit square are clipped on output only
    big = np.zeros((1, N, 4), np.float32); big[..., 2:] = 15.0
    probs4 = np.full((1, N, 21), 0.0, np.float32); probs4[0, 5, 2] = 0.9; probs4[0, :, 0] = 0.1
    b4, _, _ = dec([big, probs4])
    assert _np(b4)[0, 0].tolist() == [0.0, 0.0, 1.0, 1.0]
    _check_decode(priors, big, probs4)
    # small max_total_size, other threshold
    deltas5, z5 = __import__("tf_ssd_b200.synth", fromlist=["x"]).make_head_outputs(2, N, 21, seed=5)
    _check_decode(priors, deltas5, bo.softmax(z5), max_total=10, thr=0.3)


@pytest.mark.parametrize("q_is_L", [False, True])
def test_combined_nms_general(q_is_L):
    from tf_ssd_b200.utils import bbox_utils
    rng = np.random.default_rng(6)
    B, N, L = 3, 500, 6
    q = L if q_is_L else 1
    c = rng.random((B, N, q, 2)); wh = rng.uniform(0.05, 0.4, (B, N, q, 2))
    boxes = np.concatenate([c - wh / 2, c + wh / 2], -1).astype(np.float32)
    boxes[0, :20] = boxes[0, :20][..., [2, 3, 0, 1]]               # flipped corners are canonicalised by TF
    scores = rng.random((B, N, L)).astype(np.float32)
    kw = dict(max_output_size_per_class=30, max_total_size=64, iou_threshold=0.45, score_threshold=0.6)
    ob, os_, oc, ov = bbox_utils.non_max_suppression(boxes, scores, **kw)
    rb, rs, rc, rv = bo.combined_nms(boxes, scores, 30, 64, 0.45, 0.6, clip_boxes=True)
    np.testing.assert_array_equal(_np(ov), rv)
    np.testing.assert_array_equal(_np(os_), rs)
    np.testing.assert_array_equal(_np(oc), rc)
    np.testing.assert_array_equal(_np(ob), rb)
    # no clipping + default thresholds (score_threshold=-inf keeps everything, negatives included)
    sc2 = (scores[:1, :60] - 0.5).astype(np.float32)
    ob2, os2, oc2, ov2 = bbox_utils.non_max_suppression(boxes[:1, :60] * 1.5, sc2, max_output_size_per_class=60,
                                                        max_total_size=100, clip_boxes=False)
    rb2, rs2, rc2, rv2 = bo.combined_nms(boxes[:1, :60] * 1.5, sc2, 60, 100, 0.5, float("-inf"), clip_boxes=False)
    np.testing.assert_array_equal(_np(ov2), rv2)
    np.testing.assert_array_equal(_np(os2), rs2)
    np.testing.assert_array_equal(_np(oc2), rc2)
    np.testing.assert_array_equal(_np(ob2), rb2)
    with pytest.raises(NotImplementedError):
        bbox_utils.non_max_suppression(boxes, scores, pad_per_class=True, **kw)


def test_decoder_full_size_properties():
    """Config-5 shape (N = 24564, B = 16): properties that need no oracle."""
    from tf_ssd_b200 import synth
    from tf_ssd_b200.models.decoder import SSDDecoder
    from tf_ssd_b200.utils import bbox_utils
    fms, ars, N = CONFIGS["vgg16_512"]
    priors = bbox_utils.generate_prior_boxes(fms, ars)
    B = 16
    deltas, z = synth.make_head_outputs(B, N, 21, seed=55)
    dec = SSDDecoder(priors, VARIANCES)
    b, l, s = dec.call([deltas, z], from_logits=True)
    valid = _np(dec.last_valid_detections)
    b, l, s = _np(b), _np(l), _np(s)
    assert (valid > 0).all() and (valid <= 200).all()
    for i in range(B):
        v = valid[i]
        assert (np.diff(s[i, :v]) <= 0).all()                      # sorted by score
        assert (s[i, :v] > 0.5).all() and (l[i, :v] >= 1).all()
        assert not s[i, v:].any() and not b[i, v:].any()
        assert b[i].min() >= 0 and b[i].max() <= 1


@pytest.mark.parametrize("thr,scale", [(0.5, 3.0), (0.3, 1.0), (0.05, 1.0)])
def test_decoder_candidate_capacity(thr, scale):
    """More than one class per anchor may pass the threshold when the scores are not normalised (scale > 1) or the
    threshold is below 0.5: the candidate list must grow instead of reporting an overflow (TensorFlow has no limit)."""
    from tf_ssd_b200 import synth
    from tf_ssd_b200.models.decoder import SSDDecoder
    priors = bo.prior_boxes(*CONFIGS["mobilenet_v2"][:2])[:600]
    deltas, z = synth.make_head_outputs(2, priors.shape[0], 21, seed=5, hot_fraction=0.3, hot_boost=2.0)
    probs = (bo.softmax(z) * np.float32(scale)).astype(np.float32)
    dec = SSDDecoder(priors, VARIANCES, max_total_size=300, score_threshold=thr)
    b, l, s = dec([deltas, probs])
    rb, rl, rs, rvalid, _ = bo.ssd_decode(priors, VARIANCES, deltas, probs, 300, thr, return_aux=True)
    n_cand = ((probs > thr) & (probs.argmax(-1, keepdims=True) != 0)).sum((1, 2))
    if scale > 1.0 or thr < 0.1:
        assert (n_cand > priors.shape[0]).any()                 # really exceeds the N-candidate assumption
    np.testing.assert_array_equal(_np(dec.last_valid_detections), rvalid)
    np.testing.assert_array_equal(_np(l), rl)
    np.testing.assert_array_equal(_np(s), rs)
    np.testing.assert_allclose(_np(b), rb, rtol=1e-5, atol=1e-6)
