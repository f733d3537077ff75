"""CPU tests: the oracle against every known-answer value available for the
hot path and against independent cross-checks."""

import numpy as np
import pytest
import torch

from oracle import box_oracle as bo
from tests.conftest import CONFIGS, VARIANCES


def test_scale_kat_from_reference_tests():
    # tests/test_bbox_utils.py:19-22 of the reference: the only hot-path KAT it holds
    assert bo.scale_for_feature_map(3) == pytest.approx(0.48)
    assert [round(bo.scale_for_feature_map(k), 10) for k in range(1, 8)] == [0.2, 0.34, 0.48, 0.62, 0.76, 0.9, 1.04]


@pytest.mark.parametrize("name", list(CONFIGS))
def test_prior_counts(name):
    fms, ars, n = CONFIGS[name]
    p = bo.prior_boxes(fms, ars)
    assert p.shape == (n, 4) and p.dtype == np.float32
    assert p.min() >= 0.0 and p.max() <= 1.0


def test_prior_kats_vgg16():
    fms, ars, _ = CONFIGS["vgg16"]
    p = bo.prior_boxes(fms, ars)
    assert p.astype(np.float64).sum() == pytest.approx(17463.999998, abs=1e-5)
    exp = np.array([[0, 0, 0.11315790, 0.11315790], [0, 0, 0.08386858, 0.15457925],
                    [0, 0, 0.15457925, 0.08386858], [0, 0, 0.14354195, 0.14354195],
                    [0, 0, 0.11315790, 0.13947368]], np.float32)
    np.testing.assert_allclose(p[:5], exp, rtol=0, atol=1e-8)
    last = np.array([[0.05000001, 0.05000001, 0.94999999, 0.94999999], [0.18180194, 0, 0.81819808, 1],
                     [0, 0.18180197, 1, 0.81819803], [0.01626453, 0.01626453, 0.98373544, 0.98373544]], np.float32)
    np.testing.assert_allclose(p[-4:], last, rtol=0, atol=1e-7)


def test_prior_kats_mobilenet():
    fms, ars, _ = CONFIGS["mobilenet_v2"]
    p = bo.prior_boxes(fms, ars)
    assert p.astype(np.float64).sum() == pytest.approx(4535.999987, abs=1e-5)
    np.testing.assert_allclose(p[0], [0, 0, 0.12631579, 0.12631579], atol=1e-8)
    np.testing.assert_allclose(p[4], [0, 0, 0.12631579, 0.17894736], atol=1e-8)
    area = (p[:, 2] - p[:, 0]) * (p[:, 3] - p[:, 1])
    assert area.min() > 1.2e-2          # no 0/0 against zero-padded ground truth


def test_prior_independent_float64_rederivation():
    """Re-derive the priors in float64 with explicit loops; agree to float32 rounding."""
    fms, ars, _ = CONFIGS["mobilenet_v2"]
    rows = []
    m = len(fms)
    for i, fm in enumerate(fms):
        s, s1 = 0.2 + 0.7 / (m - 1) * i, 0.2 + 0.7 / (m - 1) * (i + 1)
        dims = [(s / np.sqrt(a), s * np.sqrt(a)) for a in ars[i]] + [(np.sqrt(s * s1),) * 2]
        for y in range(fm):
            for x in range(fm):
                cy, cx = (y + 0.5) / fm, (x + 0.5) / fm
                for h, w in dims:
                    rows.append([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2])
    ref = np.clip(np.array(rows), 0, 1)
    np.testing.assert_allclose(bo.prior_boxes(fms, ars), ref, atol=2e-7)


def test_rank_kat():
    r = bo.hard_negative_rank(np.array([[0, .5, .5, 0, 2, .1]], np.float32))
    assert r.tolist() == [[4, 1, 2, 5, 0, 3]]


def test_iou_simple_values():
    a = np.array([[0, 0, 1, 1], [0, 0, 0.5, 0.5]], np.float32)
    g = np.array([[[0, 0, 0.5, 1.0], [0, 0, 0, 0]]], np.float32)
    iou = bo.iou_map(a, g)
    assert iou.shape == (1, 2, 2)
    np.testing.assert_array_equal(iou[0, :, 0], np.array([0.5, 0.5], np.float32))
    np.testing.assert_array_equal(iou[0, :, 1], np.zeros(2, np.float32))       # padded GT -> exactly 0
    assert np.isnan(bo.iou_map(np.zeros((1, 4), np.float32), np.zeros((1, 1, 4), np.float32))).all()   # 0/0
    r2 = bo.iou_map(a, g[0], transpose_perm=[1, 0])
    np.testing.assert_array_equal(r2, iou[0])


def test_encode_decode_roundtrip_and_degenerate():
    rng = np.random.default_rng(0)
    p = bo.prior_boxes(*CONFIGS["mobilenet_v2"][:2])
    c = rng.random((p.shape[0], 2)); wh = rng.uniform(0.05, 0.5, (p.shape[0], 2))
    gt = np.concatenate([c - wh / 2, c + wh / 2], -1).astype(np.float32)
    d = bo.deltas_from_boxes(p, gt)
    back = bo.boxes_from_deltas(p, d)
    np.testing.assert_allclose(back, gt, atol=2e-6)
    z = bo.deltas_from_boxes(p, np.zeros_like(gt))
    assert not z.any()                                                  # zero GT extent -> zero delta
    zp = bo.deltas_from_boxes(np.zeros((1, 4), np.float32), np.array([[0, 0, .5, .5]], np.float32))
    np.testing.assert_allclose(zp, [[250, 250, np.log(500), np.log(500)]], rtol=1e-6)   # 1e-3 guard


def test_match_encode_rules():
    p = np.array([[0, 0, .5, .5], [.5, .5, 1, 1], [0, 0, 1, 1]], np.float32)
    gt = np.array([[[0, 0, .5, .5], [0, 0, .5, .5], [0, 0, 0, 0]]], np.float32)     # two identical GTs -> tie
    lab = np.array([[7, 9, -1]], np.int32)
    d, oh, idx, best, l = bo.match_encode(p, gt, lab, 21, 0.5, VARIANCES, return_aux=True)
    assert idx[0].tolist() == [0, 0, 0]                                 # first maximum wins
    assert l[0].tolist() == [7, 0, 0]                                   # iou 0.25 / 0 are not > 0.5
    assert oh[0, 0, 7] == 1 and oh[0, 1, 0] == 1 and oh.sum() == 3
    assert not d[0, 1:].any() and not d[0, 0].any()                     # exact match -> zero deltas too
    # exactly 0.5 is NOT positive (strict >)
    gt2 = np.array([[[0, 0, .5, 1.0]]], np.float32)
    _, _, _, b2, l2 = bo.match_encode(p[2:], gt2, np.array([[3]], np.int32), 21, 0.5, VARIANCES, return_aux=True)
    assert b2[0, 0] == np.float32(0.5) and l2[0, 0] == 0


def test_loss_quirk_final_mask_two():
    # one positive, three anchors: 3*n_pos = 3 negatives requested but only two
    # anchors are negatives -> the zero-loss positive is also picked: mask 2 (ssd_loss.py:78-85)
    y = np.zeros((1, 3, 3), np.float32); y[0, 0, 1] = 1; y[0, 1:, 0] = 1
    p = np.array([[[.2, .7, .1], [.6, .3, .1], [.5, .25, .25]]], np.float32)
    loss, ce, final, rank = bo.conf_loss(y, p, 3.0, return_aux=True)
    assert final[0].tolist() == [2.0, 1.0, 1.0]
    assert rank[0].tolist() == [2, 1, 0]
    expect = (2 * -np.log(.7) - np.log(.6) - np.log(.5)) / 1
    assert loss[0] == pytest.approx(expect, rel=1e-6)
    # no positives: divide by 1, nothing selected
    y0 = np.zeros((1, 3, 3), np.float32); y0[..., 0] = 1
    assert bo.conf_loss(y0, p)[0] == 0.0
    assert bo.loc_loss(np.zeros((1, 3, 4), np.float32), np.ones((1, 3, 4), np.float32))[0] == 0.0


def test_loss_matches_torch_reference():
    rng = np.random.default_rng(1)
    B, N, L = 3, 200, 21
    lab = np.where(rng.random((B, N)) < 0.05, rng.integers(1, L, (B, N)), 0)
    y = np.eye(L, dtype=np.float32)[lab]
    z = rng.standard_normal((B, N, L)).astype(np.float32) * 2
    p = bo.softmax(z)
    ce_t = torch.nn.functional.cross_entropy(torch.from_numpy(z).reshape(-1, L), torch.from_numpy(lab).reshape(-1),
                                             reduction="none").reshape(B, N).numpy()
    np.testing.assert_allclose(bo.categorical_ce_from_probs(y, p), ce_t, rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(bo.categorical_ce_from_logits(y, z), ce_t, rtol=2e-5, atol=1e-6)
    ad = rng.standard_normal((B, N, 4)).astype(np.float32) * (lab > 0)[..., None]
    pd = rng.standard_normal((B, N, 4)).astype(np.float32) * 2
    hub = torch.nn.functional.huber_loss(torch.from_numpy(pd), torch.from_numpy(ad), reduction="none", delta=1.0)
    pos = torch.from_numpy((ad != 0).any(-1))
    ref = (hub.sum(-1) * pos).sum(-1) / pos.sum(-1).clamp(min=1)
    np.testing.assert_allclose(bo.loc_loss(ad, pd), ref.numpy(), rtol=1e-5)
    # hard negatives == the 3*n_pos largest background CE values (no ties in random data)
    out = bo.conf_loss(y, p)
    for b in range(B):
        npos = int((lab[b] > 0).sum())
        neg_ce = np.sort(np.where(lab[b] == 0, ce_t[b], 0))[::-1][:3 * npos]
        exp = (ce_t[b][lab[b] > 0].sum() + neg_ce.sum()) / max(npos, 1)
        assert out[b] == pytest.approx(exp, rel=1e-4)


def test_loss_grads_finite_difference():
    rng = np.random.default_rng(2)
    B, N, L = 2, 40, 5
    lab = np.where(rng.random((B, N)) < 0.2, rng.integers(1, L, (B, N)), 0)
    y = np.eye(L, dtype=np.float32)[lab]
    z = rng.standard_normal((B, N, L)).astype(np.float32)
    ad = (rng.standard_normal((B, N, 4)) * (lab > 0)[..., None]).astype(np.float32)
    pd = rng.standard_normal((B, N, 4)).astype(np.float32)
    gd, gz = bo.loss_grads(ad, pd, y, z)

    def total(pd_, z_):
        return float(bo.loc_loss(ad, pd_).astype(np.float64).mean() +
                     bo.conf_loss(y, z_, from_logits=True).astype(np.float64).mean())
    for (b, n, k) in [(0, 3, 1), (1, 7, 0), (0, 11, 3)]:
        e = np.zeros_like(z, dtype=np.float32); e[b, n, k] = 1e-2
        fd = (total(pd, z + e) - total(pd, z - e)) / 2e-2
        assert gz[b, n, k] == pytest.approx(fd, abs=2e-3)
    pn = np.argwhere(lab > 0)[0]
    e = np.zeros_like(pd); e[pn[0], pn[1], 2] = 1e-2
    fd = (total(pd + e, z) - total(pd - e, z)) / 2e-2
    assert gd[pn[0], pn[1], 2] == pytest.approx(fd, abs=2e-3)


def test_nms_against_torchvision_single_class():
    import torchvision
    rng = np.random.default_rng(3)
    N = 300
    c = rng.random((N, 2)); wh = rng.uniform(0.05, 0.3, (N, 2))
    boxes = np.concatenate([c - wh / 2, c + wh / 2], -1).astype(np.float32)        # y1 x1 y2 x2
    scores = rng.random(N).astype(np.float32)
    sc = np.zeros((1, N, 2), np.float32); sc[0, :, 1] = scores
    ob, os_, oc, valid, oi = bo.combined_nms(boxes.reshape(1, N, 1, 4), sc, N, N, 0.5, 0.3, clip_boxes=False,
                                             return_indices=True)
    keep = torchvision.ops.nms(torch.from_numpy(boxes[:, [1, 0, 3, 2]]), torch.from_numpy(scores), 0.5).numpy()
    keep = keep[scores[keep] > 0.3]
    assert valid[0] == len(keep)
    assert oi[0, :valid[0]].tolist() == keep.tolist()
    assert (oc[0, :valid[0]] == 1).all()


def test_decoder_background_rule_and_order():
    priors = np.array([[.1, .1, .3, .3], [.1, .1, .3, .3], [.6, .6, .9, .9], [.5, .5, .8, .8]], np.float32)
    deltas = np.zeros((1, 4, 4), np.float32)
    probs = np.array([[[.1, .9, 0], [.05, .15, .8], [.6, .4, 0], [.2, .1, .7]]], np.float32)
    b, l, s = bo.ssd_decode(priors, VARIANCES, deltas, probs, max_total_size=5)
    # anchor 2 has argmax 0 -> dropped; anchors 0 (class 1) and 1 (class 2) overlap but differ in class
    assert s[0].tolist() == pytest.approx([.9, .8, .7, 0, 0])
    assert l[0].tolist() == [1, 2, 2, 0, 0]
    np.testing.assert_allclose(b[0, 0], priors[0], atol=1e-7)
    assert not b[0, 3:].any()
