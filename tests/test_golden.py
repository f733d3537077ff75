"""Committed golden vectors (tests/golden/hot_path_golden.npz, made by
tests/golden/make_golden.py): the oracle must keep reproducing them (CPU), and
the CUDA path must reproduce them through the C ABI (GPU)."""

import os

import numpy as np
import pytest

from oracle import box_oracle as bo

from tests.conftest import CONFIGS, VARIANCES

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hot_path_golden.npz"))
FM, ARS, _ = CONFIGS["mobilenet_v2"]


def test_oracle_reproduces_golden():
    priors = bo.prior_boxes(FM, ARS)
    assert np.array_equal(priors, GOLD["priors_mnv2"])
    pv = bo.prior_boxes(*CONFIGS["vgg16"][:2])
    assert np.array_equal(pv[:64], GOLD["priors_vgg16_head"]) and np.array_equal(pv[-64:], GOLD["priors_vgg16_tail"])
    assert abs(pv.astype(np.float64).sum() - GOLD["priors_vgg16_sum64"][0]) < 1e-9
    assert abs(GOLD["priors_vgg16_sum64"][0] - 17463.999998) < 1e-3          # SURVEY 8(c) KAT
    for tag in ("rand", "snap"):
        gt, lab = GOLD[f"gt_{tag}"], GOLD[f"lab_{tag}"]
        assert np.array_equal(bo.iou_map(priors, gt), GOLD[f"iou_{tag}"])
        d, oh = bo.match_encode(priors, gt, lab, 21, 0.5, VARIANCES)
        assert np.array_equal(d, GOLD[f"deltas_{tag}"]) and np.array_equal(oh.argmax(-1), GOLD[f"label_{tag}"])
    # the lattice case really contains exact ties and IoU == 0.5 (strict ">" and "first maximum wins" matter)
    pt, gtt, labt = GOLD["priors_tie"], GOLD["gt_tie"], GOLD["lab_tie"]
    iou = bo.iou_map(pt, gtt)
    assert np.array_equal(iou, GOLD["iou_tie"])
    assert (iou == 0.5).any()
    srt = np.sort(iou, axis=-1)
    assert ((srt[..., -1] == srt[..., -2]) & (srt[..., -1] > 0.5)).any()
    d, oh = bo.match_encode(pt, gtt, labt, 21, 0.5, VARIANCES)
    assert np.array_equal(d, GOLD["deltas_tie"]) and np.array_equal(oh.argmax(-1), GOLD["label_tie"])
    assert not (GOLD["label_tie"][0] == 9).any()            # box 2 duplicates box 0: the first one wins
    pd, z = GOLD["pred_deltas"], GOLD["pred_logits"]
    ad, al = bo.match_encode(priors, GOLD["gt_rand"], GOLD["lab_rand"], 21, 0.5, VARIANCES)
    assert np.allclose(bo.loc_loss(ad, pd, 1.0), GOLD["loc_loss"], rtol=1e-6)
    assert np.allclose(bo.conf_loss(al, bo.softmax(z), 3.0), GOLD["conf_loss_probs"], rtol=1e-6)
    assert np.allclose(bo.conf_loss(al, z, 3.0, from_logits=True), GOLD["conf_loss_logits"], rtol=1e-6)
    b, l, s = bo.ssd_decode(priors, VARIANCES, pd, bo.softmax(z))
    assert np.array_equal(l, GOLD["dec_labels"]) and np.array_equal(s, GOLD["dec_scores"])
    assert np.allclose(b, GOLD["dec_boxes"], rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
def test_cuda_reproduces_golden():
    import torch
    from tf_ssd_b200.models.decoder import SSDDecoder
    from tf_ssd_b200.ssd_loss import CustomLoss
    from tf_ssd_b200.utils import bbox_utils, train_utils
    n = lambda t: t.detach().cpu().numpy()
    priors = bbox_utils.generate_prior_boxes(FM, ARS)
    assert np.array_equal(n(priors), GOLD["priors_mnv2"])                     # bit-exact
    hp = {"total_labels": 21, "iou_threshold": 0.5, "variances": VARIANCES}
    for tag in ("rand", "snap"):
        gt, lab = GOLD[f"gt_{tag}"], GOLD[f"lab_{tag}"]
        assert np.array_equal(n(bbox_utils.generate_iou_map(priors, gt)), GOLD[f"iou_{tag}"])      # bit-exact
        d, oh, li, _ = train_utils.calculate_actual_outputs(priors, gt, lab, hp, return_indices=True)
        assert np.array_equal(n(li), GOLD[f"label_{tag}"])                    # exact indices / positive mask
        assert np.array_equal(n(oh).argmax(-1), GOLD[f"label_{tag}"])
        assert np.allclose(n(d), GOLD[f"deltas_{tag}"], rtol=1e-4, atol=1e-6)  # logf: 1e-4 rel (north_star)
    pt, gtt, labt = GOLD["priors_tie"], GOLD["gt_tie"], GOLD["lab_tie"]
    assert np.array_equal(n(bbox_utils.generate_iou_map(pt, gtt)), GOLD["iou_tie"])
    d, oh, li, _ = train_utils.calculate_actual_outputs(pt, gtt, labt, hp, return_indices=True)
    assert np.array_equal(n(li), GOLD["label_tie"]) and np.allclose(n(d), GOLD["deltas_tie"], rtol=1e-4, atol=1e-6)
    pd, z = GOLD["pred_deltas"], GOLD["pred_logits"]
    ad, al = train_utils.calculate_actual_outputs(priors, GOLD["gt_rand"], GOLD["lab_rand"], hp)
    loss = CustomLoss(3, 1)
    assert np.allclose(n(loss.loc_loss_fn(ad, pd)), GOLD["loc_loss"], rtol=1e-4)
    probs = torch.softmax(torch.from_numpy(z), -1).numpy()
    assert np.allclose(n(loss.conf_loss_fn(al, probs)), GOLD["conf_loss_probs"], rtol=1e-4)
    b, l, s = SSDDecoder(priors, VARIANCES)([pd, bo.softmax(z)])
    assert np.array_equal(n(l), GOLD["dec_labels"]) and np.array_equal(n(s), GOLD["dec_scores"])   # selection order exact
    assert np.allclose(n(b), GOLD["dec_boxes"], rtol=1e-5, atol=1e-6)
