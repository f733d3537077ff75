"""Regenerates ``tests/golden/hot_path_golden.npz`` from the CPU oracle.

    python tests/golden/make_golden.py

The reference (TensorFlow 2.0) cannot be imported in this environment, so these
vectors pin the ORACLE's behaviour (and, through ``tests/test_golden.py``, the
CUDA path's) at a fixed point in time: any later change of either shows up as a
diff against committed bytes.  Inputs are seeded; sizes are small enough for the
file to stay a few hundred kB.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import box_oracle as bo          # noqa: E402
from tf_ssd_b200 import synth                # noqa: E402   (pure NumPy input generators)

AR3 = [1., 2., 1. / 2.]
AR5 = [1., 2., 1. / 2., 3., 1. / 3.]
FM = [19, 10, 5, 3, 2, 1]
ARS = [AR3, AR5, AR5, AR5, AR3, AR3]
VAR = [0.1, 0.1, 0.2, 0.2]


def main():
    out = {}
    priors = bo.prior_boxes(FM, ARS)                                   # utils/bbox_utils.py:179-214
    out["priors_mnv2"] = priors
    pv = bo.prior_boxes([38, 19, 10, 5, 3, 1], ARS)
    out["priors_vgg16_head"] = pv[:64]
    out["priors_vgg16_tail"] = pv[-64:]
    out["priors_vgg16_sum64"] = np.array([pv.astype(np.float64).sum()])

    for tag, snap in (("rand", None), ("snap", 32)):                   # snap: exact IoU ties / IoU == 0.5
        gt, lab = synth.make_ground_truth(3, padded=8, seed=11, snap=snap)
        out[f"gt_{tag}"], out[f"lab_{tag}"] = gt, lab
        out[f"iou_{tag}"] = bo.iou_map(priors, gt)                     # utils/bbox_utils.py:24-55
        d, oh = bo.match_encode(priors, gt, lab, 21, 0.5, VAR)         # utils/train_utils.py:102-136
        out[f"deltas_{tag}"], out[f"label_{tag}"] = d, oh.argmax(-1).astype(np.int32)

    # integer-tie case: anchors AND ground truth on the k/4 lattice -> IoU exactly 0.5, exact ties,
    # duplicated ground-truth boxes (first maximum must win), zero-area and padded boxes
    grid = [k / 4.0 for k in range(5)]
    lattice = np.array([[y1, x1, y2, x2] for y1 in grid for y2 in grid if y2 > y1
                        for x1 in grid for x2 in grid if x2 > x1], np.float32)            # 100 boxes
    gt_tie = np.zeros((2, 6, 4), np.float32)
    gt_tie[0, :5] = [[0, 0, .5, .5], [0, 0, .5, 1], [0, 0, .5, .5], [.25, .25, .75, .75], [.5, .5, .5, 1]]
    gt_tie[1, :3] = [[0, 0, 1, 1], [0, .5, 1, 1], [0, 0, 1, .5]]
    lab_tie = np.array([[3, 7, 9, 1, 2, -1], [5, 6, 4, -1, -1, -1]], np.int32)
    out["priors_tie"], out["gt_tie"], out["lab_tie"] = lattice, gt_tie, lab_tie
    out["iou_tie"] = bo.iou_map(lattice, gt_tie)
    d, oh = bo.match_encode(lattice, gt_tie, lab_tie, 21, 0.5, VAR)
    out["deltas_tie"], out["label_tie"] = d, oh.argmax(-1).astype(np.int32)

    pd, z = synth.make_head_outputs(3, priors.shape[0], 21, seed=12)
    out["pred_deltas"], out["pred_logits"] = pd, z
    p = bo.softmax(z)
    gt, lab = out["gt_rand"], out["lab_rand"]
    ad, al = bo.match_encode(priors, gt, lab, 21, 0.5, VAR)
    out["loc_loss"] = bo.loc_loss(ad, pd, 1.0)                         # ssd_loss.py:26-57
    out["conf_loss_probs"] = bo.conf_loss(al, p, 3.0, from_logits=False)   # ssd_loss.py:59-91
    out["conf_loss_logits"] = bo.conf_loss(al, z, 3.0, from_logits=True)
    b, l, s = bo.ssd_decode(priors, VAR, pd, p)                        # models/decoder.py:60-93
    out["dec_boxes"], out["dec_labels"], out["dec_scores"] = b, l, s
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hot_path_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
