"""Parity against fixtures produced by the reference's OWN source.

``tests/golden/ref_*.npz`` were written by ``tests/golden/make_ref_golden.py``: the unmodified modules under
``/root/reference`` (utils/bbox_utils.py, utils/train_utils.py, ssd_loss.py, models/decoder.py, models/header.py,
models/ssd_vgg16.py, models/ssd_mobilenet_v2.py) executed on the NumPy-backed ``tensorflow`` stand-in in
``tests/tf_shim``.  Three kinds of test:

* CPU: the committed fixtures are reproducible from the reference (only where /root/reference exists);
* CPU: the oracle (``oracle/``) reproduces them -- bit-exact for priors, IoU, indices, one-hot targets, NMS selection;
* GPU: the CUDA path, through the C ABI, reproduces them -- bit-exact for integer/index results, 1e-4 relative for
  float32 box / loss tensors (north_star), fp16-storage yardstick for the networks.
"""

import os

import numpy as np
import pytest

from oracle import box_oracle as bo
from oracle import augment_oracle as ao
from oracle import net_oracle as no
from tests.golden import ref_inputs as ri

HERE = os.path.dirname(os.path.abspath(__file__))
REF = {k: np.load(os.path.join(HERE, "golden", f"ref_{k}.npz")) for k in ("priors", "box", "loss", "decode", "net", "augment")}
VAR = ri.VARIANCES
RTOL = 1e-4                  # north_star: box / loss tensors within 1e-4 relative float32
HAVE_REFERENCE = os.path.isdir(os.environ.get("SSD_REFERENCE_DIR", "/root/reference"))


def _targets():
    pri = REF["priors"]["priors_mobilenet_v2"]
    gt, lab = ri.ground_truth(3, 8, seed=11)
    assert np.array_equal(gt, REF["box"]["gt_rand"]) and np.array_equal(lab, REF["box"]["lab_rand"])
    return pri, gt, lab


# ------------------------------------------------------------ reproducibility --
@pytest.mark.skipif(not HAVE_REFERENCE, reason="/root/reference is not present on this machine")
@pytest.mark.parametrize("which", ["priors", "box", "loss", "decode", "augment"])
def test_fixtures_regenerate_from_the_reference_source(which):
    """Running the reference's modules under the shim again yields the committed bytes."""
    from tests.golden import make_ref_golden as gen
    fresh = gen.generate([which])[which]
    assert sorted(fresh) == sorted(REF[which].files)
    for k, v in fresh.items():
        assert v.dtype == REF[which][k].dtype and np.array_equal(v, REF[which][k], equal_nan=v.dtype.kind == "f"), k


def test_shim_is_independent_of_the_oracle_and_the_product():
    for dirpath, _, files in os.walk(os.path.join(HERE, "tf_shim")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "tf_ssd_b200" not in src.replace(
                    "``tf_ssd_b200``", ""), f
    src = open(os.path.join(HERE, "golden", "make_ref_golden.py")).read()
    assert "from oracle" not in src and "import oracle" not in src and "import tf_ssd_b200" not in src


# ------------------------------------------------------------- oracle (CPU) --
def test_oracle_priors_bit_exact():
    assert REF["priors"]["scale_k3"][0] == pytest.approx(0.48)          # the reference's own KAT (tests/test_bbox_utils.py:19-22)
    for name, (fm, ars) in ri.PRIOR_CONFIGS.items():
        assert np.array_equal(bo.prior_boxes(fm, ars), REF["priors"][f"priors_{name}"]), name
        for i in (0, len(fm) - 1):
            assert np.array_equal(bo.base_prior_boxes(ars[i], i + 1, len(fm)), REF["priors"][f"base_{name}_{i + 1}"])
    assert REF["priors"]["priors_vgg16"].shape == (8732, 4) and REF["priors"]["priors_vgg16_512"].shape == (24564, 4)


def test_oracle_iou_match_encode_bit_exact():
    X = REF["box"]
    pri = REF["priors"]["priors_mobilenet_v2"]
    for tag in ("rand", "snap"):
        gt, lab = X[f"gt_{tag}"], X[f"lab_{tag}"]
        assert np.array_equal(bo.iou_map(pri, gt), X[f"iou_{tag}"])
        d, oh = bo.match_encode(pri, gt, lab, 21, 0.5, VAR)
        assert np.array_equal(oh, X[f"onehot_{tag}"].astype(np.float32))
        assert np.array_equal(d, X[f"deltas_{tag}"])
    lat, gtt, labt = ri.tie_case()
    iou = bo.iou_map(lat, gtt)
    assert np.array_equal(iou, X["iou_tie"]) and (iou == 0.5).any()
    d, oh = bo.match_encode(lat, gtt, labt, 21, 0.5, VAR)
    assert np.array_equal(oh, X["onehot_tie"].astype(np.float32)) and np.array_equal(d, X["deltas_tie"])
    assert not (X["onehot_tie"][0].argmax(-1) == 9).any()          # the duplicated box never wins: first maximum
    assert np.array_equal(bo.iou_map(X["eval_boxes"], X["gt_rand"]), X["iou_eval"])
    assert np.array_equal(bo.iou_map(pri[:500], X["gt_rand"][0], transpose_perm=[1, 0]), X["iou_rank2"])
    assert np.array_equal(bo.iou_map(X["deg_a"], X["deg_g"]), X["iou_deg"], equal_nan=True)
    assert np.isnan(X["iou_deg"][0, 0, 0])                            # 0/0 (bbox_utils.py:55)
    assert np.array_equal(bo.deltas_from_boxes(X["enc_priors"], X["enc_gt"]), X["enc_deltas"])
    got = bo.boxes_from_deltas(pri, X["dec_in_deltas"] * np.array(VAR, np.float32))
    np.testing.assert_allclose(got, X["dec_boxes"], rtol=1e-6, atol=1e-7)


def test_oracle_losses():
    Ls = REF["loss"]
    pri, gt, lab = _targets()
    ad, al = bo.match_encode(pri, gt, lab, 21, 0.5, VAR)
    pd, probs = Ls["pred_deltas"], Ls["pred_probs"]
    y, p, ead, epd = ri.loss_edge_inputs()
    for ratio, alpha, tag in ((3.0, 1.0, "r3a1"), (2.0, 0.5, "r2a05")):
        for sfx in ("", "_tf20"):                                    # Huber mean*4 (TF >= 2.2) and element-wise (2.0/2.1)
            np.testing.assert_allclose(bo.loc_loss(ad, pd, alpha), Ls[f"loc_{tag}{sfx}"], rtol=1e-6)
            np.testing.assert_allclose(bo.loc_loss(ead, epd, alpha), Ls[f"edge_loc_{tag}{sfx}"], rtol=1e-6)
        np.testing.assert_allclose(bo.conf_loss(al, probs, ratio), Ls[f"conf_{tag}"], rtol=1e-6)
        np.testing.assert_allclose(bo.conf_loss(y, p, ratio), Ls[f"edge_conf_{tag}"], rtol=1e-6)
    assert Ls["rank_example"].tolist() == [[4, 1, 2, 5, 0, 3]]
    assert bo.hard_negative_rank(np.array([[0, .5, .5, 0, 2, .1]], np.float32)).tolist() == [[4, 1, 2, 5, 0, 3]]


def test_oracle_decoder_selection_exact():
    D = REF["decode"]
    for tag, cfg, B, seed, bg in ri.DECODE_CASES:
        pri = REF["priors"][f"priors_{cfg}"]
        pd, probs, _ = ri.head_outputs(B, pri.shape[0], seed=seed, background_bias=bg)
        b, l, s = bo.ssd_decode(pri, VAR, pd, probs)
        assert np.array_equal(l, D[f"{tag}_labels"]) and np.array_equal(s, D[f"{tag}_scores"]), tag
        np.testing.assert_allclose(b, D[f"{tag}_boxes"], rtol=1e-6, atol=1e-7)
    assert (D["mnv2_sparse_count"] < 200).all() and (D["mnv2_dense_count"] == 200).all()
    boxes, scores = ri.nms_tie_inputs()
    r = bo.combined_nms(boxes, scores, 10, 40, score_threshold=0.5)
    for got, k in zip(r, ("nms_boxes", "nms_scores", "nms_classes", "nms_valid")):
        assert np.array_equal(got, D[k]), k


@pytest.mark.parametrize("name", ["vgg16", "mobilenet_v2"])
def test_oracle_networks_match_the_reference_graphs(name):
    """Keras variable names / shapes and the float32 forward of the reference's model files."""
    from tf_ssd_b200.models.engine import SSDModel
    from tf_ssd_b200.utils import train_utils
    Nn = REF["net"]
    hp = train_utils.get_hyper_params(name)
    hp["total_labels"] = 21
    m = SSDModel(name, hp, seed=0)
    mine = sorted(f"{k}:{','.join(map(str, v.shape))}" for k, v in m.weights.items())
    assert mine == list(Nn[f"{name}_variables"])                      # same variables under the same Keras names
    w = ri.weights_for({k: v.shape for k, v in m.weights.items()})
    d, p = no.forward(name, w, hp, ri.image(1, 300), mode="fp32")
    rd, rp = Nn[f"{name}_deltas"], Nn[f"{name}_probs"]
    assert np.abs(d - rd).max() < 1e-4 * np.abs(rd).max() and np.abs(p - rp).max() < 1e-4


# ------------------------------------------------------------------ CUDA (GPU) --
# ------------------------------------------------------ augmentation.py (f3) --
AUG_TOL = 2e-6            # float32 images in [0,1]: device vs oracle, and operations without a mean vs the reference
# Operations that use a per-channel MEAN (expand's fill, contrast): the reference's float32 reduction order is
# TensorFlow's (NumPy's stand-in sums row by row and is itself ~1e-6 off the exact mean; the oracle and the device round
# the exact mean once); hue / saturation amplify that on near-grey pixels.  Measured: 7.9e-6 on case 0.
AUG_MEAN_TOL = 2e-5


def _aug_tol(i):
    case = ri.AUGMENT_CASES[i]
    uses_mean = case["contrast"] is not None or (case["patch"] is not None and case["patch"]["expand"] is not None)
    return AUG_MEAN_TOL if uses_mean else AUG_TOL


def _augment_plan(i):
    """The product's host planner fed with the samples the reference consumed for case ``i`` (and its crop window)."""
    from tf_ssd_b200 import augmentation as aug
    A, case = REF["augment"], ri.AUGMENT_CASES[i]
    H, W = A["img"].shape[:2]
    window = tuple(int(v) for v in A[f"case{i}_window"][:4])
    draws = aug.ReplayDraws(ri.augment_queue(case), [window] if case["patch"] is not None else [])
    plan = aug.make_plan(H, W, A["boxes"], draws)
    assert not draws.samples and not draws.crops, "the planner drew a different number of samples than the reference"
    if case["patch"] is not None and case["patch"]["expand"] is not None:
        g = plan["patch"]["expand"]
        assert (g["canvas_h"], g["canvas_w"]) == tuple(A[f"case{i}_window"][4:]), "canvas size differs from the reference's"
    return plan


def test_oracle_augmentation_matches_the_reference_source():
    """oracle/augment_oracle.py against augmentation.py:16-234 run unmodified on the shim (same draws, same window)."""
    A = REF["augment"]
    img, boxes = A["img"], A["boxes"]
    for i in range(len(ri.AUGMENT_CASES)):
        o_img, o_boxes = ao.apply(img, boxes, _augment_plan(i))
        assert o_img.shape == A[f"case{i}_img"].shape
        np.testing.assert_allclose(o_img, A[f"case{i}_img"], rtol=0, atol=_aug_tol(i), err_msg=f"case {i}")
        np.testing.assert_allclose(o_boxes, A[f"case{i}_boxes"], rtol=0, atol=1e-6, err_msg=f"case {i}")
    assert np.array_equal(ao.adjust_brightness(img, ao.uniform(A["op_brightness_u"][0], -0.12, 0.12)), A["op_brightness"])
    np.testing.assert_allclose(ao.adjust_contrast(img, ao.uniform(A["op_contrast_u"][0], 0.5, 1.5)), A["op_contrast"], rtol=0, atol=AUG_TOL)
    np.testing.assert_allclose(ao.adjust_hue(img, ao.uniform(A["op_hue_u"][0], -0.08, 0.08)), A["op_hue"], rtol=0, atol=AUG_TOL)
    np.testing.assert_allclose(ao.adjust_saturation(img, ao.uniform(A["op_saturation_u"][0], 0.5, 1.5)), A["op_saturation"], rtol=0, atol=AUG_TOL)
    assert np.array_equal(img[:, ::-1], A["op_flip_img"]) and np.array_equal(ao.flip_boxes(boxes), A["op_flip_boxes"])
    geom = ao.resolve_expand(img.shape[0], img.shape[1], *A["op_expand_u"])
    e_img, e_boxes = ao.expand_image(img, boxes, geom)
    assert e_img.shape == A["op_expand_img"].shape
    np.testing.assert_allclose(e_img, A["op_expand_img"], rtol=0, atol=AUG_TOL)
    assert np.array_equal(e_boxes, A["op_expand_boxes"])
    assert np.array_equal(ao.renormalize_boxes(boxes, [0.2, 0.1, 0.9, 0.6]), A["op_renorm"])


@pytest.mark.gpu
def test_cuda_augmentation_matches_the_reference_source():
    """ssd_augment_batch (all six cases as ONE batch, plus padded box rows) against the reference's outputs."""
    import torch
    from tf_ssd_b200 import augmentation as aug
    A = REF["augment"]
    n = len(ri.AUGMENT_CASES)
    img = np.repeat(A["img"][None], n, 0)
    boxes = np.zeros((n, 6, 4), np.float32)
    boxes[:, :4] = A["boxes"]
    plans = [_augment_plan(i) for i in range(n)]
    o_img, o_boxes = aug.apply_plans(img, boxes, plans)
    o_img, o_boxes = o_img.cpu().numpy(), o_boxes.cpu().numpy()
    for i in range(n):
        np.testing.assert_allclose(o_img[i], A[f"case{i}_img"], rtol=0, atol=_aug_tol(i), err_msg=f"case {i}")
        np.testing.assert_allclose(o_img[i], ao.apply(A["img"], A["boxes"], plans[i])[0], rtol=0, atol=AUG_TOL, err_msg=f"case {i}")
        np.testing.assert_allclose(o_boxes[i, :4], A[f"case{i}_boxes"], rtol=0, atol=1e-6, err_msg=f"case {i}")
        assert not o_boxes[i, 4:].any(), "padding rows must stay zero"
    # the reference's single operations (no clip), replaying the one sample each of them draws
    for name, fn in (("brightness", aug.random_brightness), ("contrast", aug.random_contrast), ("hue", aug.random_hue),
                     ("saturation", aug.random_saturation)):
        got, same_boxes = fn(A["img"], A["boxes"], draws=aug.ReplayDraws([A[f"op_{name}_u"][0]]))
        np.testing.assert_allclose(got.cpu().numpy(), A[f"op_{name}"], rtol=0, atol=AUG_TOL, err_msg=name)
        assert np.array_equal(same_boxes.cpu().numpy(), A["boxes"])
    f_img, f_boxes = aug.flip_horizontally(A["img"], A["boxes"])
    assert np.array_equal(f_img.cpu().numpy(), A["op_flip_img"]) and np.array_equal(f_boxes.cpu().numpy(), A["op_flip_boxes"])
    e_img, e_boxes = aug.expand_image(A["img"], A["boxes"], draws=aug.ReplayDraws(list(A["op_expand_u"])))
    assert tuple(e_img.shape) == A["op_expand_img"].shape
    np.testing.assert_allclose(e_img.cpu().numpy(), A["op_expand_img"], rtol=0, atol=AUG_TOL)
    assert np.array_equal(e_boxes.cpu().numpy(), A["op_expand_boxes"])


@pytest.mark.gpu
def test_cuda_augmentation_batch_against_the_oracle():
    """A 300x300 batch with random plans from the product's own sampler: device == oracle; deterministic."""
    import torch
    from tf_ssd_b200 import augmentation as aug
    rng = np.random.default_rng(3)
    B, S, G = 8, 300, 5
    img = (rng.integers(0, 256, (B, S, S, 3)).astype(np.float32) * np.float32(1 / 255.0))
    boxes, _ = ri.ground_truth(B, G, seed=21)
    draws = aug.RandomDraws(1234)
    plans = [aug.make_plan(S, S, boxes[i], draws) for i in range(B)]
    assert any(p["patch"] and p["patch"]["expand"] for p in plans) and any(p["patch"] is None for p in plans)
    o_img, o_boxes = aug.apply_plans(img, boxes, plans)
    again, _ = aug.apply_plans(img, boxes, plans)
    assert torch.equal(o_img, again), "augmentation must be deterministic"
    o_img, o_boxes = o_img.cpu().numpy(), o_boxes.cpu().numpy()
    for i in range(B):
        g = int((boxes[i] != 0).any(-1).sum())
        r_img, r_boxes = ao.apply(img[i], boxes[i, :g], plans[i])
        np.testing.assert_allclose(o_img[i], r_img, rtol=0, atol=AUG_TOL, err_msg=f"image {i}: {plans[i]}")
        np.testing.assert_allclose(o_boxes[i, :g], r_boxes, rtol=0, atol=1e-6)
        assert not o_boxes[i, g:].any()
    # the public call: reference signature, one example
    one_img, one_boxes = aug.apply(img[0], boxes[0, :2], draws=aug.RandomDraws(5))
    assert tuple(one_img.shape) == (S, S, 3) and tuple(one_boxes.shape) == (2, 4)
    assert float(one_img.min()) >= 0.0 and float(one_img.max()) <= 1.0


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.gpu
def test_cuda_priors_iou_match_encode():
    from tf_ssd_b200.utils import bbox_utils, train_utils
    X = REF["box"]
    for name, (fm, ars) in ri.PRIOR_CONFIGS.items():
        assert np.array_equal(_np(bbox_utils.generate_prior_boxes(fm, ars)), REF["priors"][f"priors_{name}"]), name
    pri = REF["priors"]["priors_mobilenet_v2"]
    hp = {"total_labels": 21, "iou_threshold": 0.5, "variances": VAR}
    cases = [("rand", pri, X["gt_rand"], X["lab_rand"]), ("snap", pri, X["gt_snap"], X["lab_snap"]),
             ("tie",) + ri.tie_case()]
    for tag, p, gt, lab in cases:
        assert np.array_equal(_np(bbox_utils.generate_iou_map(p, gt)), X[f"iou_{tag}"]), tag          # bit-exact
        d, oh = train_utils.calculate_actual_outputs(p, gt, lab, hp)
        assert np.array_equal(_np(oh), X[f"onehot_{tag}"].astype(np.float32)), tag                     # exact indices
        np.testing.assert_allclose(_np(d), X[f"deltas_{tag}"], rtol=RTOL, atol=1e-6)
    assert np.array_equal(_np(bbox_utils.generate_iou_map(X["eval_boxes"], X["gt_rand"])), X["iou_eval"])
    assert np.array_equal(_np(bbox_utils.generate_iou_map(pri[:500], X["gt_rand"][0], transpose_perm=[1, 0])), X["iou_rank2"])
    assert np.array_equal(_np(bbox_utils.generate_iou_map(X["deg_a"], X["deg_g"])), X["iou_deg"], equal_nan=True)
    np.testing.assert_allclose(_np(bbox_utils.get_deltas_from_bboxes(X["enc_priors"], X["enc_gt"])), X["enc_deltas"],
                               rtol=RTOL, atol=1e-6)
    import torch
    dd = torch.from_numpy(X["dec_in_deltas"] * np.array(VAR, np.float32))
    np.testing.assert_allclose(_np(bbox_utils.get_bboxes_from_deltas(pri, dd)), X["dec_boxes"], rtol=RTOL, atol=1e-6)


@pytest.mark.gpu
def test_cuda_losses():
    from tf_ssd_b200.ssd_loss import CustomLoss
    from tf_ssd_b200.utils import train_utils
    Ls = REF["loss"]
    pri, gt, lab = _targets()
    ad, al = train_utils.calculate_actual_outputs(pri, gt, lab, {"total_labels": 21, "iou_threshold": 0.5, "variances": VAR})
    pd, probs = Ls["pred_deltas"], Ls["pred_probs"]
    y, p, ead, epd = ri.loss_edge_inputs()
    for ratio, alpha, tag in ((3, 1, "r3a1"), (2, 0.5, "r2a05")):
        fn = CustomLoss(ratio, alpha)
        np.testing.assert_allclose(_np(fn.loc_loss_fn(ad, pd)), Ls[f"loc_{tag}"], rtol=RTOL)
        np.testing.assert_allclose(_np(fn.loc_loss_fn(ead, epd)), Ls[f"edge_loc_{tag}"], rtol=RTOL)
        np.testing.assert_allclose(_np(fn.conf_loss_fn(al, probs)), Ls[f"conf_{tag}"], rtol=RTOL)
        np.testing.assert_allclose(_np(fn.conf_loss_fn(y, p)), Ls[f"edge_conf_{tag}"], rtol=RTOL)


@pytest.mark.gpu
def test_cuda_decoder_selection_exact():
    from tf_ssd_b200.models.decoder import SSDDecoder
    from tf_ssd_b200.utils import bbox_utils
    D = REF["decode"]
    for tag, cfg, B, seed, bg in ri.DECODE_CASES:
        pri = REF["priors"][f"priors_{cfg}"]
        pd, probs, _ = ri.head_outputs(B, pri.shape[0], seed=seed, background_bias=bg)
        b, l, s = SSDDecoder(pri, VAR)([pd, probs])
        assert np.array_equal(_np(l), D[f"{tag}_labels"]) and np.array_equal(_np(s), D[f"{tag}_scores"]), tag   # order exact
        np.testing.assert_allclose(_np(b), D[f"{tag}_boxes"], rtol=1e-5, atol=1e-6)
    boxes, scores = ri.nms_tie_inputs()
    r = bbox_utils.non_max_suppression(boxes, scores, max_output_size_per_class=10, max_total_size=40, score_threshold=0.5)
    for got, k in zip(r, ("nms_boxes", "nms_scores", "nms_classes", "nms_valid")):
        assert np.array_equal(_np(got), D[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["vgg16", "mobilenet_v2"])
def test_cuda_networks_match_the_reference_graphs(name):
    """fp16-storage forward against the reference graph's float32 result: the distance must stay within 1.5x the
    distance of the oracle's fp16-storage simulation (the yardstick of what fp16 storage costs on this net)."""
    from tf_ssd_b200.models.engine import SSDModel
    from tf_ssd_b200.utils import train_utils
    import torch
    Nn = REF["net"]
    hp = train_utils.get_hyper_params(name)
    hp["total_labels"] = 21
    m = SSDModel(name, hp, seed=0)
    w = ri.weights_for({k: v.shape for k, v in m.weights.items()})
    m.set_weights(w)
    x = ri.image(1, 300)
    d, p = m(x)
    torch.cuda.synchronize()
    sd, sp = no.forward(name, w, hp, x, mode="fp16sim")
    rd, rp = Nn[f"{name}_deltas"], Nn[f"{name}_probs"]
    rms = lambda a, b: float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-12))
    assert rms(_np(d), rd) < 1.5 * rms(sd, rd) + 1e-3, (rms(_np(d), rd), rms(sd, rd))
    assert rms(_np(p), rp) < 1.5 * rms(sp, rp) + 1e-3, (rms(_np(p), rp), rms(sp, rp))
    assert (_np(p).argmax(-1) == rp.argmax(-1)).mean() > 0.9        # fp16 storage flips near-tied classes of a random-weight net
