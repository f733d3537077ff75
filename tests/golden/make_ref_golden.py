"""Writes ``tests/golden/ref_*.npz``: fixtures computed by the UNMODIFIED reference sources.

    python tests/golden/make_ref_golden.py            # needs /root/reference (this container only)

The reference's modules (``utils/bbox_utils.py``, ``utils/train_utils.py``, ``ssd_loss.py``, ``models/decoder.py``,
``models/header.py``, ``models/ssd_vgg16.py``, ``models/ssd_mobilenet_v2.py``) are imported from where they lie under
``/root/reference`` with ``tests/tf_shim`` first on ``sys.path``, so that ``import tensorflow`` resolves to the
NumPy-backed stand-in (same pattern as the reference's own ``tests/test_support.py:37-68``, except that the ops here
execute).  Nothing from ``oracle/`` is imported: these bytes are an independent witness for the oracle AND the CUDA
path.  What remains [TF-recall] is listed in the shim's docstrings and in DESIGN.md section 2.
"""

from __future__ import annotations

import importlib
import os
import sys
from typing import Any, Dict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get("SSD_REFERENCE_DIR", "/root/reference")
SHIM = os.path.join(ROOT, "tests", "tf_shim")

if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from tests.golden import ref_inputs as ri          # noqa: E402


class Reference(object):
    """The reference's modules, imported under the shim; ``sys.modules`` / ``sys.path`` are restored afterwards."""

    MODULES = ("utils", "utils.bbox_utils", "utils.train_utils", "ssd_loss", "models", "models.header",
               "models.decoder", "models.ssd_vgg16", "models.ssd_mobilenet_v2", "augmentation")

    def __init__(self):
        if not os.path.isdir(REFERENCE):
            raise FileNotFoundError(REFERENCE)
        saved_path = list(sys.path)
        saved = {k: v for k, v in sys.modules.items()
                 if k == "tensorflow" or k.startswith("tensorflow.") or k.split(".")[0] in ("utils", "models", "ssd_loss", "augmentation")}
        for k in saved:
            del sys.modules[k]
        sys.path[:0] = [SHIM, REFERENCE]
        try:
            self.tf = importlib.import_module("tensorflow")
            assert self.tf.__version__.endswith("numpy-shim")
            self.bbox_utils = importlib.import_module("utils.bbox_utils")
            self.train_utils = importlib.import_module("utils.train_utils")
            self.ssd_loss = importlib.import_module("ssd_loss")
            self.decoder = importlib.import_module("models.decoder")
            self.header = importlib.import_module("models.header")
            self.ssd_vgg16 = importlib.import_module("models.ssd_vgg16")
            self.ssd_mobilenet_v2 = importlib.import_module("models.ssd_mobilenet_v2")
            self.keras_layers = importlib.import_module("tensorflow.keras.layers")
            self.augmentation = importlib.import_module("augmentation")
            self.tf_image = importlib.import_module("tensorflow.image")
            for m in (self.bbox_utils, self.train_utils, self.ssd_loss, self.decoder, self.header, self.ssd_vgg16):
                assert os.path.abspath(m.__file__).startswith(os.path.abspath(REFERENCE)), m.__file__
        finally:
            sys.path[:] = saved_path
            for k in list(sys.modules):
                if k == "tensorflow" or k.startswith("tensorflow.") or k.split(".")[0] in ("utils", "models", "ssd_loss", "augmentation"):
                    del sys.modules[k]
            sys.modules.update(saved)


def _n(t: Any) -> np.ndarray:
    return np.array(t.numpy())


# ------------------------------------------------------------------ fixtures --
def priors_fixture(ref: Reference) -> Dict[str, np.ndarray]:
    """utils/bbox_utils.py:131-214 for the two reference configs and the 7-map SSD512 layout."""
    out = {"scale_k3": np.array([ref.bbox_utils.get_scale_for_nth_feature_map(3)], np.float64)}
    for name, (fm, ars) in ri.PRIOR_CONFIGS.items():
        out[f"priors_{name}"] = _n(ref.bbox_utils.generate_prior_boxes(fm, ars))
        for i in (0, len(fm) - 1):
            out[f"base_{name}_{i + 1}"] = _n(ref.bbox_utils.generate_base_prior_boxes(ars[i], i + 1, len(fm)))
    hp = ref.train_utils.get_hyper_params("vgg16")
    assert (hp["feature_map_shapes"], hp["aspect_ratios"]) == ri.PRIOR_CONFIGS["vgg16"]
    hp = ref.train_utils.get_hyper_params("mobilenet_v2")
    assert (hp["feature_map_shapes"], hp["aspect_ratios"]) == ri.PRIOR_CONFIGS["mobilenet_v2"]
    return out


def box_fixture(ref: Reference) -> Dict[str, np.ndarray]:
    """IoU (three shape modes), encode / decode, calculate_actual_outputs."""
    bu, tu = ref.bbox_utils, ref.train_utils
    out: Dict[str, np.ndarray] = {}
    priors = _n(bu.generate_prior_boxes(*ri.PRIOR_CONFIGS["mobilenet_v2"]))
    hp = tu.get_hyper_params("mobilenet_v2")
    hp["total_labels"] = 21
    for tag, snap in (("rand", 0), ("snap", 32)):
        gt, lab = ri.ground_truth(3, 8, seed=11, snap=snap)
        out[f"gt_{tag}"], out[f"lab_{tag}"] = gt, lab
        out[f"iou_{tag}"] = _n(bu.generate_iou_map(priors, gt))                      # [N,4] x [B,G,4] -> [B,N,G]  (train)
        d, oh = tu.calculate_actual_outputs(priors, gt, lab, hp)                    # train_utils.py:102-136
        out[f"deltas_{tag}"], out[f"onehot_{tag}"] = _n(d), _n(oh).astype(np.uint8)
    lattice, gt_tie, lab_tie = ri.tie_case()
    out["iou_tie"] = _n(bu.generate_iou_map(lattice, gt_tie))
    d, oh = tu.calculate_actual_outputs(lattice, gt_tie, lab_tie, hp)
    out["deltas_tie"], out["onehot_tie"] = _n(d), _n(oh).astype(np.uint8)
    # eval mode: [B,M,4] x [B,G,4] -> [B,M,G] (utils/eval_utils.py:57); rank-2 mode with transpose_perm=[1,0]
    rng = np.random.default_rng(5)
    pb = np.sort(rng.random((3, 40, 2, 2)).astype(np.float32), axis=2).transpose(0, 1, 3, 2).reshape(3, 40, 4)
    pb = pb[..., [0, 2, 1, 3]].copy()
    out["eval_boxes"] = pb
    out["iou_eval"] = _n(bu.generate_iou_map(pb, out["gt_rand"]))
    out["iou_rank2"] = _n(bu.generate_iou_map(priors[:500], out["gt_rand"][0], transpose_perm=[1, 0]))
    # degenerate boxes: 0/0 = NaN when both areas are zero (bbox_utils.py:55)
    deg_a = np.array([[0.2, 0.2, 0.2, 0.2], [0.1, 0.1, 0.4, 0.4]], np.float32)
    deg_g = np.array([[[0.2, 0.2, 0.2, 0.2], [0.0, 0.0, 0.0, 0.0], [0.1, 0.1, 0.4, 0.4]]], np.float32)
    out["deg_a"], out["deg_g"] = deg_a, deg_g
    out["iou_deg"] = _n(bu.generate_iou_map(deg_a, deg_g))
    # encode / decode on their own (bbox_utils.py:58-128), incl. zero-sized gt and zero-sized prior
    enc_p = np.concatenate([priors[::97], np.array([[0.3, 0.3, 0.3, 0.6], [0.5, 0.5, 0.5, 0.5]], np.float32)])
    enc_g = np.sort(rng.random((enc_p.shape[0], 2, 2)).astype(np.float32), axis=1).transpose(0, 2, 1).reshape(-1, 4)
    enc_g = enc_g[:, [0, 2, 1, 3]].copy()
    enc_g[3] = 0.0
    enc_g[5, 3] = enc_g[5, 1]            # zero width
    enc_g[6, 2] = enc_g[6, 0]            # zero height
    out["enc_priors"], out["enc_gt"] = enc_p, enc_g
    out["enc_deltas"] = _n(bu.get_deltas_from_bboxes(enc_p, enc_g))
    dd = rng.standard_normal((2, priors.shape[0], 4)).astype(np.float32)
    out["dec_in_deltas"] = dd
    out["dec_boxes"] = _n(bu.get_bboxes_from_deltas(priors, ref.tf.constant(dd) * ri.VARIANCES))
    return out


def loss_fixture(ref: Reference) -> Dict[str, np.ndarray]:
    """ssd_loss.py:26-91 on matched targets and on the mining corner cases; Huber under both TF reduction shapes."""
    bu, tu, tf = ref.bbox_utils, ref.train_utils, ref.tf
    out: Dict[str, np.ndarray] = {}
    priors = _n(bu.generate_prior_boxes(*ri.PRIOR_CONFIGS["mobilenet_v2"]))
    hp = tu.get_hyper_params("mobilenet_v2")
    hp["total_labels"] = 21
    gt, lab = ri.ground_truth(3, 8, seed=11)
    ad, al = tu.calculate_actual_outputs(priors, gt, lab, hp)
    pd, probs, _ = ri.head_outputs(3, priors.shape[0], seed=12)
    out["pred_deltas"], out["pred_probs"] = pd, probs
    y, p, ead, epd = ri.loss_edge_inputs()
    for ratio, alpha, tag in ((3, 1, "r3a1"), (2, 0.5, "r2a05")):
        loss = ref.ssd_loss.CustomLoss(ratio, alpha)
        for mean_mode in (True, False):
            tf.losses.Huber.mean_last_axis = mean_mode
            k = f"loc_{tag}" + ("" if mean_mode else "_tf20")
            out[k] = _n(loss.loc_loss_fn(ad, tf.constant(pd)))
            out[f"edge_{k}"] = _n(loss.loc_loss_fn(tf.constant(ead), tf.constant(epd)))
        tf.losses.Huber.mean_last_axis = True
        out[f"conf_{tag}"] = _n(loss.conf_loss_fn(al, tf.constant(probs)))
        out[f"edge_conf_{tag}"] = _n(loss.conf_loss_fn(tf.constant(y), tf.constant(p)))
    for a, b in (("loc_r3a1", "loc_r3a1_tf20"), ("edge_loc_r3a1", "edge_loc_r3a1_tf20")):
        assert np.allclose(out[a], out[b], rtol=1e-6), "Huber: mean*4 and sum disagree"
    # the documented rank example: argsort(argsort(x, DESC))
    x = tf.constant(np.array([[0, .5, .5, 0, 2, .1]], np.float32))
    out["rank_example"] = _n(tf.argsort(tf.argsort(x, direction="DESCENDING")))
    return out


def decode_fixture(ref: Reference) -> Dict[str, np.ndarray]:
    """models/decoder.py:60-93 (SSDDecoder.call) and bbox_utils.non_max_suppression."""
    bu, tf = ref.bbox_utils, ref.tf
    out: Dict[str, np.ndarray] = {}
    for name, cfg, B, seed, bg in ri.DECODE_CASES:
        priors = bu.generate_prior_boxes(*ri.PRIOR_CONFIGS[cfg])
        pd, probs, _ = ri.head_outputs(B, priors.shape[0], seed=seed, background_bias=bg)
        dec = ref.decoder.SSDDecoder(priors, ri.VARIANCES)
        b, l, s = dec.call([tf.constant(pd), tf.constant(probs)])
        out[f"{name}_boxes"], out[f"{name}_labels"], out[f"{name}_scores"] = _n(b), _n(l), _n(s)
        sc = _n(s)
        for i in range(B):                      # scores of one image are distinct: no equal-score order involved
            v = sc[i][sc[i] > 0]
            assert len(np.unique(v)) == len(v) and len(v) > 20, (name, i, len(v))
        out[f"{name}_count"] = (sc > 0).sum(-1).astype(np.int32)
    boxes, scores = ri.nms_tie_inputs()
    r = bu.non_max_suppression(tf.constant(boxes), tf.constant(scores), max_output_size_per_class=10,
                               max_total_size=40, score_threshold=0.5)
    out["nms_boxes"], out["nms_scores"], out["nms_classes"], out["nms_valid"] = [_n(t) for t in r]
    return out


def _load_named(model: Any, seed: int) -> Dict[str, tuple]:
    shapes = {k: tuple(v.shape) for k, v in model.named_variables().items()}
    w = ri.weights_for(shapes, seed)
    for k, v in model.named_variables().items():
        v.assign(w[k])
    return shapes


def net_fixture(ref: Reference) -> Dict[str, np.ndarray]:
    """models/ssd_vgg16.py:66-121, models/ssd_mobilenet_v2.py:15-47, models/header.py:54-90 run as written on the
    Keras stand-in (MobileNetV2's backbone constructor is third-party: [TF-recall], see the shim)."""
    out: Dict[str, np.ndarray] = {}
    for name, mod in (("vgg16", ref.ssd_vgg16), ("mobilenet_v2", ref.ssd_mobilenet_v2)):
        ref.keras_layers.reset_name_counters()
        hp = ref.train_utils.get_hyper_params(name)
        hp["total_labels"] = 21
        model = mod.get_model(hp)
        x = ri.image(1, hp["img_size"])
        model(x)                                            # init_model's job: builds the variables
        shapes = _load_named(model, ri.NET_SEED)
        deltas, probs = model(x)
        out[f"{name}_deltas"], out[f"{name}_probs"] = _n(deltas), _n(probs)
        out[f"{name}_variables"] = np.array(sorted(f"{k}:{','.join(map(str, s))}" for k, s in shapes.items()))
        taps = model.last_activations
        tap_names = {"vgg16": ["l2_normalization", "conv7", "conv8_2", "conv9_2", "conv10_2", "conv11_2"],
                     "mobilenet_v2": ["block_13_expand_relu", "out_relu", "extra1_2", "extra2_2", "extra3_2", "extra4_2"]}[name]
        out[f"{name}_tap_shapes"] = np.array([taps[t].shape for t in tap_names], np.int32)
        out[f"{name}_tap_absmean"] = np.array([np.abs(taps[t].numpy()).mean(dtype=np.float64) for t in tap_names])
        out[f"{name}_tap5"] = _n(taps[tap_names[4]])
    return out


def augment_fixture(ref: Reference) -> Dict[str, np.ndarray]:
    """augmentation.py:16-234 (``apply`` and each operation on its own) with the random stream and the crop window of
    ``sample_distorted_bounding_box`` fed in; utils/bbox_utils.py:217-233 through ``patch`` / ``expand_image``."""
    aug, tf, tfi = ref.augmentation, ref.tf, ref.tf_image
    img, boxes = ri.augment_inputs()
    out: Dict[str, np.ndarray] = {"img": img, "boxes": boxes}
    for i, case in enumerate(ri.AUGMENT_CASES):
        tf.random.queue[:] = ri.augment_queue(case)
        tf.random.log.clear()
        tfi.crop_queue[:] = [case["patch"]["crop"]] if case["patch"] is not None else []
        tfi.crop_log.clear()
        o_img, o_boxes = aug.apply(tf.constant(img), tf.constant(boxes))
        assert not tf.random.queue and not tfi.crop_queue, "the reference drew a different number of samples"
        out[f"case{i}_img"], out[f"case{i}_boxes"] = _n(o_img), _n(o_boxes)
        out[f"case{i}_window"] = np.array(tfi.crop_log[0] if tfi.crop_log else (0, 0, 0, 0, 0, 0), np.int32)
        out[f"case{i}_draws"] = np.array([float(np.asarray(v)) for v in tf.random.log], np.float64)
    # single operations (augmentation.py:67-139, 164-202) with their draw
    for name, fn, u in (("brightness", aug.random_brightness, 0.9), ("contrast", aug.random_contrast, 0.2),
                        ("hue", aug.random_hue, 0.35), ("saturation", aug.random_saturation, 0.77)):
        tf.random.queue[:] = [u]
        o_img, _ = fn(tf.constant(img), tf.constant(boxes))
        out[f"op_{name}"] = _n(o_img)
        out[f"op_{name}_u"] = np.array([u], np.float64)
    o_img, o_boxes = aug.flip_horizontally(tf.constant(img), tf.constant(boxes))
    out["op_flip_img"], out["op_flip_boxes"] = _n(o_img), _n(o_boxes)
    tf.random.queue[:] = [0.5, 0.25, 0.75]
    H, W = img.shape[:2]
    o_img, o_boxes = aug.expand_image(tf.constant(img), tf.constant(boxes), tf.constant(np.float32(H)), tf.constant(np.float32(W)))
    out["op_expand_img"], out["op_expand_boxes"] = _n(o_img), _n(o_boxes)
    out["op_expand_u"] = np.array([0.5, 0.25, 0.75], np.float64)
    out["op_renorm"] = _n(ref.bbox_utils.renormalize_bboxes_with_min_max(tf.constant(boxes), tf.constant(np.array([0.2, 0.1, 0.9, 0.6], np.float32))))
    return out


FIXTURES = {"augment": augment_fixture, "priors": priors_fixture, "box": box_fixture, "loss": loss_fixture, "decode": decode_fixture,
            "net": net_fixture}


def generate(which=None) -> Dict[str, Dict[str, np.ndarray]]:
    ref = Reference()
    return {k: fn(ref) for k, fn in FIXTURES.items() if which is None or k in which}


def main() -> None:
    for k, d in generate(sys.argv[1:] or None).items():
        path = os.path.join(HERE, f"ref_{k}.npz")
        np.savez_compressed(path, **d)
        print(f"{path}: {len(d)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
