"""dataset / io helpers (CPU), GPU evaluation (mAP) and the trainer.py / predictor.py flows."""

import os

import numpy as np
import pytest

from oracle import box_oracle as bo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_data_utils_contract():
    from tf_ssd_b200.utils import data_utils
    ds, info = data_utils.get_dataset("voc/2007", "test", total_items=10)
    assert data_utils.get_total_item_size(info, "test") == 10 and len(data_utils.get_labels(info)) == 20
    batches = list(ds.padded_batch(4, padded_shapes=data_utils.get_data_shapes(), padding_values=data_utils.get_padding_values()))
    assert [b[0].shape[0] for b in batches] == [4, 4, 2]
    img, gb, gl = batches[0]
    assert img.dtype == np.float32 and img.shape[1:] == (300, 300, 3) and 0.0 <= img.min() and img.max() <= 1.0
    assert gb.shape[:2] == gl.shape and gb.shape[2] == 4 and gl.dtype == np.int32
    pad = gl == -1
    assert pad.any() and not gb[pad].any()                               # padded boxes are zero, labels -1
    assert ((gl[~pad] >= 1) & (gl[~pad] <= 20)).all()                    # labels shifted by one: 0 is background
    assert (gb[~pad][:, 2] > gb[~pad][:, 0]).all() and (gb[~pad][:, 3] > gb[~pad][:, 1]).all()
    again = list(data_utils.get_dataset("voc/2007", "test", total_items=10)[0].padded_batch(4))
    assert np.array_equal(again[0][1], gb)                               # seeded: the same stream every time
    both = ds.concatenate(data_utils.get_dataset("voc/2012", "train+validation", total_items=3)[0])
    assert sum(b[0].shape[0] for b in both.padded_batch(5)) == 13
    assert len(list(ds.padded_batch(4, drop_remainder=True))) == 2
    # the default split sizes are the TFDS ones the reference trains on (trainer.py:47-58)
    assert data_utils.get_total_item_size(data_utils.get_dataset("voc/2007", "train+validation")[1], "train+validation") == 5011


def test_io_utils_and_checkpoint(tmp_path):
    from tf_ssd_b200.models.train_engine import ModelCheckpoint
    from tf_ssd_b200.utils import io_utils
    a = io_utils.handle_args(["--backbone", "vgg16", "-handle-gpu"])
    assert a.backbone == "vgg16" and a.handle_gpu
    assert io_utils.handle_args([]).backbone == "mobilenet_v2"
    io_utils.is_valid_backbone("mobilenet_v2")
    with pytest.raises(AssertionError):
        io_utils.is_valid_backbone("resnet")
    p = io_utils.get_model_path("vgg16", str(tmp_path / "trained"))
    assert p.endswith("ssd_vgg16_model_weights.npz") and os.path.isdir(os.path.dirname(p))
    assert io_utils.get_log_path("vgg16").startswith("logs/vgg16/")

    class FakeModel(object):
        saved = 0

        def save_weights(self, path):
            self.saved += 1
    cb = ModelCheckpoint(p, monitor="val_loss", save_best_only=True)
    m = FakeModel()
    cb.set_model(m)
    for epoch, v in enumerate([3.0, 2.0, 2.5, 1.0]):
        cb.on_epoch_end(epoch, {"val_loss": v})
    assert m.saved == 3 and cb.saved_epochs == [0, 1, 3] and cb.best == 1.0


def test_average_precision_arithmetic():
    from tf_ssd_b200.utils import eval_utils
    stats = eval_utils.init_stats(["bg", "a", "b"])
    assert sorted(stats) == [1, 2]
    stats[1].update(total=2, tp=[1, 0, 1], fp=[0, 1, 0], scores=[0.9, 0.8, 0.7])
    stats[2].update(total=1, tp=[0], fp=[1], scores=[0.6])
    stats, m_ap = eval_utils.calculate_mAP(stats)
    # class 1: precision [1, .5, 2/3], recall [.5, .5, 1] -> 11-point AP = (6*1 + 5*(2/3))/11
    assert stats[1]["AP"] == pytest.approx((6 + 5 * 2 / 3) / 11)
    assert stats[2]["AP"] == 0.0 and m_ap == pytest.approx(stats[1]["AP"] / 2)


def test_reference_kats_of_the_driver_modules(tmp_path, monkeypatch):
    """The known-answer tests the reference keeps for these modules (tests/test_eval_utils.py:110-147,
    tests/test_io_utils.py:17-57, tests/test_data_utils.py:31-71 of FurkanOM/tf-ssd), restated on this package."""
    import types
    from datetime import datetime
    from tf_ssd_b200.utils import data_utils, eval_utils, io_utils
    # eval_utils
    stats = eval_utils.init_stats(["bg", "person", "car"])
    assert sorted(stats.keys()) == [1, 2] and stats[1]["label"] == "person" and stats[2]["total"] == 0
    ap = eval_utils.calculate_ap(np.array([0.25, 0.5, 0.75, 1.0]), np.array([1.0, 0.75, 0.5, 0.25]))
    assert ap == pytest.approx(0.6363636364)
    st = {1: {"label": "person", "total": 1, "tp": [1, 0], "fp": [0, 1], "scores": [0.9, 0.2]},
          2: {"label": "car", "total": 1, "tp": [0, 1], "fp": [1, 0], "scores": [0.1, 0.8]}}
    out, m_ap = eval_utils.calculate_mAP(st)
    assert out[1]["AP"] == pytest.approx(1.0) and out[2]["AP"] == pytest.approx(1.0) and m_ap == pytest.approx(1.0)
    assert list(out[2]["recall"]) == [1.0, 1.0]
    # io_utils

    class FixedDateTime:
        @staticmethod
        def now():
            return datetime(2020, 1, 2, 3, 4, 5)
    monkeypatch.setattr(io_utils, "datetime", FixedDateTime)
    assert io_utils.get_log_path("mobilenet_v2", custom_postfix="_debug") == "logs/mobilenet_v2_debug/20200102-030405"
    monkeypatch.chdir(tmp_path)
    path = io_utils.get_model_path("vgg16")
    assert os.path.isdir(os.path.join(str(tmp_path), "trained"))
    assert path == "trained/ssd_vgg16_model_weights.npz"             # the reference writes .h5 (no h5py here)
    monkeypatch.setattr("sys.argv", ["prog", "-handle-gpu", "--backbone", "vgg16"])
    args = io_utils.handle_args()
    assert args.handle_gpu and args.backbone == "vgg16"
    io_utils.is_valid_backbone("mobilenet_v2")
    with pytest.raises(AssertionError):
        io_utils.is_valid_backbone("resnet50")
    # data_utils
    info = types.SimpleNamespace(splits={"train": types.SimpleNamespace(num_examples=8),
                                         "validation": types.SimpleNamespace(num_examples=3),
                                         "test": types.SimpleNamespace(num_examples=5)},
                                 features={"labels": types.SimpleNamespace(names=["person", "car", "dog"])})
    assert data_utils.get_total_item_size(info, "train+validation") == 11 and data_utils.get_total_item_size(info, "test") == 5
    assert data_utils.get_labels(info) == ["person", "car", "dog"]
    top = [str(tmp_path / "first.jpg"), str(tmp_path / "second.png")]
    os.makedirs(str(tmp_path / "nested"))
    for f in top + [str(tmp_path / "nested" / "ignored.jpg")]:
        open(f, "w").write("placeholder")
    assert sorted(p for p in data_utils.get_custom_imgs(str(tmp_path)) if not p.endswith("trained")) == sorted(top)
    assert data_utils.get_data_types() == ("float32", "float32", "int32")
    assert data_utils.get_data_shapes() == ([None, None, None], [None, None], [None])
    assert data_utils.get_padding_values() == (0, 0, -1)


def _update_stats_reference(pb, pl, ps, gb, gl, stats):
    """utils/eval_utils.py:36-91 restated with the oracle's IoU map."""
    iou = bo.iou_map(pb, gb)
    merged, gt_of = iou.max(-1), iou.argmax(-1)
    order = np.argsort(-merged, axis=-1, kind="stable")
    for lab in gl.reshape(-1):
        if lab != -1:
            stats[int(lab)]["total"] += 1
    for b in range(pb.shape[0]):
        taken = []
        for m in order[b]:
            if pl[b, m] == 0:
                continue
            lab, g = int(pl[b, m]), int(gt_of[b, m])
            stats[lab]["scores"].append(float(ps[b, m]))
            ok = merged[b, m] >= 0.5 and lab == int(gl[b, g]) and g not in taken
            stats[lab]["tp"].append(int(ok)); stats[lab]["fp"].append(int(not ok))
            if ok:
                taken.append(g)
    return stats


@pytest.mark.gpu
def test_update_stats_matches_reference_loops():
    from tf_ssd_b200 import synth
    from tf_ssd_b200.utils import eval_utils
    rng = np.random.default_rng(3)
    B, M = 3, 40
    gb, gl = synth.make_ground_truth(B, padded=6, max_boxes=5, seed=9)
    pb = np.zeros((B, M, 4), np.float32); pl = np.zeros((B, M), np.float32); ps = np.zeros((B, M), np.float32)
    for b in range(B):
        n = int((gl[b] > 0).sum())
        for m in range(30):                                               # jittered copies of the ground truth + noise
            g = m % n
            pb[b, m] = np.clip(gb[b, g] + rng.normal(0, 0.04, 4), 0, 1)
            pl[b, m] = gl[b, g] if m % 3 else (gl[b, g] % 20) + 1
            ps[b, m] = rng.random()
    labels = ["bg"] + [str(i) for i in range(1, 21)]
    got = eval_utils.update_stats(pb, pl, ps, gb, gl, eval_utils.init_stats(labels))
    ref = _update_stats_reference(pb, pl, ps, gb, gl, eval_utils.init_stats(labels))
    for k in ref:
        assert got[k]["total"] == ref[k]["total"] and got[k]["tp"] == ref[k]["tp"] and got[k]["fp"] == ref[k]["fp"], k
        assert np.allclose(got[k]["scores"], ref[k]["scores"])
    assert sum(sum(v["tp"]) for v in ref.values()) > 0


@pytest.mark.gpu
def test_trainer_and_predictor_flows(tmp_path):
    import predictor
    import trainer
    argv = ["--backbone", "mobilenet_v2", "--epochs", "2", "--batch-size", "8", "--train-items", "32", "--val-items", "16",
            "--model-dir", str(tmp_path)]
    hist = trainer.main(argv)
    assert len(hist["loss"]) == 2 and np.isfinite(hist["loss"]).all() and np.isfinite(hist["val_loss"]).all()
    path = os.path.join(str(tmp_path), "ssd_mobilenet_v2_model_weights.npz")
    assert os.path.exists(path)
    with np.load(path) as z:
        assert "block_13_expand/kernel" in z.files and "bn_Conv1/moving_mean" in z.files and z["Conv1/kernel"].shape == (3, 3, 3, 32)
    stats = predictor.main(argv)
    assert sorted(stats) == list(range(1, 21)) and all("AP" in v for v in stats.values())


def test_reference_kats_of_train_utils():
    """tests/test_train_utils.py:27-60 of the reference, restated on the package's host-side mirror."""
    from tf_ssd_b200.utils import train_utils
    params = train_utils.get_hyper_params("vgg16")
    assert params["img_size"] == 300 and params["iou_threshold"] == 0.5 and params["neg_pos_ratio"] == 3
    assert params["loc_loss_alpha"] == 1 and params["variances"] == [0.1, 0.1, 0.2, 0.2]
    over = train_utils.get_hyper_params("mobilenet_v2", img_size=320, iou_threshold=0.6)
    assert over["img_size"] == 320 and over["iou_threshold"] == 0.6
    assert train_utils.get_hyper_params("vgg16", img_size=0)["img_size"] == 300          # falsy overrides are ignored
    params["img_size"] = 320
    assert train_utils.SSD["vgg16"]["img_size"] == 300                                  # a copy, not the table itself
    assert train_utils.scheduler(0) == 1e-3 and train_utils.scheduler(110) == 1e-4 and train_utils.scheduler(130) == 1e-5
    assert train_utils.get_step_size(10, 4) == 3 and train_utils.get_step_size(8, 4) == 2
    assert train_utils.SSD["mobilenet_v2"]["feature_map_shapes"] == [19, 10, 5, 3, 2, 1]
    assert train_utils.SSD["vgg16"]["feature_map_shapes"] == [38, 19, 10, 5, 3, 1]


def test_preprocess_oracle_against_torch_interpolate():
    """The oracle's resize (TF half-pixel bilinear) agrees with torch's align_corners=False bilinear (same sampling rule)."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(4)
    for (H, W, S) in [(375, 500, 300), (120, 90, 300), (512, 512, 512), (333, 77, 64)]:
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        got = bo.preprocess_image(img, S, S)
        ref = F.interpolate(torch.from_numpy(img.astype(np.float32) / 255).permute(2, 0, 1)[None], size=(S, S), mode="bilinear",
                            align_corners=False, antialias=False)[0].permute(1, 2, 0).numpy()
        assert got.shape == (S, S, 3) and np.abs(got - ref).max() < 1e-5      # interpolation weights are rounded differently
        assert np.array_equal(bo.preprocess_image(img, S, S, flip=True), got[:, ::-1])
    b = np.array([[0.1, 0.2, 0.5, 0.6], [0, 0, 0, 0]], np.float32)
    assert np.allclose(bo.flip_boxes(b), [[0.1, 0.4, 0.5, 0.8], [0, 0, 0, 0]])


@pytest.mark.gpu
def test_device_preprocess_bit_exact():
    import torch
    from tf_ssd_b200.utils import data_utils
    rng = np.random.default_rng(8)
    batch = torch.zeros((3, 300, 300, 3), dtype=torch.float32, device="cuda")
    for i, (H, W, flip) in enumerate([(375, 500, False), (120, 90, True), (300, 300, False)]):
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        data_utils.device_preprocess(img, out=batch[i], flip=flip)
        ref = bo.preprocess_image(img, 300, 300, flip=flip)
        got = batch[i].cpu().numpy()
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (H, W, flip, np.abs(got - ref).max())
    out = data_utils.device_preprocess(rng.integers(0, 256, (64, 48, 3), dtype=np.uint8), final_height=512, final_width=512)
    assert out.shape == (512, 512, 3) and float(out.min()) >= 0 and float(out.max()) <= 1
    boxes = np.array([[[0.1, 0.2, 0.5, 0.6], [0.3, 0.0, 0.9, 1.0], [0, 0, 0, 0]]], np.float32)
    got = data_utils.device_flip_boxes(boxes.copy()).cpu().numpy()
    assert np.array_equal(got, bo.flip_boxes(boxes))



@pytest.mark.parametrize("backbone", ["vgg16", "mobilenet_v2"])
def test_h5_weight_converter_on_a_keras_shaped_tree(backbone, tmp_path):
    """tools/convert_h5_weights.py on a hand-built tree with Keras' on-disk shape (``model_weights/<layer>/<layer>/
    <variable>:0``; the reference's L2Normalization saves its unnamed variable as ``Variable:0``): every variable comes
    out under the name and shape ``SSDModel`` expects, and ``load_weights`` accepts the result (h5py itself is not
    available in this image: the file layer of the converter is the only part not exercised)."""
    import importlib.util
    from tf_ssd_b200.models.engine import SSDModel
    from tf_ssd_b200.utils import train_utils
    spec = importlib.util.spec_from_file_location("convert_h5_weights", os.path.join(ROOT, "tools", "convert_h5_weights.py"))
    conv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(conv)
    hp = train_utils.get_hyper_params(backbone)
    hp["total_labels"] = 21
    model = SSDModel(backbone, hp, seed=0)
    rng = np.random.default_rng(1)
    want = {k: rng.standard_normal(v.shape).astype(np.float32) for k, v in model.weights.items()}
    tree = {}
    for name, arr in want.items():
        layer, var = name.rsplit("/", 1)
        if layer == "l2_normalization":
            var = "Variable"                              # unnamed tf.Variable (models/ssd_vgg16.py:52)
        tree.setdefault(layer, {}).setdefault(layer, {})[var + ":0"] = arr.astype(np.float64)
    got = conv.convert_tree({"model_weights": tree, "optimizer_weights": {}} if backbone == "vgg16" else tree)
    assert sorted(got) == sorted(want)
    for k in want:
        assert got[k].dtype == np.float32 and np.array_equal(got[k], want[k]), k
    path = str(tmp_path / "w.npz")
    np.savez(path, **got)
    model.load_weights(path)
    assert all(np.array_equal(model.weights[k], want[k]) for k in want)
    with pytest.raises(ValueError):
        conv.convert_tree({"layer": {"layer": {"kernel:0": np.zeros(1), "mystery:0": np.zeros(1)}}})
    with pytest.raises(KeyError):
        model.set_weights({"no_such_layer/kernel": np.zeros(1, np.float32)})


def test_dataset_map_and_take():
    """``map`` must apply its function, ``take`` must keep concatenated parts."""
    from tf_ssd_b200.utils import data_utils
    ds, _ = data_utils.get_dataset("voc/2007", "test", total_items=6, img_size=64)
    other, _ = data_utils.get_dataset("voc/2012", "train+validation", total_items=5, img_size=64)
    seen = []
    mapped = ds.concatenate(other).map(lambda ex: (seen.append(1), data_utils.preprocessing(ex, 64, 64))[1])
    assert sum(1 for _ in mapped.take(9)) == 9 and len(seen) == 9                 # 6 of the first part + 3 of the second
    flipped = ds.map(lambda ex: (ex[0][:, ::-1], ex[1], ex[2]))
    a, b = next(iter(ds)), next(iter(flipped))
    assert np.array_equal(a[0][:, ::-1], b[0])
    with pytest.raises(ValueError):
        data_utils.preprocessing(a, 32, 32)


@pytest.mark.gpu
def test_preprocessing_of_a_tfds_style_example_and_batched_augmentation():
    """utils/data_utils.py:12-38 on a TFDS-shaped example (labels + 1, device resize, is_difficult filter, per-example
    augmentation_fn) and the batched form: ``train_utils.generator(..., augmentation_fn=augmentation.apply)``."""
    import torch
    from oracle import augment_oracle as ao
    from tf_ssd_b200 import augmentation
    from tf_ssd_b200.utils import bbox_utils, data_utils, train_utils
    rng = np.random.default_rng(0)
    example = {"image": rng.integers(0, 256, (90, 120, 3), dtype=np.uint8),
               "objects": {"bbox": np.array([[0.1, 0.2, 0.6, 0.7], [0.3, 0.3, 0.9, 0.8]], np.float32),
                           "label": np.array([4, 11], np.int64), "is_difficult": np.array([False, True])}}
    img, gb, gl = data_utils.preprocessing(example, 300, 300)
    assert np.array_equal(img.cpu().numpy(), bo.preprocess_image(example["image"], 300, 300))
    assert gl.tolist() == [5, 12] and gl.dtype == np.int32 and gb.shape == (2, 4)
    _, gb_e, gl_e = data_utils.preprocessing(example, 300, 300, evaluate=True)
    assert gl_e.tolist() == [5] and np.array_equal(gb_e, example["objects"]["bbox"][:1])
    draws = augmentation.ReplayDraws([0.25, 0.75, 0.25, 0.25, 0.25, 0.25])          # only the flip is taken
    a_img, a_gb, _ = data_utils.preprocessing(example, 300, 300, augmentation_fn=lambda i, b: augmentation.apply(i, b, draws=draws))
    assert np.array_equal(a_img.cpu().numpy(), bo.preprocess_image(example["image"], 300, 300, flip=True))
    assert np.array_equal(a_gb.cpu().numpy(), ao.flip_boxes(example["objects"]["bbox"]))
    # box helpers of utils/bbox_utils.py:217-269
    mm = [0.2, 0.1, 0.9, 0.6]
    assert np.allclose(bbox_utils.renormalize_bboxes_with_min_max(gb, mm).cpu().numpy(), ao.renormalize_boxes(gb, mm), atol=1e-7)
    px = bbox_utils.denormalize_bboxes(gb, 300, 500).cpu().numpy()
    assert np.array_equal(px, np.round(gb * np.array([300, 500, 300, 500], np.float32)))
    assert np.allclose(bbox_utils.normalize_bboxes(px, 300, 500).cpu().numpy(), gb, atol=2e-3)
    # batched: every batch of the feed is augmented on the device, targets are matched on the augmented boxes
    hp = train_utils.get_hyper_params("mobilenet_v2")
    hp["total_labels"] = 21
    priors = bbox_utils.generate_prior_boxes(hp["feature_map_shapes"], hp["aspect_ratios"])
    ds, _ = data_utils.get_dataset("voc/2007", "test", total_items=8)
    batches = ds.padded_batch(4, drop_remainder=True)
    seed_draws = augmentation.RandomDraws(7)
    feed = train_utils.generator(batches, priors, hp, augmentation_fn=lambda i, b: augmentation.apply(i, b, draws=seed_draws))
    img_b, (deltas, onehot) = next(feed)
    plain = next(train_utils.generator(batches, priors, hp))
    assert tuple(img_b.shape) == (4, 300, 300, 3) and tuple(deltas.shape) == (4, 2268, 4) and tuple(onehot.shape) == (4, 2268, 21)
    assert float(img_b.min()) >= 0.0 and float(img_b.max()) <= 1.0
    assert not torch.equal(img_b, torch.as_tensor(plain[0]).to(img_b.device))
    assert torch.isfinite(deltas).all() and float(onehot.sum(-1).min()) == 1.0
