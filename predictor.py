#!/usr/bin/env python
"""Prediction / evaluation entry point: the flow of the reference's ``predictor.py:29-116`` on the B200-native path.

hyper-parameters -> test dataset -> ``get_model`` / ``init_model`` / ``load_weights`` -> prior boxes ->
``get_decoder_model`` -> ``predict`` -> ``eval_utils.evaluate_predictions`` (mAP).  The dataset is the synthetic VOC
stand-in (no TFDS here); drawing is replaced by a printed summary.

    python predictor.py --backbone mobilenet_v2 --val-items 128
"""

from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tf_ssd_b200.models.decoder import get_decoder_model                               # noqa: E402
from tf_ssd_b200.utils import bbox_utils, data_utils, eval_utils, io_utils, train_utils   # noqa: E402


def _get_model_fns(backbone):
    if backbone == "mobilenet_v2":
        from tf_ssd_b200.models.ssd_mobilenet_v2 import get_model, init_model
    else:
        from tf_ssd_b200.models.ssd_vgg16 import get_model, init_model
    return get_model, init_model


def main(argv=None, evaluate=True):
    args = io_utils.handle_args(argv)
    if args.handle_gpu:
        io_utils.handle_gpu_compatibility()
    batch_size = args.batch_size
    backbone = args.backbone
    io_utils.is_valid_backbone(backbone)
    get_model, init_model = _get_model_fns(backbone)
    hyper_params = train_utils.get_hyper_params(backbone)
    img_size = hyper_params["img_size"]

    test_data, info = data_utils.get_dataset("voc/2007", "test", total_items=args.val_items, img_size=img_size)
    total_items = data_utils.get_total_item_size(info, "test")
    labels = ["bg"] + data_utils.get_labels(info)
    hyper_params["total_labels"] = len(labels)
    test_data = test_data.map(lambda x: data_utils.preprocessing(x, img_size, img_size, evaluate=evaluate))
    test_data = test_data.padded_batch(batch_size, padded_shapes=data_utils.get_data_shapes(),
                                       padding_values=data_utils.get_padding_values(), drop_remainder=True)

    ssd_model = get_model(hyper_params)
    init_model(ssd_model)
    ssd_model_path = io_utils.get_model_path(backbone, args.model_dir)
    if os.path.exists(ssd_model_path):
        ssd_model.load_weights(ssd_model_path)
    else:
        print(f"{ssd_model_path} not found: predicting with the seeded random initialisation")
    prior_boxes = bbox_utils.generate_prior_boxes(hyper_params["feature_map_shapes"], hyper_params["aspect_ratios"])
    ssd_decoder_model = get_decoder_model(ssd_model, prior_boxes, hyper_params)

    step_size = max(1, total_items // batch_size)
    pred_bboxes, pred_labels, pred_scores = ssd_decoder_model.predict(test_data, steps=step_size, verbose=1)
    print({"images": int(pred_bboxes.shape[0]), "detections_per_image": float((pred_scores > 0).sum(-1).mean())})
    if evaluate:
        return eval_utils.evaluate_predictions(test_data, pred_bboxes, pred_labels, pred_scores, labels, batch_size)
    return pred_bboxes, pred_labels, pred_scores


if __name__ == "__main__":
    main()
