#!/usr/bin/env python
"""Training entry point: the flow of the reference's ``trainer.py:31-127`` on the B200-native hot path.

Same steps, same names: hyper-parameters -> datasets -> ``get_model`` -> ``CustomLoss`` -> ``model.compile`` ->
``init_model`` -> prior boxes -> ``train_utils.generator`` feeds -> ``ModelCheckpoint`` (best ``val_loss``) +
``LearningRateScheduler(train_utils.scheduler)`` -> ``model.fit``.  Differences: the dataset is the seeded synthetic
VOC stand-in of ``utils/data_utils.py`` (no TFDS / no network here), weights are ``.npz``, there is no TensorBoard.
Under ``torchrun`` every rank trains on its own shard and gradients are averaged with one NCCL all-reduce per step.

    python trainer.py --backbone mobilenet_v2 --epochs 2 --train-items 256 --val-items 64
"""

from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tf_ssd_b200 import augmentation, dist_utils                                                    # noqa: E402
from tf_ssd_b200.models.train_engine import Adam, LearningRateScheduler, ModelCheckpoint   # noqa: E402
from tf_ssd_b200.ssd_loss import CustomLoss                                            # noqa: E402
from tf_ssd_b200.utils import bbox_utils, data_utils, io_utils, train_utils            # noqa: E402


def _get_model_fns(backbone):
    if backbone == "mobilenet_v2":
        from tf_ssd_b200.models.ssd_mobilenet_v2 import get_model, init_model
    else:
        from tf_ssd_b200.models.ssd_vgg16 import get_model, init_model
    return get_model, init_model


def main(argv=None):
    import torch
    args = io_utils.handle_args(argv)
    if args.handle_gpu:
        io_utils.handle_gpu_compatibility()
    rank, local_rank, world = dist_utils.env_rank()
    torch.cuda.set_device(local_rank)
    dist_utils.init_from_env("nccl")

    batch_size, epochs, load_weights, with_voc_2012 = args.batch_size, args.epochs, False, True
    backbone = args.backbone
    io_utils.is_valid_backbone(backbone)
    get_model, init_model = _get_model_fns(backbone)
    hyper_params = train_utils.get_hyper_params(backbone)
    img_size = hyper_params["img_size"]

    train_data, info = data_utils.get_dataset("voc/2007", "train+validation", total_items=args.train_items, img_size=img_size)
    val_data, val_info = data_utils.get_dataset("voc/2007", "test", total_items=args.val_items, img_size=img_size)
    train_data.seed += 1000 * rank                                  # every rank draws its own shard of the synthetic stream
    train_total_items = data_utils.get_total_item_size(info, "train+validation")
    val_total_items = data_utils.get_total_item_size(val_info, "test")
    if with_voc_2012 and not args.train_items:
        voc_2012_data, voc_2012_info = data_utils.get_dataset("voc/2012", "train+validation", img_size=img_size)
        train_total_items += data_utils.get_total_item_size(voc_2012_info, "train+validation")
        train_data = train_data.concatenate(voc_2012_data)
    labels = ["bg"] + data_utils.get_labels(info)
    hyper_params["total_labels"] = len(labels)

    train_data = train_data.map(lambda x: data_utils.preprocessing(x, img_size, img_size))
    val_data = val_data.map(lambda x: data_utils.preprocessing(x, img_size, img_size))
    data_shapes, padding_values = data_utils.get_data_shapes(), data_utils.get_padding_values()
    train_data = train_data.shuffle(batch_size * 4).padded_batch(batch_size, padded_shapes=data_shapes,
                                                                 padding_values=padding_values, drop_remainder=True)
    val_data = val_data.padded_batch(batch_size, padded_shapes=data_shapes, padding_values=padding_values, drop_remainder=True)

    ssd_model = get_model(hyper_params)
    ssd_custom_losses = CustomLoss(hyper_params["neg_pos_ratio"], hyper_params["loc_loss_alpha"])
    ssd_model.compile(optimizer=Adam(learning_rate=1e-3), loss=[ssd_custom_losses.loc_loss_fn, ssd_custom_losses.conf_loss_fn])
    init_model(ssd_model)

    ssd_model_path = io_utils.get_model_path(backbone, args.model_dir)
    if load_weights:
        ssd_model.load_weights(ssd_model_path)
    prior_boxes = bbox_utils.generate_prior_boxes(hyper_params["feature_map_shapes"], hyper_params["aspect_ratios"])
    # trainer.py:68 maps augmentation.apply over the examples; here every padded batch is augmented on the device
    augmentation_fn = None if args.no_augmentation else augmentation.apply
    ssd_train_feed = train_utils.generator(train_data, prior_boxes, hyper_params, augmentation_fn=augmentation_fn)
    ssd_val_feed = train_utils.generator(val_data, prior_boxes, hyper_params)

    callbacks = [LearningRateScheduler(train_utils.scheduler)]
    checkpoint_callback = ModelCheckpoint(ssd_model_path, monitor="val_loss", save_best_only=True, save_weights_only=True)
    if rank == 0:
        callbacks.insert(0, checkpoint_callback)
    step_size_train = max(1, train_total_items // batch_size)
    step_size_val = max(1, val_total_items // batch_size)
    history = ssd_model.fit(ssd_train_feed, steps_per_epoch=step_size_train, validation_data=ssd_val_feed,
                            validation_steps=step_size_val, epochs=epochs, callbacks=callbacks, verbose=1 if rank == 0 else 0)
    if rank == 0:
        print({"loss": history["loss"], "val_loss": history["val_loss"], "saved_epochs": checkpoint_callback.saved_epochs,
               "weights": ssd_model_path})
    return history


if __name__ == "__main__":
    main()
