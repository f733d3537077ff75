"""Dataset plumbing with the reference's names (``utils/data_utils.py`` of FurkanOM/tf-ssd).

The reference reads PASCAL VOC through ``tensorflow_datasets``; neither the data nor TFDS exist in this
environment, so ``get_dataset`` serves a SEEDED SYNTHETIC dataset with the same tensor contract
(``utils/data_utils.py:33-37,140-155``): images NHWC float32 in [0,1], boxes normalised ``[y1,x1,y2,x2]``,
labels shifted by +1 (0 = background), batches padded with box 0 / label -1.  Host-side NumPy only."""

from __future__ import annotations

import zlib
from typing import Any, Dict, Iterator, List, Tuple

import numpy as np

VOC_LABELS = ["aeroplane", "bicycle", "bird", "boat", "bottle", "bus", "car", "cat", "chair", "cow", "diningtable",
              "dog", "horse", "motorbike", "person", "pottedplant", "sheep", "sofa", "train", "tvmonitor"]
# split sizes of the TFDS voc builders the reference uses (trainer.py:47-58)
_SPLIT_ITEMS = {("voc/2007", "train+validation"): 5011, ("voc/2007", "test"): 4952, ("voc/2012", "train+validation"): 11540}


class SyntheticVOC(object):
    """Iterable of ``(img [S,S,3], gt_boxes [g,4], gt_labels [g])`` examples; ``padded_batch`` mirrors tf.data."""

    def __init__(self, total_items: int, img_size: int = 300, seed: int = 0, max_boxes: int = 8):
        self.total_items, self.img_size, self.seed, self.max_boxes = int(total_items), int(img_size), int(seed), int(max_boxes)
        self._parts: List["SyntheticVOC"] = [self]
        self._fns: List[Any] = []
        self._limit = -1

    def _view(self, total_items: int) -> "SyntheticVOC":
        out = SyntheticVOC(total_items, self.img_size, self.seed, self.max_boxes)
        out._parts, out._fns, out._limit = list(self._parts), list(self._fns), self._limit
        return out

    def concatenate(self, other: "SyntheticVOC") -> "SyntheticVOC":
        out = SyntheticVOC(self.total_items + other.total_items, self.img_size, self.seed, self.max_boxes)
        out._parts = self._parts + other._parts
        return out

    def map(self, fn) -> "SyntheticVOC":
        """``tf.data.Dataset.map``: ``fn`` is applied to every example when the dataset is iterated."""
        out = self._view(self.total_items)
        out._fns.append(fn)
        return out

    def shuffle(self, buffer_size: int) -> "SyntheticVOC":
        """The synthetic stream is i.i.d. already: shuffling it changes nothing a training run can observe."""
        return self

    def take(self, n: int) -> "SyntheticVOC":
        """``tf.data.Dataset.take``: the first ``n`` examples (across concatenated parts)."""
        out = self._view(min(int(n), self.total_items))
        out._limit = out.total_items
        return out

    def _example(self, rng: np.random.Generator) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        S = self.img_size
        g = int(rng.integers(1, self.max_boxes + 1))
        c, wh = rng.random((g, 2)), rng.uniform(0.1, 0.6, (g, 2))
        boxes = np.clip(np.concatenate([c - wh / 2, c + wh / 2], -1), 0.0, 1.0).astype(np.float32)
        labels = rng.integers(1, len(VOC_LABELS) + 1, g).astype(np.int32)      # +1: background is class 0 (:36)
        img = np.full((S, S, 3), 0.4, np.float32) + 0.1 * rng.random((S, S, 3), dtype=np.float32)
        for (y1, x1, y2, x2), lab in zip(boxes, labels):                          # class-coloured rectangles: learnable
            col = np.array([(lab * 37 % 255) / 255.0, (lab * 91 % 255) / 255.0, (lab * 53 % 255) / 255.0], np.float32)
            img[int(y1 * S):max(int(y2 * S), int(y1 * S) + 1), int(x1 * S):max(int(x2 * S), int(x1 * S) + 1)] = col
        return img, boxes, labels

    def __iter__(self) -> Iterator[Tuple[np.ndarray, np.ndarray, np.ndarray]]:
        seen = 0
        for k, part in enumerate(self._parts):
            rng = np.random.default_rng(part.seed + 7919 * k)
            for _ in range(part.total_items):
                if 0 <= self._limit <= seen:
                    return
                example = part._example(rng)
                for fn in self._fns:
                    example = fn(example)
                seen += 1
                yield example

    def padded_batch(self, batch_size: int, padded_shapes: Any = None, padding_values: Any = None, drop_remainder: bool = False):
        """``tf.data.Dataset.padded_batch`` (trainer.py:75-84): ragged ground truth padded to the batch maximum."""
        ds = self

        class _Batched(object):
            def __iter__(self_inner):
                imgs, boxes, labels = [], [], []
                for img, b, l in ds:
                    imgs.append(_host(img)); boxes.append(_host(b)); labels.append(_host(l))
                    if len(imgs) == batch_size:
                        yield _pad(imgs, boxes, labels)
                        imgs, boxes, labels = [], [], []
                if imgs and not drop_remainder:
                    yield _pad(imgs, boxes, labels)
        return _Batched()


def _host(x: Any) -> np.ndarray:
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def _pad(imgs, boxes, labels):
    g = max(b.shape[0] for b in boxes)
    gb = np.zeros((len(imgs), g, 4), np.float32)                      # padding value 0 (:151)
    gl = np.full((len(imgs), g), -1, np.int32)                        # padding value -1 (:153)
    for i, (b, l) in enumerate(zip(boxes, labels)):
        gb[i, :b.shape[0]] = b
        gl[i, :l.shape[0]] = l
    return np.stack(imgs), gb, gl


def get_dataset(name: str, split: str, data_dir: str = "~/tensorflow_datasets", total_items: int = 0, img_size: int = 300):
    """utils/data_utils.py:40-56 -> ``(dataset, info)``; synthetic stand-in (see the module docstring)."""
    n = int(total_items) or _SPLIT_ITEMS.get((name, split), 1000)
    seed = zlib.crc32(f"{name}:{split}".encode()) % (1 << 16)
    return SyntheticVOC(n, img_size=img_size, seed=seed), {"name": name, "splits": {split: n}, "labels": list(VOC_LABELS)}


def get_total_item_size(info: Any, split: str) -> int:
    """utils/data_utils.py:59-72: ``"train+validation"`` sums the named splits.  Accepts the synthetic ``info`` dict of
    ``get_dataset`` as well as a TFDS ``DatasetInfo`` (``info.splits[name].num_examples``)."""
    splits = info["splits"] if isinstance(info, dict) else info.splits
    if split in splits and not hasattr(splits[split], "num_examples"):
        return int(splits[split])
    total = 0
    for name in split.split("+"):
        item = splits[name]
        total += int(getattr(item, "num_examples", item))
    return total


def get_labels(info: Any) -> List[str]:
    """utils/data_utils.py:75-85 (TFDS: ``info.features["labels"].names``)."""
    if isinstance(info, dict):
        return list(info["labels"])
    return list(info.features["labels"].names)


def get_custom_imgs(custom_image_path: str) -> List[str]:
    """utils/data_utils.py:88-101: the files directly inside ``custom_image_path`` (sub-directories are ignored)."""
    import os
    for path, _, files in os.walk(custom_image_path):
        return [os.path.join(path, name) for name in files]
    return []


def preprocessing(image_data: Any, final_height: int, final_width: int, augmentation_fn: Any = None, evaluate: bool = False):
    """utils/data_utils.py:12-38 for one example -> ``(img, gt_boxes, gt_labels)``.

    A TFDS-style example (``{"image": uint8 [H,W,3], "objects": {"bbox", "label", "is_difficult"}}``) goes through the
    reference's steps: labels + 1 (:35), ``convert_image_dtype`` + ``resize`` on the device (:36-37,
    ``ssd_preprocess_image``), the ``is_difficult`` filter when ``evaluate`` (:38-41), then ``augmentation_fn`` (:42-43).
    The synthetic examples of ``get_dataset`` are ``(img, gt_boxes, gt_labels)`` tuples that are already resized and
    normalised: only ``augmentation_fn`` applies to them.  (The fast path for training augments whole batches on the
    device instead: ``train_utils.generator(..., augmentation_fn=augmentation.apply)``.)"""
    if isinstance(image_data, dict):
        objects = image_data["objects"]
        gt_boxes = np.asarray(_host(objects["bbox"]), np.float32).reshape(-1, 4)
        gt_labels = (np.asarray(_host(objects["label"])).astype(np.int64) + 1).astype(np.int32)
        img = device_preprocess(image_data["image"], final_height=final_height, final_width=final_width)
        if evaluate:
            not_diff = np.logical_not(np.asarray(_host(objects["is_difficult"]), bool))
            gt_boxes, gt_labels = gt_boxes[not_diff], gt_labels[not_diff]
    else:
        img, gt_boxes, gt_labels = image_data
        if tuple(np.shape(img)[:2]) != (final_height, final_width):
            raise ValueError(f"synthetic example is {np.shape(img)[:2]}, expected {(final_height, final_width)}")
    if augmentation_fn is not None:
        img, gt_boxes = augmentation_fn(img, gt_boxes)
    return img, gt_boxes, gt_labels


def get_data_types():
    """utils/data_utils.py:127-137 (NumPy dtypes instead of TensorFlow's)."""
    return (np.dtype("float32"), np.dtype("float32"), np.dtype("int32"))


def get_data_shapes():
    """utils/data_utils.py:140-147."""
    return ([None, None, None], [None, None], [None])


def get_padding_values():
    """utils/data_utils.py:150-155."""
    return (np.float32(0), np.float32(0), np.int32(-1))


def device_preprocess(img_u8: Any, out: Any = None, final_height: int = 300, final_width: int = 300, flip: bool = False):
    """The per-example work of ``preprocessing`` (utils/data_utils.py:33-37) on the DEVICE: uint8 ``[H,W,3]`` ->
    float32 ``[final_height, final_width, 3]`` in [0,1] (convert_image_dtype + bilinear resize, half-pixel centres),
    optionally mirrored (augmentation.py:flip_horizontally), written into ``out`` -- typically one slot
    ``batch[i]`` of the NHWC batch buffer the network reads, so no host-side resize / float image ever exists."""
    import torch
    from tf_ssd_b200 import _ffi
    src = _ffi.to_dev(img_u8, dtype=torch.uint8)
    if src.dim() != 3 or src.shape[2] != 3:
        raise ValueError("expected a uint8 image [H, W, 3]")
    if out is None:
        out = torch.empty((final_height, final_width, 3), dtype=torch.float32, device=src.device)
    if tuple(out.shape) != (final_height, final_width, 3) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("out must be a contiguous float32 [final_height, final_width, 3] CUDA tensor")
    _ffi.check(_ffi.lib().ssd_preprocess_image(_ffi.ptr(src), int(src.shape[0]), int(src.shape[1]), _ffi.ptr(out), final_height,
                                               final_width, int(bool(flip)), _ffi.stream()), "ssd_preprocess_image")
    return out


def device_flip_boxes(gt_boxes: Any):
    """augmentation.py:128-137 on the device, in place on a float32 ``[..., 4]`` CUDA tensor (padding rows stay zero)."""
    from tf_ssd_b200 import _ffi
    b = _ffi.to_dev(gt_boxes)
    _ffi.check(_ffi.lib().ssd_flip_boxes(_ffi.ptr(b), int(b.numel() // 4), _ffi.stream()), "ssd_flip_boxes")
    return b

