"""CPU oracle for the tf-ssd hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the arithmetic of the reference's hot path
(FurkanOM/tf-ssd: ``utils/bbox_utils.py``, ``utils/train_utils.py``,
``ssd_loss.py``, ``models/decoder.py``, ``models/header.py``,
``models/ssd_mobilenet_v2.py``, ``models/ssd_vgg16.py``).  Every function cites
the reference file:line it follows.

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  Nothing under
``tf_ssd_b200/`` imports it; the product path fails loudly when its CUDA
library is missing instead of falling back to this code.

PARITY STATUS: *pinned by the reference's own source, executed on a NumPy
stand-in for TensorFlow*.  The reference is pure TensorFlow 2.0/Keras;
TensorFlow itself is not installable in this environment and the reference's own
tests execute no tensor op.  ``tests/golden/make_ref_golden.py`` imports the
UNMODIFIED modules under ``/root/reference`` (``utils/bbox_utils.py``,
``utils/train_utils.py``, ``ssd_loss.py``, ``models/decoder.py``,
``models/header.py``, ``models/ssd_vgg16.py``, ``models/ssd_mobilenet_v2.py``)
with ``tests/tf_shim`` (a NumPy-backed ``tensorflow`` module written from
TensorFlow's documented op semantics, which never imports this package) and
commits what they compute as ``tests/golden/ref_*.npz``.  This oracle and the CUDA
path are both held to those bytes (``tests/test_ref_golden.py``): priors, IoU (three
shape modes, 0/0 = NaN), encode/decode, ``calculate_actual_outputs`` (ties, padding),
both losses (mining quirk, both Huber reduction shapes), ``SSDDecoder.call``, and
the float32 forward of both model files (variable names and shapes included).

Also pinned: the reference's only hot-path known-answer test
(``tests/test_bbox_utils.py:19-22``, scale(3) == 0.48), its hyper-param
expectations (``tests/test_train_utils.py:27-60``), the derived KATs of SURVEY.md
section 8(c) (``tests/test_oracle_kats.py``), and ``torchvision.ops.nms`` as an
independent cross-check of greedy suppression.

What is STILL ``[TF-recall]`` (rules of TensorFlow that neither /root/reference nor
TensorFlow's documentation pins; the shim and this oracle encode the same recalled
rule, so the fixtures cannot disprove it): the visiting order of EQUAL scores
inside ``tf.image.combined_non_max_suppression`` and its strict ``>`` comparisons;
``tf.argmax`` returning the first maximum; the topology of
``keras_applications.mobilenet_v2`` 1.0.8 (third-party; not under /root/reference);
BatchNormalization epsilon / momentum and the Adam epsilon of Keras' defaults.
"""
