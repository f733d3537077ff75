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

PARITY STATUS: *parity unpinned against TensorFlow*.  The reference is pure
TensorFlow 2.0/Keras; TensorFlow is not installable in this environment and
the reference's own tests execute no tensor op.  What IS pinned:

* the reference's only hot-path known-answer test
  (``tests/test_bbox_utils.py:19-22``, scale(3) == 0.48) and its hyper-param
  expectations (``tests/test_train_utils.py:27-60``);
* the derived KATs of SURVEY.md section 8(c) (prior sums / rows / counts, the
  rank example), re-derived independently in ``tests/test_oracle_kats.py``;
* ``torchvision.ops.nms`` as an independent cross-check of greedy suppression.

Third-party arithmetic that is restated from the published TensorFlow
algorithm (tensorflow==2.0.0, keras-applications==1.0.8, ``environment.yml``)
is marked ``[TF-recall]`` where it is used.
"""
