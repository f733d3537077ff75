"""GPU parity tests for the SSD loss (ssd_loss.py) forward and backward."""

import numpy as np
import pytest
import torch

from oracle import box_oracle as bo
from tests.conftest import CONFIGS, VARIANCES

pytestmark = pytest.mark.gpu
RTOL = 1e-4          # north_star: loss tensors within 1e-4 relative fp32


def _np(t):
    return t.detach().cpu().numpy()


def _targets(B, name="mobilenet_v2", seed=0, G=16):
    from tf_ssd_b200 import synth
    priors = bo.prior_boxes(*CONFIGS[name][:2])
    gt, lab = synth.make_ground_truth(B, padded=G, seed=seed)
    ad, al = bo.match_encode(priors, gt, lab, 21, 0.5, VARIANCES)
    pd, z = synth.make_head_outputs(B, priors.shape[0], 21, seed=seed + 1)
    return ad, al, pd.astype(np.float32), z


@pytest.mark.parametrize("name,B", [("mobilenet_v2", 8), ("vgg16", 4), ("vgg16_512", 2)])
def test_loss_forward_parity(name, B):
    from tf_ssd_b200.ssd_loss import CustomLoss, ssd_loss
    ad, al, pd, z = _targets(B, name, seed=B)
    probs = bo.softmax(z)
    fn = CustomLoss(3, 1)
    np.testing.assert_allclose(_np(fn.loc_loss_fn(ad, pd)), bo.loc_loss(ad, pd, 1.0), rtol=RTOL)
    np.testing.assert_allclose(_np(fn.conf_loss_fn(al, probs)), bo.conf_loss(al, probs, 3.0), rtol=RTOL)
    np.testing.assert_allclose(_np(fn.conf_loss_fn(al, z, from_logits=True)),
                               bo.conf_loss(al, z, 3.0, from_logits=True), rtol=RTOL)
    both = _np(ssd_loss(ad, pd, al, probs))
    np.testing.assert_allclose(both, bo.loc_loss(ad, pd) + bo.conf_loss(al, probs), rtol=RTOL)
    fn2 = CustomLoss(2, 0.5)
    np.testing.assert_allclose(_np(fn2.loc_loss_fn(ad, pd)), bo.loc_loss(ad, pd, 0.5), rtol=RTOL)
    np.testing.assert_allclose(_np(fn2.conf_loss_fn(al, probs)), bo.conf_loss(al, probs, 2.0), rtol=RTOL)


def test_loss_edge_cases():
    from tf_ssd_b200.ssd_loss import CustomLoss
    fn = CustomLoss(3, 1)
    rng = np.random.default_rng(4)
    B, N, L = 4, 300, 21
    # image 0: no positives; image 1: every anchor positive; image 2: quirk (3*npos > #neg); image 3: ties
    lab = np.zeros((B, N), np.int64)
    lab[1] = rng.integers(1, L, N)
    lab[2, :100] = rng.integers(1, L, 100)
    lab[3, :5] = 3
    y = np.eye(L, dtype=np.float32)[lab]
    p = bo.softmax(rng.standard_normal((B, N, L)).astype(np.float32))
    p[3, 5:] = p[3, 5]                                  # identical rows -> identical CE -> rank ties
    ref, ce, final, rank = bo.conf_loss(y, p, 3.0, return_aux=True)
    assert final[2].max() == 2.0 and ref[0] == 0.0
    np.testing.assert_allclose(_np(fn.conf_loss_fn(y, p)), ref, rtol=RTOL)
    ad = (rng.standard_normal((B, N, 4)) * (lab > 0)[..., None]).astype(np.float32)
    pd = (3 * rng.standard_normal((B, N, 4))).astype(np.float32)
    got = _np(fn.loc_loss_fn(ad, pd))
    np.testing.assert_allclose(got, bo.loc_loss(ad, pd), rtol=RTOL)
    assert got[0] == 0.0
    # probability clip: p < 1e-7 on the true class -> CE = -log(1e-7)
    yc = np.zeros((1, 4, 3), np.float32); yc[0, :, 1] = 1
    pc = np.tile(np.array([1.0, 0.0, 0.0], np.float32), (1, 4, 1))
    np.testing.assert_allclose(_np(fn.conf_loss_fn(yc, pc)), bo.conf_loss(yc, pc), rtol=1e-6)
    # unnormalised probabilities are renormalised (Keras)
    pu = (p[:1] * 3.0).astype(np.float32)
    np.testing.assert_allclose(_np(fn.conf_loss_fn(y[:1], pu)), bo.conf_loss(y[:1], pu), rtol=RTOL)
    # N == 1 and tiny shapes
    np.testing.assert_allclose(_np(fn.conf_loss_fn(y[:, :1], p[:, :1])), bo.conf_loss(y[:, :1], p[:, :1]), rtol=RTOL)
    with pytest.raises(ValueError):
        fn.loc_loss_fn(ad, pd[:, :10])


def test_hard_negative_selection_exact():
    """Mask (which anchors are mined) must equal the oracle's rank rule exactly on tie-heavy input."""
    from tf_ssd_b200.ssd_loss import CustomLoss
    from tf_ssd_b200 import _ffi
    rng = np.random.default_rng(8)
    B, N, L = 6, 2268, 21
    lab = np.where(rng.random((B, N)) < 0.03, rng.integers(1, L, (B, N)), 0)
    y = np.eye(L, dtype=np.float32)[lab]
    # quantised probabilities: many exactly equal CE values
    base = np.round(rng.random((B, N, 1)) * 8) / 8
    p = np.concatenate([base * 0.5 + 0.25, np.full((B, N, L - 1), 1.0)], -1).astype(np.float32)
    p[..., 1:] = (1 - p[..., :1]) / (L - 1)
    ref, ce, final, rank = bo.conf_loss(y, p, 3.0, return_aux=True)
    fn = CustomLoss(3, 1)
    got = fn.conf_loss_fn(y, p)
    np.testing.assert_allclose(_np(got), ref, rtol=RTOL)
    # read the final mask back from the workspace layout: [ce | masked | hub | flags | fmask]
    bn = B * N
    up = lambda v: (v + 255) // 256 * 256
    off = up(bn * 4) * 3 + up(bn)
    fmask = _np(fn._ws[off:off + bn]).reshape(B, N)
    np.testing.assert_array_equal(fmask, final.astype(np.uint8))


def test_loss_backward_parity():
    from tf_ssd_b200.ssd_loss import CustomLoss
    ad, al, pd, z = _targets(4, "mobilenet_v2", seed=21)
    fn = CustomLoss(3, 1)
    loc, conf, gd, gz = fn.forward_backward(ad, pd, al, z)
    np.testing.assert_allclose(_np(loc), bo.loc_loss(ad, pd), rtol=RTOL)
    np.testing.assert_allclose(_np(conf), bo.conf_loss(al, z, from_logits=True), rtol=RTOL)
    rgd, rgz = bo.loss_grads(ad, pd, al, z)
    np.testing.assert_allclose(_np(gd), rgd, rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(_np(gz), rgz, rtol=1e-4, atol=1e-7)
    # torch autograd on the same masked objective as a second opinion
    _, _, final, _ = bo.conf_loss(al, z, 3.0, from_logits=True, return_aux=True)
    zt = torch.tensor(z, requires_grad=True)
    ce = torch.nn.functional.cross_entropy(zt.reshape(-1, 21), torch.tensor(al.argmax(-1)).reshape(-1),
                                           reduction="none").reshape(al.shape[:2])
    npos = torch.tensor((al[..., 1:] != 0).any(-1).sum(-1)).clamp(min=1)
    ((ce * torch.tensor(final)).sum(-1) / npos).mean().backward()
    np.testing.assert_allclose(_np(gz), zt.grad.numpy(), rtol=1e-4, atol=1e-7)
