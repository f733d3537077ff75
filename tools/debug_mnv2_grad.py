"""Per-variable gradient error of the MobileNetV2 training step against the torch-autograd oracle (forward order)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import train_oracle as to
from tests.test_train_gpu import _mnv2_setup, _l2
from tf_ssd_b200.models.train_engine import Trainer

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 256.0
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ratio = float(sys.argv[3]) if len(sys.argv) > 3 else 3.0
model, hp, img, ad, al = _mnv2_setup(B)
stats = {}
ref_loss, ref_grads = to.train_step(model.weights, hp, img, ad, al, model.l2_kernels, neg_pos_ratio=ratio, backbone="mobilenet_v2", fp16sim=len(sys.argv) > 4, stats=stats)
tr = Trainer(model, loss_scale=scale, neg_pos_ratio=ratio)
out = tr.forward_backward(img, ad, al)
torch.cuda.synchronize()
print("loss", out["loc"].cpu().numpy(), ref_loss["loc"], out["conf"].cpu().numpy(), ref_loss["conf"])
st = tr._state[B]
for name, v in tr.vars.items():
    g = v["grad"].cpu().numpy() / tr.loss_scale
    layer, var = name.rsplit("/", 1)
    if layer.endswith("_conv_head"):
        continue
    if var == "kernel":
        ref = ref_grads[name].transpose(3, 0, 1, 2)
        g = g[..., :ref.shape[3]]
    elif var == "depthwise_kernel":
        ref = ref_grads[name][..., 0]
    else:
        ref = ref_grads[name]
    print(f"{name:45s} err {_l2(g, ref):8.4f}  |ref| {np.abs(ref).max():9.3e}  |got| {np.abs(g).max():9.3e}")
plan = st["plan"]
for s in plan.steps:
    if s.kind == "bn":
        mean, var, _ = stats[s.name]
        Cc = mean.shape[0]
        save = s.meta["save"].cpu().numpy()
        print(f"FWD {s.name:35s} mean err {_l2(save[:Cc], mean):9.3e} rstd err {_l2(save[Cc:], 1.0 / np.sqrt(var + 1e-3)):9.3e}  min var {var.min():9.3e} max rstd {save[Cc:].max():8.2f}")
for s in plan.steps:
    if s.kind == "bn":
        gbuf = st["grad_of"].get(s.meta["out"].data_ptr())
        if gbuf is not None:
            a = gbuf.float().abs()
            print(f"dY {s.name:35s} max {float(a.max()):9.3e} nonzero-frac {float((a > 0).float().mean()):6.3f} inf {bool(torch.isinf(gbuf).any())}")
