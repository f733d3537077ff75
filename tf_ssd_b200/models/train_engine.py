"""Training step behind ``model.compile(...)`` / ``model.fit(...)``
(trainer.py:86-127 of the reference: Keras compile + fit with ``Adam(1e-3)`` and the two
``CustomLoss`` callables).

One step = forward plan (the inference launch list, activations kept) -> fused SSD loss forward
and backward (``ssd_loss_fwd/bwd``, from logits) -> backward launch list derived by walking the
forward plan in reverse -> optional gradient all-reduce (``dist_utils.GradBuckets``) -> fused
Adam on fp32 master weights that also refreshes the fp16 working copies the plans point at.

Mixed precision: activations and activation gradients are fp16 (loss-scaled), variable gradients,
master weights and Adam moments are fp32.

Scope: both graphs.  SSD300-VGG16 (BASELINE.json config 3, and the ``vgg16_512`` extension) trains on its
inference plan (no BatchNorm).  SSD300-MobileNetV2 (config 4) trains on ``SSDModel.train_plan``: BatchNorm in
training mode (``ssd_bn_train_fwd/bwd``: batch statistics, moving-average update), depthwise data / filter
gradients (``ssd_depthwise3x3_dgrad/wgrad``), shortcut adds folded into the BatchNorm kernels.
"""

from __future__ import annotations

import ctypes as C
import math
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from tf_ssd_b200 import _ffi, dist_utils
from tf_ssd_b200._ffi_conv import ACT_NONE, ConvDesc
from tf_ssd_b200.models.engine import SSDModel, Step, _no_gc_during_capture

L2_REG = 5e-4                  # models/ssd_vgg16.py:76  reg_factor


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


class Adam(object):
    """Stand-in for ``tensorflow.keras.optimizers.Adam`` (trainer.py:8,92): a bag of hyper-parameters
    consumed by ``model.compile``; the update itself is the fused ``ssd_adam_step`` kernel."""

    def __init__(self, learning_rate: float = 1e-3, beta_1: float = 0.9, beta_2: float = 0.999, epsilon: float = 1e-7):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon


class LearningRateScheduler(object):
    """``tensorflow.keras.callbacks.LearningRateScheduler(schedule)`` (trainer.py:7,116)."""

    def __init__(self, schedule):
        self.schedule = schedule

    def on_epoch_begin(self, epoch, logs=None):
        return float(self.schedule(epoch))


class ModelCheckpoint(object):
    """``tensorflow.keras.callbacks.ModelCheckpoint(path, monitor="val_loss", save_best_only=True,
    save_weights_only=True)`` (trainer.py:108-113): writes ``model.save_weights(path)`` when the monitored value
    improves (or every epoch without ``save_best_only``)."""

    def __init__(self, filepath: str, monitor: str = "val_loss", save_best_only: bool = False, save_weights_only: bool = True):
        self.filepath, self.monitor, self.save_best_only = filepath, monitor, save_best_only
        self.best = math.inf
        self.model = None
        self.saved_epochs: List[int] = []

    def set_model(self, model) -> None:
        self.model = model

    def on_epoch_end(self, epoch, logs=None):
        value = (logs or {}).get(self.monitor)
        if self.save_best_only and (value is None or not value < self.best):
            return
        if value is not None:
            self.best = min(self.best, value)
        trainer = getattr(self.model, "trainer", None)
        if trainer is not None:
            trainer.sync_weights_to_host()
        self.model.save_weights(self.filepath)
        self.saved_epochs.append(epoch)


class Trainer(object):
    """Owns the master weights, Adam state, gradient buffers and the backward launch list."""

    def __init__(self, model: SSDModel, learning_rate: float = 1e-3, neg_pos_ratio: float = 3.0, loc_loss_alpha: float = 1.0,
                 beta_1: float = 0.9, beta_2: float = 0.999, epsilon: float = 1e-7, loss_scale: float = 1024.0,
                 use_cuda_graph: bool = True, dynamic_loss_scale: bool = True, scale_check_every: int = 100,
                 scale_growth_interval: int = 2000):
        _ffi.check_device()
        self.init_kwargs = dict(learning_rate=learning_rate, neg_pos_ratio=neg_pos_ratio, loc_loss_alpha=loc_loss_alpha,
                                beta_1=beta_1, beta_2=beta_2, epsilon=epsilon, loss_scale=loss_scale,
                                use_cuda_graph=use_cuda_graph, dynamic_loss_scale=dynamic_loss_scale,
                                scale_check_every=scale_check_every, scale_growth_interval=scale_growth_interval)
        # Mixed-precision guard: a step whose gradients hold a non-finite value is SKIPPED on the device (guarded Adam);
        # the host looks at the skip counter every ``scale_check_every`` steps (one synchronisation) and halves the loss
        # scale after an overflow / doubles it after ``scale_growth_interval`` clean steps (re-capturing the step graph).
        self.dynamic_loss_scale = bool(dynamic_loss_scale)
        self.scale_check_every, self.scale_growth_interval = int(scale_check_every), int(scale_growth_interval)
        self._since_check = self._clean_steps = 0
        self.skipped_steps = 0
        self._pending: List[Any] = []                # in-flight gradient all-reduces of the current step
        self.model = model
        model.trainer = self                         # the model's training-mode variables now belong to this trainer
        self.use_cuda_graph = bool(use_cuda_graph)
        self._eval_dirty = True
        self.lib = _ffi.lib()
        self.dev = _ffi.require_cuda()
        self.lr, self.b1, self.b2, self.eps = float(learning_rate), float(beta_1), float(beta_2), float(epsilon)
        self.neg_pos_ratio, self.alpha = float(neg_pos_ratio), float(loc_loss_alpha)
        self.loss_scale = float(loss_scale)
        self.t = 0
        self._bwd: Dict[int, List[Tuple[Any, tuple, tuple]]] = {}
        self._state: Dict[int, Dict[str, Any]] = {}
        self._build_variables()

    # -- variables: fp32 masters in the kernels' layouts, views into flat gradient buckets ---------------------
    def _build_variables(self) -> None:
        m = self.model
        plan = m.train_plan(1)                        # forces every packed (fp16, OHWI) tensor into existence
        self.vars: Dict[str, Dict[str, Any]] = {}
        shapes = []
        for s in plan.steps:
            if s.kind == "conv":
                w16, b32 = s.meta["w"], s.meta["bias"]
                self.vars[s.name + "/kernel"] = dict(w16=w16, master=w16.float().clone(), l2=0.0)
                shapes.append((s.name + "/kernel", tuple(w16.shape)))
                if b32 is not None:                   # layers followed by BatchNorm have no bias
                    self.vars[s.name + "/bias"] = dict(w16=None, master=b32, l2=0.0)    # bias buffers are fp32 already
                    shapes.append((s.name + "/bias", tuple(b32.shape)))
            elif s.kind == "dw":
                w16 = s.meta["w"]
                self.vars[s.name + "/depthwise_kernel"] = dict(w16=w16, master=w16.float().clone(), l2=0.0)
                shapes.append((s.name + "/depthwise_kernel", tuple(w16.shape)))
            elif s.kind == "bn":
                for var in ("gamma", "beta"):
                    self.vars[f"{s.name}/{var}"] = dict(w16=None, master=s.meta[var], l2=0.0)
                    shapes.append((f"{s.name}/{var}", tuple(s.meta[var].shape)))
            elif s.kind == "l2norm":
                sc = s.meta["scale"]
                self.vars[s.name + "/scale"] = dict(w16=None, master=sc, l2=0.0)
                shapes.append((s.name + "/scale", tuple(sc.shape)))
        for k in m.l2_kernels:                        # Keras name "<layer>/kernel" == step name + "/kernel"
            if k in self.vars:
                self.vars[k]["l2"] = 2.0 * L2_REG
        self.grads = dist_utils.GradBuckets(shapes, self.dev)
        for name, v in self.vars.items():
            v["grad"] = self.grads.views[name]
            v["m"] = torch.zeros_like(v["master"])
            v["v"] = torch.zeros_like(v["master"])
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=self.dev)
        self._guard = torch.zeros(2, dtype=torch.int32, device=self.dev)     # [overflow of this step, skipped steps]
        # weights of layers WITHOUT BatchNorm live in the model's packed cache and are updated in place by Adam: they
        # must survive the cache being cleared (sync_weights_to_host), or a plan built later for another batch size would
        # re-upload stale host copies that Adam never touches
        self._owned_packed = {k: v for k, v in m._packed.items()
                              if any(v[0] is s.meta.get("w") or v[0] is s.meta.get("scale") for s in plan.steps if s.meta)}
        # descriptor table of the multi-tensor Adam launch (device array of struct ssd_adam_var)
        from tf_ssd_b200._ffi_conv import AdamVar
        table = (AdamVar * len(self.vars))()
        for i, v in enumerate(self.vars.values()):
            table[i].w, table[i].m, table[i].v = v["master"].data_ptr(), v["m"].data_ptr(), v["v"].data_ptr()
            table[i].grad = v["grad"].data_ptr()
            table[i].w16 = v["w16"].data_ptr() if v["w16"] is not None else None
            table[i].n, table[i].l2 = v["master"].numel(), v["l2"]
        raw = np.frombuffer(bytes(table), dtype=np.uint8).copy()
        self._adam_table = torch.from_numpy(raw).to(self.dev)
        self._adam_max_n = max(v["master"].numel() for v in self.vars.values())

    def sync_weights_to_host(self) -> None:
        """Write the trained variables back into ``model.weights`` (Keras names / layouts)."""
        m = self.model
        for name, v in self.vars.items():
            arr = v["master"].detach().cpu().numpy()
            layer, var = name.rsplit("/", 1)
            if layer.endswith("_conv_head"):
                idx = layer.split("_")[0]
                nl = m.weights[f"{idx}_conv_label_output/bias"].shape[0]
                if var == "kernel":
                    hwio = arr.transpose(1, 2, 3, 0)
                    m.weights[f"{idx}_conv_label_output/kernel"] = np.ascontiguousarray(hwio[..., :nl])
                    m.weights[f"{idx}_conv_boxes_output/kernel"] = np.ascontiguousarray(hwio[..., nl:])
                else:
                    m.weights[f"{idx}_conv_label_output/bias"] = arr[:nl].copy()
                    m.weights[f"{idx}_conv_boxes_output/bias"] = arr[nl:].copy()
            elif var == "kernel":
                cin = m.weights[name].shape[2]
                m.weights[name] = np.ascontiguousarray(arr.transpose(1, 2, 3, 0)[:, :, :cin, :])
            elif var == "depthwise_kernel":
                m.weights[name] = np.ascontiguousarray(arr[..., None])                   # [3,3,C] -> [3,3,C,1]
            else:
                m.weights[name] = arr.copy()
        if m.has_batchnorm:
            for key, t in m._train_vars.items():     # moving statistics are updated by the forward kernels
                if key.endswith("/moving_mean") or key.endswith("/moving_variance"):
                    m.weights[key] = t.detach().cpu().numpy().copy()
            m._packed.clear()                        # inference plans fold BatchNorm: rebuild them from the new values
            m._packed.update(self._owned_packed)     # ... except the trainer-owned (BatchNorm-free) layers: same tensors
            m._plans.clear()
            m._version = getattr(m, "_version", 0) + 1
            self._eval_dirty = False

    # -- backward launch list ---------------------------------------------------------------------------------
    def _prepare(self, B: int) -> Dict[str, Any]:
        if B in self._state:
            return self._state[B]
        m, lib, dev = self.model, self.lib, self.dev
        plan = m.train_plan(B)
        N, L = m.n_anchors, m.total_labels
        st: Dict[str, Any] = dict(plan=plan)
        st["g_deltas"] = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
        st["g_logits"] = torch.empty((B, N, L), dtype=torch.float32, device=dev)
        st["loc"] = torch.empty((B,), dtype=torch.float32, device=dev)
        st["conf"] = torch.empty((B,), dtype=torch.float32, device=dev)
        st["ws"] = _ffi.workspace(lib.ssd_loss_workspace_bytes(B, N, L))
        st["ad"] = torch.zeros((B, N, 4), dtype=torch.float32, device=dev)      # static target buffers (graph inputs)
        st["al"] = torch.zeros((B, N, L), dtype=torch.float32, device=dev)
        keep: List[Any] = []
        launches: List[Tuple[Any, tuple, str]] = []
        grad_of: Dict[int, torch.Tensor] = {}       # activation data_ptr -> fp16 gradient buffer
        written: Dict[int, bool] = {}

        def grad_buf(t: torch.Tensor) -> torch.Tensor:
            if t.data_ptr() not in grad_of:
                grad_of[t.data_ptr()] = torch.zeros_like(t)
                written[t.data_ptr()] = False
            return grad_of[t.data_ptr()]

        def add(fn, args, what, writes=()):
            launches.append((fn, args, what))
            for name in writes:                          # last launch that writes each variable's gradient
                last_writer[name] = len(launches) - 1

        last_writer: Dict[str, int] = {}

        first_conv_input = None
        for s in plan.steps:
            if s.kind == "cast":
                first_conv_input = s.keep[0].data_ptr()

        for s in reversed(plan.steps):
            mt = s.meta
            if s.kind == "conv":
                x, w16 = mt["x"], mt["w"]
                cout, k, stride, dil = mt["cout"], mt["k"], mt["stride"], mt["dilation"]
                Ho, Wo = mt["Ho"], mt["Wo"]
                H, W, cin = x.shape[1], x.shape[2], x.shape[3]
                ldy = _pad8(cout)
                if "head" in mt:
                    off, cnt, A = mt["head"]
                    dy = torch.zeros((B, Ho, Wo, ldy), dtype=torch.float16, device=dev)
                    add(lib.ssd_head_grad_gather, (_ffi.ptr(st["g_logits"]), _ffi.ptr(st["g_deltas"]), _ffi.ptr(dy), B, N, L,
                                                   off, Ho * Wo, A, ldy), s.name + ":gather")
                else:
                    out = mt["out0"]
                    dy = grad_of[out.data_ptr()]
                    assert written[out.data_ptr()], s.name
                    if mt["act"] != ACT_NONE:
                        add(lib.ssd_relu_bwd, (_ffi.ptr(dy), _ffi.ptr(out), out.numel()), s.name + ":relu")
                keep.append(dy)
                # filter + bias gradients
                d = ConvDesc()
                d.inp = x.data_ptr()
                d.B, d.H, d.W, d.Cin, d.Ho, d.Wo, d.Cout = B, H, W, cin, Ho, Wo, cout
                d.KH = d.KW = k
                d.stride, d.dilation, d.pad_top, d.pad_left = stride, dil, mt["ph"][0], mt["pw"][0]
                keep.append(d)
                add(lib.ssd_conv2d_wgrad, (C.byref(d), _ffi.ptr(dy), ldy, _ffi.ptr(self.vars[s.name + "/kernel"]["grad"])),
                    s.name + ":wgrad", writes=(s.name + "/kernel",))
                if s.name + "/bias" in self.vars:
                    add(lib.ssd_bias_grad, (_ffi.ptr(dy), _ffi.ptr(self.vars[s.name + "/bias"]["grad"]), B * Ho * Wo, ldy, cout),
                        s.name + ":bgrad", writes=(s.name + "/bias",))
                # data gradient (not needed for the layer fed by the image)
                if x.data_ptr() == first_conv_input:
                    continue
                wt = torch.zeros((cin, k, k, ldy), dtype=torch.float16, device=dev)
                add(lib.ssd_filter_flip_transpose, (_ffi.ptr(w16), _ffi.ptr(wt), cout, k, k, cin, ldy), s.name + ":flip")
                src, Hs, Ws = dy, Ho, Wo
                if stride > 1:
                    Hs, Ws = (Ho - 1) * stride + 1, (Wo - 1) * stride + 1
                    src = torch.zeros((B, Hs, Ws, ldy), dtype=torch.float16, device=dev)
                    add(lib.ssd_upsample_zero, (_ffi.ptr(dy), _ffi.ptr(src), B, Ho, Wo, ldy, Hs, Ws, stride), s.name + ":up")
                dx = grad_buf(x)
                acc = written[x.data_ptr()]
                g = ConvDesc()
                g.inp, g.weight, g.bias = src.data_ptr(), wt.data_ptr(), None
                g.residual = dx.data_ptr() if acc else None
                g.out0 = dx.data_ptr()
                g.B, g.H, g.W, g.Cin, g.Ho, g.Wo, g.Cout = B, Hs, Ws, ldy, H, W, cin
                g.KH = g.KW = k
                g.stride, g.dilation = 1, dil
                g.pad_top, g.pad_left = (k - 1) * dil - mt["ph"][0], (k - 1) * dil - mt["pw"][0]
                g.act, g.out_f32, g.split = ACT_NONE, 0, cin
                g.img_stride0, g.pix_stride0 = H * W * cin, cin
                keep += [wt, src, g]
                add(lib.ssd_conv2d, (C.byref(g),), s.name + ":dgrad")
                written[x.data_ptr()] = True
            elif s.kind == "bn":
                # y = act(BN(x)) (+ res):  dX of the normalisation into grad(x); the shortcut receives dY itself
                x, out, res = mt["x"], mt["out"], mt["res"]
                dy = grad_of[out.data_ptr()]
                assert written[out.data_ptr()], s.name
                dx = grad_buf(x)
                dres, acc_res = None, 0
                if res is not None:
                    dres = grad_buf(res)
                    acc_res = int(written[res.data_ptr()])
                    written[res.data_ptr()] = True
                add(lib.ssd_bn_train_bwd, (_ffi.ptr(x), _ffi.ptr(dy), _ffi.ptr(mt["gamma"]), _ffi.ptr(mt["beta"]),
                                           _ffi.ptr(mt["save"]), mt["M"], mt["C"], mt["act"], _ffi.ptr(dx), _ffi.ptr(dres),
                                           acc_res, _ffi.ptr(self.vars[s.name + "/gamma"]["grad"]),
                                           _ffi.ptr(self.vars[s.name + "/beta"]["grad"]), _ffi.ptr(mt["ws"]),
                                           mt["ws"].numel()), s.name + ":bn", writes=(s.name + "/gamma", s.name + "/beta"))
                written[x.data_ptr()] = True
            elif s.kind == "dw":
                x, out = mt["x"], mt["out"]
                dy = grad_of[out.data_ptr()]
                assert written[out.data_ptr()], s.name
                geo = (B, x.shape[1], x.shape[2], x.shape[3], mt["Ho"], mt["Wo"], mt["stride"], mt["ph"][0], mt["pw"][0])
                add(lib.ssd_depthwise3x3_wgrad, (_ffi.ptr(x), _ffi.ptr(dy),
                                                 _ffi.ptr(self.vars[s.name + "/depthwise_kernel"]["grad"])) + geo,
                    s.name + ":dw_wgrad", writes=(s.name + "/depthwise_kernel",))
                dx = grad_buf(x)
                acc = written[x.data_ptr()]
                add(lib.ssd_depthwise3x3_dgrad, (_ffi.ptr(dy), _ffi.ptr(mt["w"]), _ffi.ptr(dx)) + geo + (int(acc),),
                    s.name + ":dw_dgrad")
                written[x.data_ptr()] = True
            elif s.kind == "pool":
                x, out = mt["x"], mt["out"]
                dy = grad_of[out.data_ptr()]
                dx = grad_buf(x)
                acc = written[x.data_ptr()]
                add(lib.ssd_maxpool_bwd, (_ffi.ptr(x), _ffi.ptr(out), _ffi.ptr(dy), _ffi.ptr(dx), B, x.shape[1], x.shape[2],
                                          x.shape[3], out.shape[1], out.shape[2], mt["k"], mt["stride"], mt["ph"][0],
                                          mt["pw"][0], int(acc)), s.name + ":pool")
                written[x.data_ptr()] = True
            elif s.kind == "l2norm":
                x, out = mt["x"], mt["out"]
                dy = grad_of[out.data_ptr()]
                dx = grad_buf(x)
                acc = written[x.data_ptr()]
                add(lib.ssd_l2norm_bwd, (_ffi.ptr(x), _ffi.ptr(mt["scale"]), _ffi.ptr(dy), _ffi.ptr(dx),
                                         _ffi.ptr(self.vars[s.name + "/scale"]["grad"]), B * x.shape[1] * x.shape[2],
                                         x.shape[3], int(acc)), s.name + ":l2norm", writes=(s.name + "/scale",))
                written[x.data_ptr()] = True
        st["launches"], st["keep"], st["grad_of"] = launches, keep, grad_of
        # Segments for the overlapped all-reduce: bucket k is complete after the last launch that writes one of its
        # variables.  The backward pass runs in reverse layer order, so the LAST bucket completes first.
        ready_at: Dict[int, int] = {}
        for name, idx in last_writer.items():
            k = self.grads.bucket_of[name]
            ready_at[k] = max(ready_at.get(k, -1), idx)
        assert len(ready_at) == len(self.grads.buckets), "every bucket must be written by the backward pass"
        cuts = sorted(set(ready_at.values()))
        segments, begin = [], 0
        for c in cuts:
            segments.append((begin, c + 1, sorted(k for k, v in ready_at.items() if v == c)))
            begin = c + 1
        if begin < len(launches):
            segments.append((begin, len(launches), []))
        st["segments"] = segments                       # [(first launch, end launch, buckets complete afterwards)]
        st["graphs"] = None
        self._state[B] = st
        return st

    # -- one step ---------------------------------------------------------------------------------------------
    def _enqueue_step(self, st: Dict[str, Any], B: int, upto: Optional[int] = None) -> None:
        """Forward plan, fused loss forward/backward and the backward launch list on the current stream.  Nothing
        here allocates or synchronises, so the whole sequence is captured once per batch size into a CUDA graph
        (several hundred small launches per MobileNetV2 step would otherwise be bound by the host's launch rate)."""
        m, lib = self.model, self.lib
        plan = st["plan"]
        N, L = m.n_anchors, m.total_labels
        ad, al = st["ad"], st["al"]
        plan.run(parallel=False)
        stream = _ffi.stream()
        _ffi.check(lib.ssd_loss_fwd(_ffi.ptr(ad), _ffi.ptr(plan.deltas), _ffi.ptr(al), _ffi.ptr(plan.logits), B, N, L,
                                    self.neg_pos_ratio, self.alpha, 1, _ffi.ptr(st["loc"]), _ffi.ptr(st["conf"]),
                                    _ffi.ptr(st["ws"]), st["ws"].numel(), stream), "ssd_loss_fwd")
        # Keras reduces each loss with SUM_OVER_BATCH_SIZE (mean over the batch) and sums the two (trainer.py:91-94)
        _ffi.check(lib.ssd_loss_bwd(_ffi.ptr(ad), _ffi.ptr(plan.deltas), _ffi.ptr(al), _ffi.ptr(plan.logits), B, N, L,
                                    self.alpha, self.loss_scale / B, _ffi.ptr(st["g_deltas"]), _ffi.ptr(st["g_logits"]),
                                    _ffi.ptr(st["ws"]), st["ws"].numel(), stream), "ssd_loss_bwd")
        self.grads.zero_()
        self._enqueue_launches(st, 0, len(st["launches"]) if upto is None else upto)

    def _enqueue_launches(self, st: Dict[str, Any], first: int, last: int) -> None:
        stream = _ffi.stream()
        for fn, args, what in st["launches"][first:last]:
            rc = fn(*args, stream)
            if rc != 0:
                _ffi.check(rc, what)

    def _feed(self, plan, images: Any) -> None:
        """Training plans read the float32 image buffer: a uint8 batch is converted first (convert_image_dtype,
        utils/data_utils.py:36)."""
        if SSDModel._is_u8(images):
            images = _ffi.to_dev(images, dtype=torch.uint8).to(torch.float32) * (1.0 / 255.0)
        self.model._to_image_buffer(plan, images)

    def forward_backward(self, images: Any, actual_deltas: Any, actual_labels: Any, reduce: bool = True) -> Dict[str, torch.Tensor]:
        """Forward, loss, backward: fills the gradient buckets (loss-scaled) and returns the per-image losses.

        With several ranks (``reduce=True``) the step is replayed as one CUDA graph PER SEGMENT of the backward pass; after
        each segment the gradient buckets it completed are all-reduced asynchronously (NCCL's stream picks up behind the
        segment), so the exchange of the late layers' gradients overlaps the backward pass of the early layers.
        ``apply_gradients`` waits for the exchanges.  Single process: one graph, no exchange."""
        m = self.model
        B = int(images.shape[0])
        st = self._prepare(B)
        st["ad"].copy_(_ffi.to_dev(actual_deltas), non_blocking=True)
        st["al"].copy_(_ffi.to_dev(actual_labels), non_blocking=True)
        self._feed(st["plan"], images)
        self._eval_dirty = True
        distributed = reduce and dist_utils.world_size() > 1
        segments = st["segments"] if distributed else [(0, len(st["launches"]), [])]
        self._pending = []

        def run_segment(i):
            first, last, _ = segments[i]
            if i == 0:
                self._enqueue_step(st, B, upto=last)
            else:
                self._enqueue_launches(st, first, last)

        key = "graphs_dist" if distributed else "graphs"
        # launches recorded (or run eagerly) below carry no programmatic-dependent-launch edges: measured 7.85 vs 8.11 ms
        # per MobileNetV2 step (the inference plans keep them)
        self.lib.ssd_set_pdl(0)
        try:
            return self._forward_backward_recorded(st, B, key, segments, run_segment, distributed)
        finally:
            self.lib.ssd_set_pdl(-1)

    def _forward_backward_recorded(self, st, B, key, segments, run_segment, distributed) -> Dict[str, torch.Tensor]:
        if self.use_cuda_graph and st.get(key) is None:
            # the first step runs eagerly (workspaces, function attributes, split-K regions) ...
            for i in range(len(segments)):
                run_segment(i)
            torch.cuda.synchronize()
            graphs = []
            for i in range(len(segments)):           # ... and is then recorded, without executing, for every later step
                g = torch.cuda.CUDAGraph()
                with _no_gc_during_capture():
                    with torch.cuda.graph(g):
                        run_segment(i)
                graphs.append(g)
            st[key] = graphs
        for i, (_, _, done) in enumerate(segments):
            if self.use_cuda_graph:
                st[key][i].replay()
            else:
                run_segment(i)
            if distributed:
                for k in done:
                    self._pending.append(self.grads.allreduce_sum_async(k))
        return dict(loc=st["loc"], conf=st["conf"])

    def apply_gradients(self, learning_rate: Optional[float] = None, reduced: Optional[bool] = None) -> None:
        """Waits for the gradient exchange (sum over ranks; the 1 / world of the mean is folded into the optimizer's
        gradient scale), then ONE fused Adam launch over every variable, skipped on the device when a gradient is
        non-finite."""
        world = dist_utils.world_size()
        if self._pending:
            for w in self._pending:
                w.wait()
            self._pending = []
        elif world > 1 and reduced is not True:
            for w in [self.grads.allreduce_sum_async(k) for k in range(len(self.grads.buckets))]:
                w.wait()
        self.t += 1
        lr = self.lr if learning_rate is None else float(learning_rate)
        lr_t = lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        self.sumsq.zero_()
        st = _ffi.stream()
        _ffi.check(self.lib.ssd_grad_nonfinite_multi(_ffi.ptr(self._adam_table), len(self.vars), self._adam_max_n,
                                                     _ffi.ptr(self._guard), st), "ssd_grad_nonfinite_multi")
        _ffi.check(self.lib.ssd_adam_step_multi_guarded(_ffi.ptr(self._adam_table), len(self.vars), self._adam_max_n, lr_t,
                                                        self.b1, self.b2, self.eps, 1.0 / (self.loss_scale * world),
                                                        _ffi.ptr(self.sumsq), _ffi.ptr(self._guard), st),
                   "ssd_adam_step_multi_guarded")
        self._since_check += 1
        if self.dynamic_loss_scale and self._since_check >= self.scale_check_every:
            self.update_loss_scale()

    def update_loss_scale(self) -> float:
        """Host side of the loss-scale controller (one device synchronisation): halve after a skipped step, double after
        ``scale_growth_interval`` clean ones.  The scale is an argument of captured kernels, so a change drops the
        captured step graphs (they are re-captured on the next step)."""
        skipped = int(self._guard[1].item())
        self._guard[1].zero_()
        new = self.loss_scale
        if skipped > 0:
            self.skipped_steps += skipped
            self.t -= skipped                          # skipped steps did not advance Adam's bias correction
            self._clean_steps = 0
            new = max(self.loss_scale / 2.0, 1.0)
        else:
            self._clean_steps += self._since_check
            if self._clean_steps >= self.scale_growth_interval:
                new, self._clean_steps = min(self.loss_scale * 2.0, 65536.0), 0
        self._since_check = 0
        if new != self.loss_scale:
            self.loss_scale = new
            for st in self._state.values():
                st["graphs"] = st["graphs_dist"] = None
        return self.loss_scale

    def evaluate_batch(self, images: Any, targets: Tuple[Any, Any]) -> Dict[str, float]:
        """Validation loss of one batch (no gradient, no update): mean loc + mean conf (+ the regulariser of the
        last training step, like Keras adds ``model.losses`` to ``val_loss``)."""
        m, lib = self.model, self.lib
        B = int(images.shape[0])
        st = self._prepare(B)
        plan = st["plan"]
        if m.has_batchnorm:                          # Keras validates with training=False: moving statistics, folded BN
            if self._eval_dirty:
                self.sync_weights_to_host()
            plan = m.plan(B)
        ad, al = _ffi.to_dev(targets[0]), _ffi.to_dev(targets[1])
        self._feed(plan, images)
        plan.run()
        _ffi.check(lib.ssd_loss_fwd(_ffi.ptr(ad), _ffi.ptr(plan.deltas), _ffi.ptr(al), _ffi.ptr(plan.logits), B, m.n_anchors,
                                    m.total_labels, self.neg_pos_ratio, self.alpha, 1, _ffi.ptr(st["loc"]), _ffi.ptr(st["conf"]),
                                    _ffi.ptr(st["ws"]), st["ws"].numel(), _ffi.stream()), "ssd_loss_fwd")
        loc, conf = float(st["loc"].mean()), float(st["conf"].mean())
        reg = L2_REG * float(self.sumsq)
        return dict(loss=loc + conf + reg, loc_loss=loc, conf_loss=conf, reg_loss=reg)

    def train_step(self, images: Any, targets: Tuple[Any, Any], learning_rate: Optional[float] = None) -> torch.Tensor:
        """One optimisation step WITHOUT a host synchronisation: returns the step's total loss (mean loc + mean conf +
        regulariser, like Keras' ``loss``) as a device scalar.  ``fit`` accumulates these on the device and reads the
        epoch mean once."""
        out = self.forward_backward(images, targets[0], targets[1])
        self.apply_gradients(learning_rate)
        return out["loc"].mean() + out["conf"].mean() + L2_REG * self.sumsq[0]

    def train_on_batch(self, images: Any, targets: Tuple[Any, Any], learning_rate: Optional[float] = None) -> Dict[str, float]:
        """One optimisation step; ``targets = (actual_deltas, actual_labels)`` as ``train_utils.generator`` yields.
        Returns host floats (one synchronisation) like Keras' ``train_on_batch``."""
        out = self.forward_backward(images, targets[0], targets[1])
        self.apply_gradients(learning_rate)
        loc, conf = float(out["loc"].mean()), float(out["conf"].mean())
        reg = L2_REG * float(self.sumsq)
        return dict(loss=loc + conf + reg, loc_loss=loc, conf_loss=conf, reg_loss=reg)
