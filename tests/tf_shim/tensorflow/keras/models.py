"""tf.keras.models.Model: evaluates a graph recorded by the layer stand-ins."""

from __future__ import annotations

from typing import Any, Dict, List

import numpy as np

import tensorflow as tf
from .layers import Layer, SymbolicTensor


class Model(Layer):
    def __init__(self, inputs: Any = None, outputs: Any = None, name: str = "model", **kwargs: Any):
        super().__init__(name=name)
        self.inputs_, self.outputs_ = inputs, outputs
        self.input, self.output = inputs, outputs
        self.layers: List[Layer] = []
        seen = set()

        def visit(node):
            if isinstance(node, (list, tuple)):
                for n in node:
                    visit(n)
                return
            if id(node) in seen or not isinstance(node, SymbolicTensor):
                return
            seen.add(id(node))
            if node.layer is not None:
                visit(node.inputs)
                if node.layer not in self.layers:
                    self.layers.append(node.layer)

        visit(outputs)

    def get_layer(self, name: str) -> Layer:
        for l in self.layers:
            if l.name == name:
                return l
        raise ValueError(f"No such layer: {name}")

    def __call__(self, x: Any, training: bool = False) -> Any:
        feed = tf.convert_to_tensor(x)
        memo: Dict[int, Any] = {}
        taps: Dict[str, Any] = {}

        def ev(node):
            if isinstance(node, (list, tuple)):
                return [ev(n) for n in node]
            if id(node) in memo:
                return memo[id(node)]
            if node.layer is None:
                if node is not self.inputs_:
                    raise ValueError("graph input is not the model input")
                val = feed
            else:
                val = node.layer._run(ev(node.inputs))
                taps[node.layer.name] = val
            memo[id(node)] = val
            return val

        out = ev(self.outputs_)
        self.last_activations = taps
        return out

    def predict(self, x: Any, steps=None, verbose=0):
        out = self(x)
        return [o.numpy() for o in out] if isinstance(out, list) else out.numpy()

    def named_variables(self) -> Dict[str, tf.Variable]:
        """``{"<layer>/<variable>": Variable}`` -- the names an .h5 checkpoint would use."""
        out = {}
        for l in self.layers:
            for k, v in l.variables().items():
                out[f"{l.name}/{k}"] = v
            if not l.variables() and hasattr(l, "scale") and isinstance(getattr(l, "scale"), tf.Variable):
                out[f"{l.name}/scale"] = l.scale      # custom layers that keep a tf.Variable attribute (L2Normalization)
        return out
