"""Minimal stand-in for the TF-1 variable machinery the reference relies on (oracle side).

The reference addresses all state by scope-qualified variable name created through
tf.variable_scope(name, reuse=) / tf.get_variable(name, ...) (common/ops/conv2d.py:59,142,213;
common/ops/linear.py:45,140,177; common/ops/sn.py:28,32; common/ops/normalization.py:43,49,51;
common/ops/embedding.py:28,40).  This shim reproduces the naming and reuse rules with torch tensors:

* get_variable(name, initializer) under nested scopes yields "a/b/name";
* when the variable already exists it is returned unchanged (the initializer value is discarded, exactly
  as the reference draws and discards NumPy values on every reuse call, conv2d.py:124-144);
* trainable flags are recorded so that trainable_variables() can be filtered by substring
  (SNGAN/gan_cifar_resnet.py:507,512).
"""
from __future__ import annotations

import contextlib
from collections import OrderedDict

import numpy as np
import torch


class Graph:
    """Holds variables (name -> tensor) and the scope stack."""

    def __init__(self, dtype=torch.float64, u_seed=2):
        self.dtype = dtype
        self.u_rng = np.random.RandomState(u_seed)  # stands in for TF's op-level RNG for `u` (sn.py:32)
        self.vars: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        self.trainable: "OrderedDict[str, bool]" = OrderedDict()
        self._scopes: list[str] = []
        self.collections: dict[str, list] = {}
        # True while the "graph" is being built: like the reference, every layer call then draws its NumPy
        # initial values even when the variable already exists (conv2d.py:124-144).  Training-time calls of
        # the eager oracle correspond to session.run on the already-built graph and draw nothing.
        self.draw_on_reuse = False

    # -- scopes ---------------------------------------------------------------------------------
    @contextlib.contextmanager
    def variable_scope(self, name, reuse=None):
        self._scopes.append(name)
        try:
            yield
        finally:
            self._scopes.pop()

    def full_name(self, name: str) -> str:
        return "/".join([s for s in self._scopes if s] + [name])

    # -- variables ------------------------------------------------------------------------------
    def get_variable(self, name, initializer=None, shape=None, trainable=True):
        full = self.full_name(name)
        if full in self.vars:
            if self.draw_on_reuse and callable(initializer):
                initializer(shape)  # drawn and discarded
            return self.vars[full]
        if initializer is None:
            raise ValueError(f"variable {full} needs an initializer")
        if callable(initializer):
            value = initializer(shape)
        else:
            value = initializer
        t = torch.as_tensor(np.asarray(value), dtype=self.dtype).clone()
        t.requires_grad_(bool(trainable))
        self.vars[full] = t
        self.trainable[full] = bool(trainable)
        return t

    def trainable_variables(self, substring: str = ""):
        return [(n, v) for n, v in self.vars.items() if self.trainable[n] and substring in n]

    def assign(self, name_or_tensor, value):
        """u.assign(value) -- in-place, outside autograd."""
        t = self.vars[name_or_tensor] if isinstance(name_or_tensor, str) else name_or_tensor
        with torch.no_grad():
            t.copy_(value.detach().reshape(t.shape))

    def add_to_collection(self, key, op):
        self.collections.setdefault(key, []).append(op)


def constant_initializer(value):
    return lambda shape: np.full(shape, value, dtype="float32")


def truncated_normal(shape, rng: np.random.RandomState):
    """tf.truncated_normal_initializer(): N(0,1) re-drawn while |x| > 2.

    TF's own random stream cannot be reproduced; parity tests inject `u` explicitly."""
    out = rng.standard_normal(size=shape)
    bad = np.abs(out) > 2
    while bad.any():
        out[bad] = rng.standard_normal(size=int(bad.sum()))
        bad = np.abs(out) > 2
    return out.astype("float32")
