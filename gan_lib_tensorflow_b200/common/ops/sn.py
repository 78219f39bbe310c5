"""Spectral normalisation with the reference's signature (common/ops/sn.py:8-69).

The returned object stands for W / sigma; Conv2D / Linear consume it without materialising the division
(1/sigma is applied in the GEMM epilogue).  u state and update_collection semantics follow sn.py:48-65:
None -> u <- u' on every evaluation of the network; "NO_OPS" -> nothing stored; any other string -> the
assignment is appended to that collection for the caller to run.
"""
from __future__ import annotations

import warnings

import torch

from ... import kernels as K
from ...framework import Var, Variable, get_store, truncated_normal

NO_OPS = 'NO_OPS'

_collections: dict[str, list] = {}
_warned = False


class SNWeight:
    """W_bar = W / sigma as (weight variable, spectral-norm state)."""

    def __init__(self, w: Variable, entry, sigma):
        self.w, self.entry, self.sigma = w, entry, sigma

    def materialize(self) -> torch.Tensor:
        """fp32 copy of W / sigma (diagnostics and tests; the layers never need it)."""
        out = torch.empty_like(self.w.data)
        flat = out.reshape(-1, 1)
        K.sgemm_small(self.w.data.reshape(-1, 1), torch.ones(1, 1, device=out.device), flat, flat.shape[0], 1, 1,
                      False, False, self.entry.inv_sigma, None, 0.0)
        return out


def get_collection(key):
    return _collections.get(key, [])


def run_update_collection(key):
    """Executes the u assignments a caller collected under `key` (sn.py:65)."""
    for u_var, value in _collections.pop(key, []):
        u_var.data.copy_(value)
        get_store().bump_u(u_var.root)


def spectral_normed_weight(W, u=None, num_iters=1, update_collection=None, with_sigma=False, reuse=False):
    """common/ops/sn.py:15-69"""
    if num_iters != 1:
        raise NotImplementedError("only num_iters=1 (the value every reference call-site uses) is built")
    store = get_store()
    with store.variable_scope('spectral_norm'):
        c = W.data.shape[-1]
        if u is None:
            u = store.get_variable("u", shape=[1, c], trainable=False,
                                   initializer=lambda s: truncated_normal(s, store.u_rng))
        elif not isinstance(u, Variable):
            raise TypeError("u must be a framework.Variable (persistent state)")
    entry = store.sn_acquire(W, u, assign=update_collection is None)
    if update_collection is None:
        global _warned
        if not _warned:
            warnings.warn('Setting update_collection to None will make u being updated every W execution. This '
                          'maybe undesirable. Please consider using a update collection instead.', stacklevel=2)
            _warned = True
    else:
        if update_collection != NO_OPS:
            _collections.setdefault(update_collection, []).append((u, entry.u_out.clone().reshape(u.data.shape)))
    w_bar = SNWeight(W, entry, entry.sigma)
    if with_sigma:
        return w_bar, Var(entry.sigma)
    return w_bar
