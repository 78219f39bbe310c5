"""Step machinery shared by the train loops of the reference's scripts (ACGAN/train.py, PGGAN/train.py,
Pix2Pix/train.py): two tf.train.AdamOptimizer instances over the `d_net` / `g_net` variable lists, each minimising its
own loss with the other network's variables left out of var_list."""
from __future__ import annotations

import math

import torch

from . import kernels as K
from .framework import get_store


class AdamState:
    """tf.train.AdamOptimizer(lr, beta1, beta2, epsilon) over one network's flat buffers (one fused launch)."""

    def __init__(self, flat, beta1=0.0, beta2=0.9, eps=1e-8):
        self.flat, self.beta1, self.beta2, self.eps = flat, beta1, beta2, eps
        self.t = 0
        self.lr_t = torch.zeros(1, dtype=torch.float32, device=flat.params.device)

    def set_lr(self, lr: float) -> None:
        """Advances the step count and uploads lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)."""
        self.t += 1
        self.lr_t.fill_(lr * math.sqrt(1.0 - self.beta2 ** self.t) / (1.0 - self.beta1 ** self.t))

    def apply(self, grad_scale: float = 1.0) -> None:
        K.adam(self.flat.params, self.flat.grads, self.flat.m, self.flat.v, self.lr_t, self.beta1, self.beta2,
               self.eps, grad_scale)


class TwoPlayer:
    """`minimize(loss, var_list=<root>_vars)` for the two root scopes of a GAN.  loss_fn() builds the forward pass on
    the tape and returns the loss Var; gradients of the other root are not computed (store.frozen_scopes)."""

    def __init__(self, d_root: str = "d_net", g_root: str = "g_net", beta1: float = 0.0, beta2: float = 0.9,
                 eps: float = 1e-8, world_size: int = 1, grad_allreduce=None, store=None):
        self.store = store or get_store()
        self.roots = {"d": d_root, "g": g_root}
        self.world_size, self.grad_allreduce = world_size, grad_allreduce
        self.betas = (beta1, beta2, eps)
        self.opt = {}

    def finalize(self) -> None:
        """Call once every variable exists: flat parameter / gradient / Adam-slot buffers per network."""
        self.store.finalize()
        for k, root in self.roots.items():
            self.opt[k] = AdamState(self.store.flat[root], *self.betas)

    def gradients(self, which: str, loss_fn):
        """Zeroes the gradients of network `which`, runs loss_fn() on a tape with the other network frozen and
        back-propagates.  Returns the loss Var."""
        st = self.store
        root, other = self.roots[which], self.roots["g" if which == "d" else "d"]
        st.zero_grad(root)
        with st.gradient_tape() as tape, st.frozen_scopes(other):
            loss = loss_fn()
            tape.backward(loss)
        return loss

    def step(self, which: str, loss_fn, lr: float):
        loss = self.gradients(which, loss_fn)
        root = self.roots[which]
        if self.grad_allreduce is not None:
            self.grad_allreduce(self.store.flat[root].grads)
        self.opt[which].set_lr(lr)
        self.opt[which].apply(1.0 / self.world_size)
        self.store.bump(root)
        group = self.store.pack_groups.get(root)
        if group is not None and group.entries:
            group.refresh()          # bf16 operand copies follow the update
        return loss
