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
        self._graphs, self.graph_launches, self.graph_loss = {}, {}, {}

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
        self._update(which)          # Adam, then the bf16 operand copies follow the update
        return loss

    # ------------------------------------------------------------------------------------------ CUDA graphs
    def _update(self, which: str) -> None:
        root = self.roots[which]
        self.opt[which].apply(1.0 / self.world_size)
        self.store.bump(root)
        group = self.store.pack_groups.get(root)
        if group is not None and group.entries:
            group.refresh()

    def _invalidate_sn(self) -> None:
        """Every captured compute graph re-evaluates the spectral-norm state it uses (the host-side version cache
        does not run at replay)."""
        st = self.store
        groups = list(st.sn_groups.values()) + [g for shadows in st.sn_shadow.values() for g in shadows]
        for g in groups:
            g.valid_for = None
            g.fresh_for = None

    def capture(self, which: str, loss_fn) -> None:
        """Captures step(which, loss_fn, lr) into CUDA graphs.  loss_fn must read its inputs from tensors that keep
        their addresses (the trainer's static buffers), and at least one eager step of this kind must have run
        (workspaces, descriptor tables and operand copies exist).  With a gradient collective the compute and update
        halves are captured separately and the all-reduce runs between them at replay; the learning rate stays a
        device scalar written by AdamState.set_lr before each replay."""
        st = self.store
        for root in self.roots.values():          # operand copies are current before any compute graph runs
            group = st.pack_groups.get(root)
            if group is not None and group.entries:
                group.refresh()
        loss_out = torch.zeros(1, dtype=torch.float32, device=st.device)

        def compute():
            loss = self.gradients(which, loss_fn)
            loss_out.copy_(loss.data.reshape(-1)[:1])

        def full():
            compute()
            self._update(which)

        parts = (("full", full),) if self.grad_allreduce is None else (("compute", compute),
                                                                        ("update", lambda: self._update(which)))
        torch.cuda.synchronize()
        for name, body in parts:
            self._invalidate_sn()
            before = K.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body()
            self._graphs[(which, name)] = g
            self.graph_launches[(which, name)] = K.launch_count() - before
        self._invalidate_sn()
        self.graph_loss[which] = loss_out

    def captured(self, which: str) -> bool:
        return which in self.graph_loss

    def replay(self, which: str, lr: float) -> torch.Tensor:
        """One captured step of network `which`; returns the (static) loss scalar."""
        self.opt[which].set_lr(lr)
        if (which, "full") in self._graphs:
            self._graphs[(which, "full")].replay()
        else:
            self._graphs[(which, "compute")].replay()
            self.grad_allreduce(self.store.flat[self.roots[which]].grads)
            self._graphs[(which, "update")].replay()
        return self.graph_loss[which]

    def launches(self, which: str) -> int:
        return sum(n for (w, _), n in self.graph_launches.items() if w == which)


def reference_loop(trainer, max_iter: int, step_fn, scalars_fn, dev_cost_fn, samples_fn, *, out_dir: str = '.',
                   checkpoint_dir: str | None = None, display_interval: int = 100, out_image_interval: int = 1000,
                   restore: bool = False, capture_after: int | None = 1, capture_fn=None, log=print):
    """The body shared by ACGAN/train.py:176-233 and PGGAN/train.py:168-226: optional restore of the latest checkpoint
    (optimistic_restore), then per step `step_fn(step)` (the script's G step / n_dis D steps), a progress line every
    display_interval steps, and every out_image_interval steps the dev-set critic loss (`lib.plot.plot('dev_cost')`),
    the fixed-noise sample grid (`samples_<step>.png`) and `saver.save(... 'model.ckpt', global_step=step)`;
    `lib.plot.tick()` closes every step.  The Inception-score hook (evaluation_interval) needs the external Inception
    graph and is not run.  capture_after: the step after which capture_fn() turns the training ops into CUDA graphs."""
    import os

    from .common import misc as lib_misc
    from .common import plot as lib_plot

    checkpoint_dir = checkpoint_dir or os.path.join(out_dir, 'checkpoint')
    lib_plot.set_output_dir(out_dir)
    opts = (trainer.players.opt['g'], trainer.players.opt['d'])          # g_opt is created first in both scripts
    if restore:
        ckpts = [f for f in os.listdir(checkpoint_dir) if f.startswith('model.ckpt-')] if os.path.isdir(checkpoint_dir) else []
        if ckpts:
            latest = max(ckpts, key=lambda f: int(f.split('-')[1].split('.')[0]))
            log('Restore model from: {}...'.format(latest))
            prefix = latest.split('.npz')[0].split('.index')[0].split('.data-')[0]     # .npz or a TF tensor bundle
            restored_path = os.path.join(checkpoint_dir, prefix)
            lib_misc.restore_checkpoint(restored_path, opts)
            # the trainers' global_step (TF name `Variable`) drives the polynomial learning-rate decay of ACGAN / Pix2Pix
            if hasattr(trainer, 'global_step') and os.path.exists(restored_path + '.npz'):
                import numpy as np
                with np.load(restored_path + '.npz') as z:
                    if 'Variable' in z.files:
                        trainer.global_step = int(z['Variable'])
        else:
            log('No checkpoint found in: {}'.format(checkpoint_dir))
    captured = False
    for step in range(max_iter):
        step_fn(step)
        if capture_fn is not None and capture_after is not None and not captured and step >= capture_after:
            capture_fn()
            captured = True
        if step % display_interval == display_interval - 1:
            log('step: {}, '.format(step) + ', '.join('{}: {}'.format(k, float(v.reshape(-1)[0]))
                                                      for k, v in scalars_fn().items()))
        if step % out_image_interval == out_image_interval - 1:
            costs = dev_cost_fn(step)
            if costs:
                lib_plot.plot('dev_cost', torch.stack([c.reshape(-1)[0] for c in costs]).mean())
                lib_plot.flush()
            lib_misc.save_images(samples_fn(step), os.path.join(out_dir, 'samples_{}.png'.format(step)))
            if not os.path.exists(checkpoint_dir):
                os.mkdir(checkpoint_dir)
            extra = {'Variable': int(trainer.global_step)} if hasattr(trainer, 'global_step') else None
            lib_misc.save_checkpoint(os.path.join(checkpoint_dir, 'model.ckpt-{}'.format(step)), opts, extra=extra)
        lib_plot.tick()
    return trainer
