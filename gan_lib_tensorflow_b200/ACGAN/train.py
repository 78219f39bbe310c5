"""Training steps of ACGAN/train.py:78-146, 191-203 on the B200 layer ops (config 2 of BASELINE.json): critic loss
get_loss(loss_type) + auxiliary-classifier cross-entropy on the real batch, generator loss get_loss + acgan_scale_G *
cross-entropy on the fake batch, two Adam(beta1 = 0, beta2 = 0.9) optimisers on
tf.train.polynomial_decay(4e-4 -> 2e-4 over max_iter / 2 generator steps); per iteration one generator step (skipped
at step 0) and n_dis critic steps.

The WGAN-GP gradient penalty the script always adds (train.py:97-105) is ACGAN/gp.py: a third pass of D over the
interpolates whose input gradient is built from differentiable backward ops, so the penalty reaches D's parameters
through the backward pass (grad-grad of the batch norms included).  gradient_penalty=False drops the term."""
from __future__ import annotations

import numpy as np
import torch

from .. import functional as F
from .. import kernels as K
from ..framework import Var, get_store
from ..training import TwoPlayer
from . import gp as GP
from . import model as M


class Trainer:
    def __init__(self, batch_size: int = 64, z_dim: int = 128, loss_type: str = 'HINGE', acgan_scale_G: float = 0.1,
                 max_iter: int = 100000, gradient_penalty: bool = True, seed: int | None = 0, world_size: int = 1,
                 grad_allreduce=None):
        self.gradient_penalty = bool(gradient_penalty)
        self.store = get_store()
        self.model = M.ACGAN()
        self.batch, self.z_dim, self.loss_type, self.scale_g = batch_size, z_dim, loss_type, acgan_scale_G
        self.max_iter, self.global_step = max_iter, 0
        if seed is not None:
            np.random.seed(seed)
        dev = self.store.device
        lab = torch.zeros(2, dtype=torch.int32, device=dev)
        with self.store.building():          # train.py:89-94: D(real), G, D(fake, reuse)
            self.model.get_discriminator(torch.zeros(2, 32, 32, 3, device=dev), lab, update_collection=None)
            fake = self.model.get_generator(torch.zeros(2, z_dim, device=dev), lab)
            self.model.get_discriminator(fake, lab, update_collection='NO_OPS', reuse=True)
        self.players = TwoPlayer("d_net", "g_net", beta1=0.0, beta2=0.9, world_size=world_size,
                                 grad_allreduce=grad_allreduce)
        self.players.finalize()

    def learning_rate(self) -> float:
        """tf.train.polynomial_decay(0.0004, global_step, max_iter // 2, 0.0002) (train.py:141)"""
        steps = self.max_iter // 2
        t = min(self.global_step, steps) / float(steps)
        return (0.0004 - 0.0002) * (1.0 - t) + 0.0002

    def preprocess(self, real_int, deq_noise):
        """int [B, 3072] CHW -> NHWC float in [-1, 1) + U(0, 1/128) (train.py:78-81)"""
        b = real_int.shape[0]
        return K.preprocess_real(real_int, deq_noise, b, 1024).reshape(b, 32, 32, 3)

    def d_loss(self, real, real_labels, z, fake_labels, alpha=None):
        """alpha: fp32 [batch] interpolation coefficients of the penalty (tf.random_uniform, train.py:98); drawn here
        when not given."""
        m = self.model
        fake = m.get_generator(z, fake_labels, reuse=True)                                   # g_net frozen
        disc_real, disc_real_acgan = m.get_discriminator(Var(real), real_labels, update_collection=None, reuse=True)
        disc_fake, _ = m.get_discriminator(Var(fake.data), fake_labels, update_collection='NO_OPS', reuse=True)
        loss, self.last_d = M.discriminator_losses(disc_real, disc_real_acgan, real_labels, disc_fake, self.loss_type)
        if self.gradient_penalty:
            if alpha is None:
                alpha = torch.rand(real.shape[0], device=real.device)
            fake32 = fake.data if fake.data.dtype == torch.float32 else K.cast(fake.data, torch.float32)
            gp = GP.gradient_penalty(real, fake32, alpha)
            self.last_d['gradient_penalty'] = gp.data
            loss = F.add_scalars(loss, gp)
        return loss

    def g_loss(self, z, fake_labels):
        m = self.model
        fake = m.get_generator(z, fake_labels, reuse=True)
        disc_fake, disc_fake_acgan = m.get_discriminator(fake, fake_labels, update_collection='NO_OPS', reuse=True)
        loss, self.last_g = M.generator_losses(disc_fake, disc_fake_acgan, fake_labels, self.loss_type, self.scale_g)
        return loss

    def d_step(self, real, real_labels, z, fake_labels, alpha=None):
        if self.players.captured("d"):
            s = self.static
            s['real'].copy_(real, non_blocking=True)
            s['real_labels'].copy_(real_labels, non_blocking=True)
            s['z_d'].copy_(z, non_blocking=True)
            s['fake_labels_d'].copy_(fake_labels, non_blocking=True)
            if alpha is None:
                s['alpha'].uniform_(0.0, 1.0)
            else:
                s['alpha'].copy_(alpha, non_blocking=True)
            return self.players.replay("d", self.learning_rate())
        return self.players.step("d", lambda: self.d_loss(real, real_labels, z, fake_labels, alpha),
                                 self.learning_rate())

    def g_step(self, z, fake_labels):
        if self.players.captured("g"):
            self.static['z_g'].copy_(z, non_blocking=True)
            self.static['fake_labels_g'].copy_(fake_labels, non_blocking=True)
            loss = self.players.replay("g", self.learning_rate())
        else:
            loss = self.players.step("g", lambda: self.g_loss(z, fake_labels), self.learning_rate())
        self.global_step += 1                # g_opt.minimize(..., global_step=global_step)
        return loss

    def capture(self):
        """Both training ops as CUDA graphs over static input buffers (the step is launch-bound at CIFAR size: ~500
        kernels of a few microseconds each).  Call after at least one eager d_step and g_step (every workspace,
        descriptor table and operand copy exists); afterwards d_step / g_step copy their arguments into the static
        buffers and replay."""
        dev, b = self.store.device, self.batch
        f32, i32 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.int32, device=dev)
        s = self.static = {
            'real': torch.zeros(b, 32, 32, 3, **f32), 'real_labels': torch.zeros(b, **i32),
            'z_d': torch.zeros(b, self.z_dim, **f32), 'fake_labels_d': torch.zeros(b, **i32),
            'alpha': torch.zeros(b, **f32), 'z_g': torch.zeros(b, self.z_dim, **f32),
            'fake_labels_g': torch.zeros(b, **i32),
        }
        d_fn = lambda: self.d_loss(s['real'], s['real_labels'], s['z_d'], s['fake_labels_d'], s['alpha'])  # noqa: E731
        g_fn = lambda: self.g_loss(s['z_g'], s['fake_labels_g'])                                             # noqa: E731
        self.players.capture("d", d_fn)
        self.players.capture("g", g_fn)

    def _noise(self):
        dev = self.store.device
        z = torch.randn(self.batch, self.z_dim, device=dev)
        labels = (torch.rand(self.batch, device=dev) * 10).to(torch.int32)
        return z, labels

    def train_iteration(self, step: int, batches, n_dis: int = 5):
        """train.py:191-203.  `batches` yields (NHWC float real images, int32 labels) on the device."""
        g = None
        if step > 0:
            g = self.g_step(*self._noise())
        d = None
        for _ in range(n_dis):
            real, labels = next(batches)
            d = self.d_step(real, labels, *self._noise())
        return d, g

    # ------------------------------------------------------------------------------------------ the script's loop
    def train(self, max_iter: int, train_gen, dev_gen=None, n_dis: int = 5, **kw):
        """ACGAN/train.py:148-233.  train_gen / dev_gen: epoch generator factories yielding (int pixels [B, 3072] CHW,
        int labels [B]) like common.data.cifar10.load; the G step is skipped at step 0 (:192), then n_dis critic steps;
        fixed-noise samples use labels 0..9 repeated (:150-152).  Remaining keywords: training.reference_loop."""
        from ..common import misc as lib_misc
        from ..training import reference_loop

        dev = self.store.device
        fixed_z = torch.from_numpy(lib_misc.get_z(100, n_hidden=self.z_dim)).to(dev)                # :148
        fixed_labels = torch.from_numpy(np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9] * 10, dtype='int32')).to(dev)

        def inf_train_gen():
            while True:
                for images_, labels_ in train_gen():
                    yield images_, labels_

        gen = inf_train_gen()

        def feed(images, labels):
            real_int = torch.as_tensor(images, dtype=torch.int32).to(dev)
            deq = torch.empty(real_int.shape[0], real_int.shape[1], device=dev).uniform_(0.0, 1.0 / 128)   # :80-81
            return self.preprocess(real_int, deq), torch.as_tensor(labels, dtype=torch.int32).to(dev)

        def step_fn(step):
            self.train_iteration(step, (feed(*next(gen)) for _ in iter(int, 1)), n_dis=n_dis)

        def scalars():
            out = {}
            if getattr(self, 'last_g', None):      # the kernel folds acgan_scale_G into the term; the script logs it unscaled
                out.update({'g_loss_gan': self.last_g['g_loss_gan'],
                            'g_loss_acgan': self.last_g['g_loss_acgan_scaled'] / self.scale_g})
            d_gan = self.last_d['d_loss_gan']
            if self.last_d.get('gradient_penalty') is not None:     # the script logs d_loss_gan after `+= gradient_penalty`
                d_gan = d_gan + self.last_d['gradient_penalty']
            out.update({'d_loss_gan': d_gan, 'd_loss_acgan': self.last_d['d_loss_acgan']})
            return out

        def dev_costs(_step):
            if dev_gen is None:
                return []
            # d_loss() rebinds self.last_d to fresh tensors; the captured graphs keep writing the training scalars into the
            # ones they were captured with, so those are put back (otherwise every later progress line would show the last
            # dev batch)
            keep = self.last_d
            try:
                return [self.d_loss(*feed(images, labels), *self._noise()).data.clone() for images, labels in dev_gen()]
            finally:
                self.last_d = keep

        def samples(_step):
            return self.model.get_generator(fixed_z, fixed_labels, reuse=True).data

        return reference_loop(self, max_iter, step_fn, scalars, dev_costs, samples, capture_fn=self.capture, **kw)
