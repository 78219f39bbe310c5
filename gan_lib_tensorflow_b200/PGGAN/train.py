"""Training steps of PGGAN/train.py:83-136, 182-190 on the B200 layer ops: hinge losses, two Adam(1e-4, beta1 = 0,
beta2 = 0.9) optimisers, per iteration one generator step followed by n_dis critic steps, alpha = step / max_iter.
Real images arrive as NHWC float tensors already resized to the stage's resolution (tf.image.resize_images,
train.py:89-93, is an input-pipeline step outside the hot path)."""
from __future__ import annotations

import numpy as np
import torch

from .. import functional as F
from ..framework import Var, get_store
from ..training import TwoPlayer
from .model_nvidia import PGGAN

LR = 0.0001      # train.py:125, 128
N_DIS = 5        # --n_dis
Z_DIM = 512      # --z_dim


class Trainer:
    def __init__(self, block_count: int, trans: bool, inputs_norm: bool = False, batch_size: int = 16,
                 z_dim: int = Z_DIM, max_iter: int = 100000, seed: int | None = 0, world_size: int = 1,
                 grad_allreduce=None, model: str = "nvidia"):
        self.store = get_store()
        if model == "nvidia":                    # train.py:61-66 (--model)
            self.model = PGGAN(block_count=block_count, trans=trans, inputs_norm=inputs_norm)
        elif model == "resnet":
            from .model_resnet import PGGAN as PGGANResNet
            self.model = PGGANResNet(block_count=block_count, trans=trans, inputs_norm=inputs_norm)
        else:
            raise NotImplementedError('Not supported model!')
        self.batch, self.z_dim, self.max_iter = batch_size, z_dim, max_iter
        self.size = 4 * 2 ** block_count
        if seed is not None:
            np.random.seed(seed)
        dev = self.store.device
        # graph construction order of train.py:103-107: D(real), G, D(fake, reuse)
        with self.store.building():
            real0 = torch.zeros(2, self.size, self.size, 3, device=dev)
            self.model.get_discriminator(real0, 0.0, update_collection="NO_OPS")
            fake0 = self.model.get_generator(torch.zeros(2, z_dim, device=dev), 0.0)
            self.model.get_discriminator(fake0, 0.0, update_collection="NO_OPS", reuse=True)
        self.players = TwoPlayer("d_net", "g_net", beta1=0.0, beta2=0.9, world_size=world_size,
                                 grad_allreduce=grad_allreduce)
        self.players.finalize()

    def alpha(self, step: int) -> float:
        return (step * 1.0) / self.max_iter          # train.py:184

    # ------------------------------------------------------------------------------------------ losses
    def d_loss(self, real, z, alpha):
        """train.py:103-113: D(real) assigns u (update_collection=None), D(G(z)) runs with NO_OPS; G gets no gradient."""
        m = self.model
        fake = m.get_generator(z, alpha, reuse=True)                      # no tape entry survives: g_net is frozen
        disc_real = m.get_discriminator(Var(real), alpha, update_collection=None, reuse=True)
        disc_fake = m.get_discriminator(Var(fake.data), alpha, update_collection="NO_OPS", reuse=True)
        return F.gan_loss(F.concat_rows(disc_real, disc_fake), 'd', n_real=real.shape[0], loss_type="HINGE")

    def g_loss(self, z, alpha):
        m = self.model
        fake = m.get_generator(z, alpha, reuse=True)
        disc_fake = m.get_discriminator(fake, alpha, update_collection="NO_OPS", reuse=True)
        return F.gan_loss(disc_fake, 'g', loss_type="HINGE")

    # ------------------------------------------------------------------------------------------ steps
    def d_step(self, real, z, alpha):
        if self.players.captured("d"):
            s = self.static
            s['real'].copy_(real, non_blocking=True)
            s['z_d'].copy_(z, non_blocking=True)
            s['alpha'].fill_(float(alpha))
            return self.players.replay("d", LR)
        return self.players.step("d", lambda: self.d_loss(real, z, alpha), LR)

    def g_step(self, z, alpha):
        if self.players.captured("g"):
            self.static['z_g'].copy_(z, non_blocking=True)
            self.static['alpha'].fill_(float(alpha))
            return self.players.replay("g", LR)
        return self.players.step("g", lambda: self.g_loss(z, alpha), LR)

    def capture(self):
        """Both training ops as CUDA graphs over static buffers (call after one eager d_step and g_step).  alpha lives
        in a device scalar that functional.lerp reads at run time (the placeholder of train.py:83), so the graphs
        follow alpha = step / max_iter without being re-captured."""
        dev = self.store.device
        f32 = dict(dtype=torch.float32, device=dev)
        s = self.static = {'real': torch.zeros(self.batch, self.size, self.size, 3, **f32),
                           'z_d': torch.zeros(self.batch, self.z_dim, **f32),
                           'z_g': torch.zeros(self.batch, self.z_dim, **f32), 'alpha': torch.zeros(1, **f32)}
        self.players.capture("d", lambda: self.d_loss(s['real'], s['z_d'], s['alpha']))
        self.players.capture("g", lambda: self.g_loss(s['z_g'], s['alpha']))

    def train_iteration(self, step: int, batches, n_dis: int = N_DIS):
        """train.py:182-190.  `batches` yields NHWC float real images [batch, size, size, 3]."""
        a = self.alpha(step)
        dev = self.store.device
        g = self.g_step(torch.randn(self.batch, self.z_dim, device=dev), a)
        d = None
        for _ in range(n_dis):
            d = self.d_step(next(batches), torch.randn(self.batch, self.z_dim, device=dev), a)
        return d, g

    # ------------------------------------------------------------------------------------------ the script's loop
    def train(self, train_gen, dev_gen=None, max_iter: int | None = None, n_dis: int = N_DIS, **kw):
        """PGGAN/train.py:138-226.  train_gen / dev_gen: epoch generator factories yielding NHWC float images
        [batch, size, size, 3] already resized to this stage (train.py:89-93 is an input-pipeline step) -- labels, if a
        tuple is yielded, are ignored like in the reference; every step runs the G step first (:183), then n_dis critic
        steps, all with alpha = step / max_iter.  Remaining keywords: training.reference_loop."""
        from ..common import misc as lib_misc
        from ..training import reference_loop

        if max_iter is not None:
            self.max_iter = max_iter                 # alpha = step / args.max_iter (:184)
        max_iter = self.max_iter
        dev = self.store.device
        fixed_z = torch.from_numpy(lib_misc.get_z(100, n_hidden=self.z_dim)).to(dev)                # :138

        def images_of(item):
            x = item[0] if isinstance(item, (tuple, list)) else item
            return torch.as_tensor(x, dtype=torch.float32).to(dev)

        def inf_train_gen():
            while True:
                for item in train_gen():
                    yield images_of(item)

        gen = inf_train_gen()
        last = {}

        def step_fn(step):
            d, g = self.train_iteration(step, gen, n_dis=n_dis)
            last['d_loss'], last['g_loss'] = d.data if hasattr(d, 'data') else d, g.data if hasattr(g, 'data') else g

        def z():
            return torch.randn(self.batch, self.z_dim, device=dev)

        def dev_costs(step):
            if dev_gen is None:
                return []
            # eager evaluations between graph replays: the replays move the weights behind the host-side spectral-norm
            # cache, so it is dropped before and after (D(real) assigns u here too, like the reference's dev loop)
            self.players._invalidate_sn()
            costs = [self.d_loss(images_of(item), z(), self.alpha(step)).data.clone() for item in dev_gen()]
            self.players._invalidate_sn()
            return costs

        def samples(step):
            return self.model.get_generator(fixed_z, self.alpha(step), reuse=True).data

        return reference_loop(self, max_iter, step_fn, lambda: dict(g_loss=last['g_loss'], d_loss=last['d_loss']),
                              dev_costs, samples, capture_fn=self.capture, **kw)
