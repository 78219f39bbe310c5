"""Training steps of Pix2Pix/train.py:447-539, 694-729 on the B200 layer ops (config 4 of BASELINE.json): U-Net
generator, spectrally-normalised PatchGAN, lib.misc.get_loss(loss_type) for both players, gen_loss = gan_weight *
GAN + l1_weight * L1, two Adam(beta1 = 0, beta2 = 0.9) optimisers with a linear learning-rate decay driven by the
generator's global step; per iteration n_dis critic steps on the batch, then one generator step.

Every discriminator call uses update_collection=None in the reference (train.py:459-478): u is re-assigned by D(real),
by D(fake) and again by D(fake) of the generator step, so each call sees a different sigma (framework.sn_acquire keeps
one spectral-norm state per evaluation).  The WGAN-GP penalty (train.py:489-503) is Pix2Pix/gp.py: a third such pass."""
from __future__ import annotations

import numpy as np
import torch

from .. import functional as F
from ..framework import Var, get_store
from ..training import TwoPlayer
from . import networks  # noqa: F401
from .model import Pix2Pix


class Trainer:
    def __init__(self, ngf: int = 64, ndf: int = 64, size: int = 256, loss_type: str = 'HINGE',
                 gan_weight: float = 1.0, l1_weight: float = 100.0, initial_lr: float = 0.0002, end_lr: float = 0.0001,
                 beta1: float = 0.0, beta2: float = 0.9, max_steps: int = 23600, seed: int | None = 0,
                 world_size: int = 1, grad_allreduce=None, net_type: str = 'UNet_Attention', conv_type: str = 'conv2d',
                 channel_multiplier: int = 0):
        """net_type / conv_type / channel_multiplier: the flags of Pix2Pix/train.py:31-36 ('UNet_Attention' = unet_g /
        unet_d, the 256x256 topology of config 4; 'UNet' = the 512x512 pair)."""
        if loss_type == 'WGAN-GP' and conv_type != 'conv2d':
            raise NotImplementedError('the gradient penalty (Pix2Pix/gp.py) is built for conv_type conv2d')
        self.store = get_store()
        self.model = Pix2Pix()
        self.ngf, self.ndf, self.loss_type = ngf, ndf, loss_type
        self.net = dict(net_type=net_type, conv_type=conv_type, channel_multiplier=channel_multiplier)
        self.gan_weight, self.l1_weight = gan_weight, l1_weight
        self.initial_lr, self.end_lr, self.max_steps = initial_lr, end_lr, max_steps
        self.global_step = 0
        if seed is not None:
            np.random.seed(seed)
        dev = self.store.device
        with self.store.building():                      # create_model(): G, D(real), D(fake, reuse)
            x0 = torch.zeros(1, size, size, 3, device=dev)
            out0 = self.model.get_generator(x0, 3, ngf=ngf, **self.net)
            self.model.get_discriminator(x0, x0, ndf=ndf, update_collection="NO_OPS", **self.net)
            self.model.get_discriminator(x0, out0, ndf=ndf, update_collection="NO_OPS", reuse=True, **self.net)
        self.players = TwoPlayer("d_net", "g_net", beta1=beta1, beta2=beta2, world_size=world_size,
                                 grad_allreduce=grad_allreduce)
        self.players.finalize()

    def learning_rate(self) -> float:
        """tf.train.polynomial_decay(initial_lr, global_step, max_steps, end_lr), power 1 (train.py:521-526)."""
        t = min(self.global_step, self.max_steps) / float(self.max_steps)
        return (self.initial_lr - self.end_lr) * (1.0 - t) + self.end_lr

    # ------------------------------------------------------------------------------------------ losses
    def d_loss(self, inputs, targets, keep_masks=None, gp_alpha=None):
        """discrim_loss of train.py:459-503.  gp_alpha: the tf.random_uniform([batch]) draw of the WGAN-GP term (drawn
        here when None)."""
        m = self.model
        outputs = m.get_generator(inputs, 3, ngf=self.ngf, reuse=True, keep_masks=keep_masks, **self.net)  # g_net frozen
        predict_real = m.get_discriminator(inputs, targets, ndf=self.ndf, update_collection=None, reuse=True, **self.net)
        predict_fake = m.get_discriminator(inputs, Var(outputs.data), ndf=self.ndf, update_collection=None, reuse=True,
                                           **self.net)
        pr, pf = F.reshape(predict_real, (-1,)), F.reshape(predict_fake, (-1,))
        loss = F.gan_loss(F.concat_rows(pr, pf), 'd', n_real=pr.shape[0], loss_type=self.loss_type)
        if self.loss_type == 'WGAN-GP':                    # train.py:489-503, a third update_collection=None pass of D
            from . import gp
            x = inputs.data if isinstance(inputs, Var) else inputs
            t = targets.data if isinstance(targets, Var) else targets
            if gp_alpha is None:
                gp_alpha = torch.rand(x.shape[0], device=x.device)
            penalty = gp.gradient_penalty(x, t, outputs.data, gp_alpha,
                                          n_layers=4 if self.net['net_type'] == 'UNet' else 3)
            self.last_penalty = penalty.data
            loss = F.add_scalars(loss, penalty)
        return loss

    def g_loss(self, inputs, targets, keep_masks=None):
        m = self.model
        outputs = m.get_generator(inputs, 3, ngf=self.ngf, reuse=True, keep_masks=keep_masks, **self.net)
        predict_fake = m.get_discriminator(inputs, outputs, ndf=self.ndf, update_collection=None, reuse=True, **self.net)
        gen_loss_gan = F.gan_loss(F.reshape(predict_fake, (-1,)), 'g', scale=self.gan_weight, loss_type=self.loss_type)
        gen_loss_l1 = F.l1_loss(targets, outputs, scale=self.l1_weight)
        self.last = {'gen_loss_GAN_weighted': gen_loss_gan.data, 'gen_loss_L1_weighted': gen_loss_l1.data}
        return F.add_scalars(gen_loss_gan, gen_loss_l1)

    # ------------------------------------------------------------------------------------------ steps
    def _stage(self, inputs, targets, keep_masks):
        s = self.static
        s['inputs'].copy_(inputs, non_blocking=True)
        s['targets'].copy_(targets, non_blocking=True)
        if (keep_masks is None) != (s['masks'] is None):
            raise ValueError("captured with%s dropout masks: pass keep_masks accordingly" % ("out" if s['masks'] is None else ""))
        if keep_masks is not None:
            for dst, src in zip(s['masks'], keep_masks):
                dst.copy_(src, non_blocking=True)

    def d_step(self, inputs, targets, keep_masks=None, gp_alpha=None):
        if self.players.captured("d"):
            self._stage(inputs, targets, keep_masks)
            if self.loss_type == 'WGAN-GP':
                if gp_alpha is None:
                    self.static['alpha'].uniform_()
                else:
                    self.static['alpha'].copy_(gp_alpha, non_blocking=True)
            return self.players.replay("d", self.learning_rate())
        return self.players.step("d", lambda: self.d_loss(inputs, targets, keep_masks, gp_alpha), self.learning_rate())

    def g_step(self, inputs, targets, keep_masks=None):
        if self.players.captured("g"):
            self._stage(inputs, targets, keep_masks)
            loss = self.players.replay("g", self.learning_rate())
        else:
            loss = self.players.step("g", lambda: self.g_loss(inputs, targets, keep_masks), self.learning_rate())
        self.global_step += 1                             # gen_optim.apply_gradients(..., global_step=global_step)
        return loss

    def capture(self, inputs, targets, keep_masks=None):
        """Both training ops as CUDA graphs over static buffers shaped like the given batch (call after one eager
        d_step and g_step with the same shapes); later d_step / g_step calls copy their arguments in and replay."""
        self.static = {'inputs': torch.empty_like(inputs), 'targets': torch.empty_like(targets),
                       'masks': None if keep_masks is None else [torch.empty_like(m) for m in keep_masks],
                       'alpha': torch.rand(inputs.shape[0], device=inputs.device)}
        s = self.static
        self.players.capture("d", lambda: self.d_loss(s['inputs'], s['targets'], s['masks'], s['alpha']))
        self.players.capture("g", lambda: self.g_loss(s['inputs'], s['targets'], s['masks']))

    def train_iteration(self, inputs, targets, n_dis: int = 5, mask_fn=None):
        """train.py:703-729: n_dis critic steps on the batch, then the generator step.  mask_fn() -> three dropout
        keep masks (fresh ones for every session.run in the reference)."""
        d = None
        for _ in range(n_dis):
            d = self.d_step(inputs, targets, mask_fn() if mask_fn else None)
        g = self.g_step(inputs, targets, mask_fn() if mask_fn else None)
        return d, g

    # ------------------------------------------------------------------------------------------ the script's loop
    def train(self, batches, max_steps: int | None = None, n_dis: int = 5, progress_freq: int = 50,
              display_freq: int = 0, save_freq: int = 4000, val_batches=(), out_dir: str = '.', mask_fn=None,
              capture_after: int | None = 0, log=print):
        """Pix2Pix/train.py:694-772.  `batches`: a sequence of (inputs, targets) device tensor pairs in [-1, 1]
        (train_data[idx] after the script's preprocessing; indexed step % len like :695); per step n_dis critic
        steps on the batch, then the generator step; every progress_freq steps the three losses (:746-749, GAN and L1
        unweighted like the `gen_loss_GAN` / `gen_loss_L1` fetches); every display_freq steps the
        inputs | targets | outputs grid of the batch; every save_freq steps the validation pass (:751-766: the script
        saves images there, not a model) over val_batches.  should(freq) also fires on the last step (:701-702)."""
        import os
        import time

        from ..common import misc as lib_misc

        max_steps = self.max_steps if max_steps is None else max_steps
        start = time.time()
        captured = False

        def should(freq, step):
            return freq > 0 and ((step + 1) % freq == 0 or step == max_steps - 1)

        def triple(inputs, targets):
            out = self.model.get_generator(inputs, 3, ngf=self.ngf, reuse=True, keep_masks=None, **self.net).data
            return torch.cat([inputs, targets, out.to(inputs.dtype)], dim=0)       # plumbing: the display fetches

        for step in range(max_steps):
            inputs, targets = batches[step % len(batches)]
            masks = (lambda: mask_fn()) if mask_fn else (lambda: None)
            d = None
            for _ in range(n_dis):
                d = self.d_step(inputs, targets, masks())
            self.g_step(inputs, targets, masks())
            if capture_after is not None and not captured and step >= capture_after:
                self.capture(inputs, targets, masks())
                captured = True
            if should(progress_freq, step):
                rate = (step + 1) * inputs.shape[0] / (time.time() - start)
                log("progress  step %d  image/sec %0.1f" % (self.global_step, rate))
                log("discrim_loss", float(d.reshape(-1)[0] if torch.is_tensor(d) else d.data.reshape(-1)[0]))
                log("gen_loss_GAN", float(self.last['gen_loss_GAN_weighted'].reshape(-1)[0]) / self.gan_weight)
                log("gen_loss_L1", float(self.last['gen_loss_L1_weighted'].reshape(-1)[0]) / self.l1_weight)
            if should(display_freq, step):
                lib_misc.save_images(triple(inputs, targets), os.path.join(out_dir, 'train_%08d.png' % self.global_step))
            if should(save_freq, step):
                for i, (vi, vt) in enumerate(val_batches):
                    lib_misc.save_images(triple(vi, vt), os.path.join(out_dir, 'val_%04d.png' % i))
                    log("evaluated image", 'val_%04d.png' % i)
        return self
