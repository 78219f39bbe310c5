"""ACGAN CIFAR-10 ResNet (config 2 of BASELINE.json) on the B200 layer ops, with the reference's class, method names,
arguments and variable scopes (ACGAN/model.py:21-90) plus the losses of ACGAN/train.py:89-121.

G = Linear -> three 'up' residual blocks with class-conditional batch norm (library Normalize dispatch,
common/resnet_block.py:32-50) -> batch norm -> relu -> 3x3 conv -> tanh.  D = OptimizedResBlockDisc1 + one 'down' and
two plain residual blocks, NOT spectrally normalised, so the library dispatch gives every D block plain batch norm;
leaky-ReLU throughout; spatial mean; two Linear heads (critic logit, 10-way auxiliary classifier).

The gradient penalty of ACGAN/train.py:97-105 lives in ACGAN/gp.py (it needs the interpolated images, not only the
logits); `discriminator_losses(gradient_penalty=True)` points there."""
from __future__ import annotations

import torch

from .. import functional as F
from ..common import resnet_block as rb
from ..common.ops import conv2d as conv2d_ops
from ..common.ops import linear as linear_ops
from ..framework import get_store

BF16 = torch.bfloat16


class ACGAN(object):
    G_INPUT_DIM = 1024    # channels of the 4x4 seed (ACGAN/model.py:34-35)
    G_DIM = 256           # channels of the three 'up' blocks (:37-42); ACGAN/model_.py uses 128 / 128

    def __init__(self):
        pass

    def get_generator(self, z_var, labels=None, training=True, reuse=False):
        """ACGAN/model.py:27-57.  Returns Var [n, 32, 32, 3] in (-1, 1)."""
        store = get_store()
        with store.variable_scope('g_net', reuse=reuse):
            z_var_ = F.as_var(z_var)
            z_var_ = F.reshape(z_var_, (z_var_.shape[0], -1))
            output = linear_ops.Linear(z_var_, z_var_.shape[-1], 4 * 4 * self.G_INPUT_DIM, 'G.Input', out_dtype=BF16)
            output = F.reshape(output, (-1, 4, 4, self.G_INPUT_DIM))
            for i in (1, 2, 3):
                output = rb.ResidualBlock(output, output.shape[-1], self.G_DIM, 3, 'G.%d' % i, resample='up', labels=labels,
                                          activation_fn='relu', out_dtype=BF16)
            # Normalize('G.OutputN', output) has no labels -> plain batch norm; fused with the relu behind it
            output, _ = rb._norm_act('G.OutputN', output, None, rb._normalize_kind('G.OutputN', None, False), 'relu')
            output = conv2d_ops.Conv2D(output, output.shape[-1], 3, 3, 1, 'G.Output', he_init=False, biases=True)
            return F.activation(output, 'tanh')

    def get_discriminator(self, x_var, labels=None, update_collection=None, reuse=False):
        """ACGAN/model.py:59-90.  Returns (output_wgan Var [n], output_acgan Var [n, 10])."""
        store = get_store()
        kw = dict(spectral_normed=False, update_collection=update_collection, labels=labels, biases=True,
                  activation_fn='lrelu')
        with store.variable_scope('d_net', reuse=reuse):
            output = rb.OptimizedResBlockDisc1(F.as_var(x_var), activation_fn='lrelu')
            output = rb.ResidualBlock(output, output.shape[-1], 128, 3, 'D.DownBlock.2', resample='down', **kw)
            output = rb.ResidualBlock(output, output.shape[-1], 128, 3, 'D.NoneBlock.3', resample=None, **kw)
            output = rb.ResidualBlock(output, output.shape[-1], 128, 3, 'D.NoneBlock.4', resample=None, **kw)
            output = F.act_mean_hw(output, 'lrelu')
            output_wgan = linear_ops.Linear(output, output.shape[-1], 1, 'D.Output', spectral_normed=False,
                                            update_collection=update_collection, biases=True)
            output_wgan = F.reshape(output_wgan, (-1,))
            output_acgan = linear_ops.Linear(output, output.shape[-1], 10, 'D.ACGANOutput', spectral_normed=False,
                                             update_collection=update_collection, biases=True)
            return output_wgan, output_acgan


def discriminator_losses(disc_real, disc_real_acgan, real_labels, disc_fake, loss_type='HINGE',
                         gradient_penalty=False):
    """d_loss = d_loss_gan (+ gradient penalty) + d_loss_acgan (ACGAN/train.py:94-115).
    Returns (d_loss Var, device scalars {d_loss_gan, d_loss_acgan})."""
    if gradient_penalty:
        raise NotImplementedError("the WGAN-GP term is a function of the images: add ACGAN.gp.gradient_penalty(real, "
                                  "fake, alpha) to this loss (ACGAN.train.Trainer does)")
    logits = F.concat_rows(disc_real, disc_fake)
    n_real = disc_real.shape[0]
    d_loss_gan = F.gan_loss(logits, 'd', n_real=n_real, loss_type=loss_type)
    d_loss_acgan = F.softmax_xent(disc_real_acgan, real_labels)
    return F.add_scalars(d_loss_gan, d_loss_acgan), {'d_loss_gan': d_loss_gan.data, 'd_loss_acgan': d_loss_acgan.data}


def generator_losses(disc_fake, disc_fake_acgan, fake_labels, loss_type='HINGE', acgan_scale_G=0.1):
    """g_loss = g_loss_gan + acgan_scale_G * g_loss_acgan (ACGAN/train.py:117-121)."""
    g_loss_gan = F.gan_loss(disc_fake, 'g', loss_type=loss_type)
    g_loss_acgan = F.softmax_xent(disc_fake_acgan, fake_labels, scale=acgan_scale_G)
    return F.add_scalars(g_loss_gan, g_loss_acgan), {'g_loss_gan': g_loss_gan.data,
                                                     'g_loss_acgan_scaled': g_loss_acgan.data}
