"""TEST INFRASTRUCTURE -- CPU oracle (torch) restating ACGAN/model.py:21-90 and the losses of ACGAN/train.py:89-121
and common/misc.py:310-394 on top of oracle.ops / oracle.resnet_block.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import it.  Parity unpinned by the reference (no upstream tests / golden vectors;
TensorFlow 1.5 not installable)."""
from __future__ import annotations

import torch

from . import ops
from . import resnet_block as rb


class ACGAN:
    def get_generator(self, g, z_var, labels=None, training=True, reuse=False):
        """ACGAN/model.py:27-57"""
        with g.variable_scope("g_net", reuse=reuse):
            z_var_ = z_var.reshape(z_var.shape[0], -1)                                           # :37
            output = ops.Linear(g, z_var_, z_var_.shape[-1], 4 * 4 * 1024, "G.Input")            # :40
            output = output.reshape(-1, 4, 4, 1024)                                              # :41
            for i in (1, 2, 3):                                                                  # :43-48
                output = rb.ResidualBlock(g, output, output.shape[-1], 256, 3, "G.%d" % i, resample="up",
                                          labels=labels, activation_fn="relu")
            output = rb.Normalize(g, "G.OutputN", output)                                        # :49
            output = rb.nonlinearity(output, activation_fn="relu")                               # :50
            output = ops.Conv2D(g, output, output.shape[-1], 3, 3, 1, "G.Output", he_init=False, biases=True)
            return torch.tanh(output)                                                            # :53

    def get_discriminator(self, g, x_var, labels=None, update_collection=None, reuse=False):
        """ACGAN/model.py:59-90"""
        kw = dict(spectral_normed=False, update_collection=update_collection, labels=labels, biases=True,
                  activation_fn="lrelu")
        with g.variable_scope("d_net", reuse=reuse):
            output = rb.OptimizedResBlockDisc1(g, x_var, activation_fn="lrelu")                  # :69
            output = rb.ResidualBlock(g, output, output.shape[-1], 128, 3, "D.DownBlock.2", resample="down", **kw)
            output = rb.ResidualBlock(g, output, output.shape[-1], 128, 3, "D.NoneBlock.3", resample=None, **kw)
            output = rb.ResidualBlock(g, output, output.shape[-1], 128, 3, "D.NoneBlock.4", resample=None, **kw)
            output = rb.nonlinearity(output, activation_fn="lrelu")                              # :80
            output = torch.mean(output, dim=(1, 2))                                              # :81
            output_wgan = ops.Linear(g, output, output.shape[-1], 1, "D.Output", spectral_normed=False,
                                     update_collection=update_collection, biases=True).reshape(-1)   # :84-87
            output_acgan = ops.Linear(g, output, output.shape[-1], 10, "D.ACGANOutput", spectral_normed=False,
                                      update_collection=update_collection, biases=True)          # :90-93
            return output_wgan, output_acgan


def get_loss(disc_real, disc_fake, loss_type="HINGE"):
    """common/misc.py:310-394 -> (d_loss, g_loss)"""
    sp = torch.nn.functional.softplus
    if loss_type == "HINGE":
        d_loss = torch.relu(1.0 - disc_real).mean() + torch.relu(1.0 + disc_fake).mean()
        g_loss = -disc_fake.mean()
    elif loss_type in ("WGAN", "WGAN-GP"):
        d_loss = -disc_real.mean() + disc_fake.mean()
        g_loss = -disc_fake.mean()
    elif loss_type == "LSGAN":
        d_loss = (torch.square(1.0 - disc_real).mean() + torch.square(disc_fake).mean()) / 2.0
        g_loss = torch.square(1.0 - disc_fake).mean() / 2.0
    elif loss_type == "CGAN":   # sigmoid cross entropy with labels 1 / 0
        d_loss = sp(-disc_real).mean() + sp(disc_fake).mean()
        g_loss = sp(-disc_fake).mean()
    elif loss_type == "Modified_MiniMax":
        d_loss = -torch.log(torch.sigmoid(disc_real)).mean() - torch.log(1.0 - torch.sigmoid(disc_fake)).mean()
        g_loss = -torch.log(torch.sigmoid(disc_fake)).mean()
    elif loss_type == "MiniMax":
        d_loss = -torch.log(torch.sigmoid(disc_real)).mean() - torch.log(1.0 - torch.sigmoid(disc_fake)).mean()
        g_loss = torch.log(1.0 - torch.sigmoid(disc_fake)).mean()
    else:
        raise ValueError(loss_type)
    return d_loss, g_loss


def sparse_softmax_xent_mean(logits, labels):
    """tf.reduce_mean(tf.nn.sparse_softmax_cross_entropy_with_logits(...)) (ACGAN/train.py:110-121)"""
    return torch.nn.functional.cross_entropy(logits, labels.long(), reduction="mean")


def gradient_penalty(g, model, real, fake, alpha, labels):
    """ACGAN/train.py:97-104 (torch double backward stands in for tf.gradients of tf.gradients)"""
    interpolates = (real + alpha.reshape(-1, 1, 1, 1) * (fake - real)).detach().requires_grad_(True)
    d = model.get_discriminator(g, interpolates, labels, update_collection=ops.NO_OPS, reuse=True)[0]
    gradients = torch.autograd.grad(d.sum(), interpolates, create_graph=True)[0]
    slopes = torch.sqrt(torch.sum(gradients ** 2, dim=(1, 2, 3)) + 1e-10)
    return 10 * torch.mean((slopes - 1.0) ** 2)
