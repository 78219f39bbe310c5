"""TEST INFRASTRUCTURE -- CPU oracle (torch fp32/fp64) restating PGGAN/model_nvidia.py:15-237 (lrelu, minibatch_std,
generator_block, get_generator, discriminator_block, get_discriminator incl. the fade-in skip connections) line by
line on top of oracle.ops.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.
Parity unpinned by the reference (no upstream tests / golden vectors; TensorFlow 1.5 is not installable here)."""
from __future__ import annotations

import torch

from . import ops
from . import resnet_block as rb


def lrelu(x, leakiness=0.2):
    """model_nvidia.py:15-17: tf.maximum(x, 0.2 x) (gradient 1 at x == 0)."""
    assert leakiness <= 1, "leakiness must be <= 1"
    return torch.where(x >= 0, x, leakiness * x)


def minibatch_std(x):
    """model_nvidia.py:20-28"""
    m = torch.mean(x, dim=0, keepdim=True)                               # :22
    v = torch.mean((x - m) * (x - m), dim=0, keepdim=True)              # :24
    std = torch.mean(torch.sqrt(v + 1e-8))                               # :25
    std = std.reshape(1, 1, 1, 1).expand(x.shape[0], x.shape[1], x.shape[2], 1)   # :26
    return torch.cat([x, std], dim=3)                                    # :28


def avg_pool2(x):
    """tf.nn.avg_pool(ksize 2, stride 2, VALID) (:160, :185)"""
    return rb.mean_pool2(x)


class PGGAN:
    def __init__(self, block_count, trans, inputs_norm):
        self.bc = block_count          # :37
        self.trans = trans             # :38
        self.inputs_norm = inputs_norm  # :39

    def get_dim(self, stage):
        """:41-47 (a float in the reference under Python 3; used as a channel count)"""
        return int(min(2048 / (2 ** stage), 512))

    def generator_block(self, g, inputs, out_dim, name="generator_block"):
        """:49-71"""
        with g.variable_scope(name):
            output = rb.upsample2(inputs)                                                        # :58-59
            output = ops.Conv2D(g, output, output.shape[-1], out_dim, 3, 1, "Conv.1",
                                inputs_norm=self.inputs_norm, he_init=True, biases=True)         # :61-62
            output = lrelu(ops.pixel_norm(output))                                               # :63-64
            output = ops.Conv2D(g, output, output.shape[-1], out_dim, 3, 1, "Conv.2",
                                inputs_norm=self.inputs_norm, he_init=True, biases=True)         # :66-67
            output = lrelu(ops.pixel_norm(output))                                               # :68-69
        return output

    def get_generator(self, g, z_var, alpha, training=True, reuse=False):
        """:73-129"""
        with g.variable_scope("g_net", reuse=reuse):
            z_var_ = z_var.reshape(z_var.shape[0], -1)                                           # :83
            output = ops.Linear(g, z_var_, z_var_.shape[-1], 4 * 4 * 512, "G.Input",
                                inputs_norm=self.inputs_norm)                                    # :86-87
            output = output.reshape(-1, 4, 4, 512)                                               # :88
            output = lrelu(ops.pixel_norm(output))                                               # :89-90
            output = ops.Conv2D(g, output, output.shape[-1], 512, 3, 1, "G.Conv",
                                inputs_norm=self.inputs_norm, he_init=True, biases=True)         # :93-94
            output = lrelu(ops.pixel_norm(output))                                               # :95-96
            for i in range(self.bc - 1):                                                         # :99-101
                output = self.generator_block(g, output, self.get_dim(i), "G.UpBlock.{}".format(i + 1))
            if self.trans:                                                                       # :103-118
                toRGB1 = self.generator_block(g, output, self.get_dim(self.bc - 1), "G.UpBlock.{}".format(self.bc))
                toRGB1 = ops.Conv2D(g, toRGB1, toRGB1.shape[-1], 3, 1, 1, "G.{}_toRGB1".format(self.bc),
                                    inputs_norm=self.inputs_norm, he_init=True, biases=True)
                toRGB2 = rb.upsample2(output)
                toRGB2 = ops.Conv2D(g, toRGB2, toRGB2.shape[-1], 3, 1, 1, "G.{}_toRGB2".format(self.bc),
                                    inputs_norm=self.inputs_norm, he_init=True, biases=True)
                toRGB = (1 - alpha) * toRGB2 + alpha * toRGB1                                    # :118
            else:                                                                                # :119-127
                if self.bc > 0:
                    toRGB = self.generator_block(g, output, self.get_dim(self.bc - 1), "G.UpBlock.{}".format(self.bc))
                else:
                    toRGB = output
                toRGB = ops.Conv2D(g, toRGB, toRGB.shape[-1], 3, 1, 1, "G.{}_toRGB".format(self.bc),
                                   inputs_norm=self.inputs_norm, he_init=True, biases=True)
        return toRGB

    def discriminator_block(self, g, inputs, out_dim, name, spectral_normed=False, update_collection=None,
                            reuse=False):
        """:131-162"""
        with g.variable_scope(name):
            output = ops.Conv2D(g, inputs, inputs.shape[-1], inputs.shape[-1], 3, 1, "Conv.1",
                                spectral_normed=spectral_normed, update_collection=update_collection, reuse=reuse,
                                he_init=True, biases=True)                                       # :142-147
            output = lrelu(output)                                                               # :149
            output = ops.Conv2D(g, output, output.shape[-1], out_dim, 3, 1, "Conv.2",
                                spectral_normed=spectral_normed, update_collection=update_collection, reuse=reuse,
                                he_init=True, biases=True)                                       # :151-155
            output = lrelu(output)                                                               # :157
            output = avg_pool2(output)                                                           # :160
        return output

    def get_discriminator(self, g, x_var, alpha, spectral_normed=True, update_collection=None, reuse=False):
        """:164-237"""
        kw = dict(spectral_normed=spectral_normed, update_collection=update_collection, reuse=reuse)
        with g.variable_scope("d_net", reuse=reuse):
            if self.trans:                                                                       # :175-200
                fromRGB1 = ops.Conv2D(g, x_var, x_var.shape[-1], self.get_dim(self.bc - 1), 1, 1,
                                      "D.{}_fromRGB1".format(self.bc), he_init=True, biases=True, **kw)
                fromRGB1 = self.discriminator_block(g, fromRGB1, self.get_dim(self.bc - 1),
                                                    "D.Block.{}".format(self.bc), **kw)
                fromRGB2 = avg_pool2(x_var)
                fromRGB2 = ops.Conv2D(g, fromRGB2, fromRGB2.shape[-1], self.get_dim(self.bc - 1), 1, 1,
                                      "D.{}_fromRGB2".format(self.bc), he_init=True, biases=True, **kw)
                x_code = (1 - alpha) * fromRGB2 + alpha * fromRGB1                               # :200
            else:                                                                                # :201-215
                x_code = ops.Conv2D(g, x_var, x_var.shape[-1], self.get_dim(self.bc - 1), 1, 1,
                                    "D.{}_fromRGB".format(self.bc), he_init=True, biases=True, **kw)
                if self.bc > 0:
                    x_code = self.discriminator_block(g, x_code, self.get_dim(self.bc - 1),
                                                      "D.Block.{}".format(self.bc), **kw)
            for i in range(1, self.bc):                                                          # :217-223
                x_code = self.discriminator_block(g, x_code, self.get_dim(self.bc - 1 - i),
                                                  "D.Block.{}".format(self.bc - i), **kw)
            output = minibatch_std(x_code)                                                       # :225
            output = ops.Conv2D(g, output, output.shape[-1], self.get_dim(self.bc - 1), 3, 1, "D.Conv",
                                he_init=True, biases=True, **kw)                                 # :226-231
            output = lrelu(output)                                                               # :232
            output = torch.mean(output, dim=(1, 2))                                              # :234
            logits = ops.Linear(g, output, output.shape[-1], 1, "D.Output")                      # :235
            return logits.reshape(-1)                                                            # :236


class PGGANResNet:
    """PGGAN/model_resnet.py:14-70: the ResNet variant (common/resnet_block.py:192-349) behind the same two methods."""

    def __init__(self, block_count, trans, inputs_norm):
        self.bc, self.trans, self.inputs_norm = block_count, trans, inputs_norm

    def get_generator(self, g, z_var, alpha, training=True, reuse=False):
        with g.variable_scope("g_net", reuse=reuse):                                                 # :33-38
            z_var_ = z_var.reshape(z_var.shape[0], -1)
            return rb.Generator_PGGAN(g, z_var_, self.bc, self.trans, alpha, self.inputs_norm, training=training)

    def get_discriminator(self, g, x_var, alpha, labels=None, update_collection=None, reuse=False):
        with g.variable_scope("d_net", reuse=reuse):                                                 # :50-69
            return rb.Discriminator_PGGAN(g, x_var, labels, self.bc, self.trans, alpha, self.inputs_norm,
                                          update_collection=update_collection, reuse=reuse)


class PGGANLosses:
    """Losses of PGGAN/train.py:103-113 (hinge) over the model above: returns (cost, params, grads) like the SNGAN
    oracle; D(real) runs with update_collection=None, the fake branch with NO_OPS."""

    def __init__(self, g, block_count, trans, inputs_norm, size, z_dim=512, model="nvidia"):
        cls = PGGAN if model == "nvidia" else PGGANResNet                                     # train.py:61-66
        self.g, self.model = g, cls(block_count, trans, inputs_norm)
        with torch.no_grad():   # graph construction order: D(real), G, D(fake, reuse)
            g.draw_on_reuse = True
            self.model.get_discriminator(g, torch.zeros(2, size, size, 3), 0.0, update_collection=ops.NO_OPS)
            f = self.model.get_generator(g, torch.zeros(2, z_dim), 0.0)
            self.model.get_discriminator(g, f, 0.0, update_collection=ops.NO_OPS, reuse=True)
            g.draw_on_reuse = False

    def d_grads(self, real, z, alpha):
        m, g = self.model, self.g
        with torch.no_grad():
            fake = m.get_generator(g, z, alpha, reuse=True)
        disc_real = m.get_discriminator(g, real, alpha, update_collection=None, reuse=True)
        disc_fake = m.get_discriminator(g, fake, alpha, update_collection=ops.NO_OPS, reuse=True)
        cost = torch.relu(1.0 - disc_real).mean() + torch.relu(1.0 + disc_fake).mean()      # train.py:110-112
        params = g.trainable_variables("d_net")
        return cost.detach(), params, torch.autograd.grad(cost, [p for _, p in params], allow_unused=True)

    def g_grads(self, z, alpha):
        m, g = self.model, self.g
        disc_fake = m.get_discriminator(g, m.get_generator(g, z, alpha, reuse=True), alpha,
                                        update_collection=ops.NO_OPS, reuse=True)
        cost = -disc_fake.mean()                                                              # train.py:113
        params = g.trainable_variables("g_net")
        return cost.detach(), params, torch.autograd.grad(cost, [p for _, p in params], allow_unused=True)
