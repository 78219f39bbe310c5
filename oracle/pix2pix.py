"""TEST INFRASTRUCTURE -- CPU oracle (torch fp32/fp64) restating Pix2Pix/networks.py:25-43, 174-354 (unet_g, unet_d,
norm_layer) line by line on top of oracle.ops.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
may import it.  Parity unpinned by the reference (no tests / golden vectors exist upstream, TensorFlow 1.5 is not
installable here); dropout keep-masks are inputs because TF's op RNG cannot be restated.  The Self_Attn calls of the
reference are unrunnable and are omitted, as in the product."""
from __future__ import annotations

import torch

from . import ops
from . import resnet_block as rb


def norm_layer(g, inputs, decay=0.9, epsilon=1e-5, is_training=True, norm_type="BN"):
    """Pix2Pix/networks.py:25-43"""
    if norm_type == "BN":
        return ops.batch_norm(g, inputs, decay=decay, epsilon=epsilon, is_training=True)
    if norm_type == "IN":
        return ops.instance_norm(g, inputs, epsilon=epsilon)
    raise NotImplementedError("Normalization [%s] is not implemented!" % norm_type)


_CONV_TYPE = ["conv2d", 0]   # (--conv_type, --channel_multiplier) of Pix2Pix/train.py:31-33, set per network call


def _conv(g, x, out_channels, stride, padding, sn=False, uc=None):
    return ops.Conv2D(g, x, x.shape[-1], out_channels, 4, stride, "Conv2D", conv_type=_CONV_TYPE[0],
                      channel_multiplier=_CONV_TYPE[1], padding=padding, spectral_normed=sn,
                      update_collection=uc, inputs_norm=False, he_init=True, biases=True)


def unet_generator(g, generator_inputs, generator_outputs_channels, ngf, padding="SAME", keep_masks=None):
    """Pix2Pix/networks.py:359-472: unet_g with a ninth encoder / decoder level (512x512 -> 1x1)."""
    return unet_g(g, generator_inputs, generator_outputs_channels, ngf, padding, keep_masks, deep=5)


def unet_discriminator(g, discrim_inputs, discrim_targets, ndf, spectral_normed, update_collection, padding="VALID"):
    """Pix2Pix/networks.py:475-536: unet_d with n_layers = 4."""
    return unet_d(g, discrim_inputs, discrim_targets, ndf, spectral_normed, update_collection, padding, n_layers=4)


def unet_g(g, generator_inputs, generator_outputs_channels, ngf, padding="SAME", keep_masks=None, deep=4,
           conv_type="conv2d", channel_multiplier=0):
    """Pix2Pix/networks.py:174-284 (deep = 4); deep = 5 gives the layer lists of unet_generator, :373-383 / :406-415"""
    _CONV_TYPE[:] = [conv_type, channel_multiplier]
    layers = []
    with g.variable_scope("encoder_1"):
        layers.append(_conv(g, generator_inputs, ngf, 2, padding))                               # :178-185
    for out_channels in (ngf * 2, ngf * 4, ngf * 8) + (ngf * 8,) * deep:                         # :187-195
        with g.variable_scope("encoder_%d" % (len(layers) + 1)):
            rectified = rb.nonlinearity(layers[-1], "lrelu", 0.2)                                 # :199
            convolved = _conv(g, rectified, out_channels, 2, padding)                            # :201-206
            layers.append(norm_layer(g, convolved, epsilon=1e-5, norm_type="IN"))                # :207
    layer_specs = ([(ngf * 8, 0.5)] * 3 + [(ngf * 8, 0.0)] * (deep - 3)
                   + [(ngf * 4, 0.0), (ngf * 2, 0.0), (ngf, 0.0)])                               # :214-222
    num_encoder_layers = len(layers)
    for decoder_layer, (out_channels, dropout) in enumerate(layer_specs):
        skip_layer = num_encoder_layers - decoder_layer - 1
        with g.variable_scope("decoder_%d" % (skip_layer + 1)):
            if decoder_layer == 0:
                inputs = layers[-1]
            else:
                inputs = torch.cat([layers[-1], layers[skip_layer]], dim=3)                      # :233
            rectified = torch.relu(inputs)                                                       # :235
            resized = rb.upsample2(rectified)                                                    # :241-243 (nearest 2x)
            output = _conv(g, resized, out_channels, 1, padding)                                 # :247-251
            output = norm_layer(g, output, epsilon=1e-5, norm_type="IN")                         # :256
            if dropout > 0.0 and keep_masks is not None:                                         # :263-264
                output = output * keep_masks[decoder_layer] / (1.0 - dropout)
            layers.append(output)
    with g.variable_scope("decoder_1"):                                                          # :268-282
        inputs = torch.cat([layers[-1], layers[0]], dim=3)
        resized = rb.upsample2(torch.relu(inputs))
        output = torch.tanh(_conv(g, resized, generator_outputs_channels, 1, padding))
        layers.append(output)
    unet_g.last_layers = layers   # kept for layer-by-layer parity probes
    return layers[-1]


def unet_d(g, discrim_inputs, discrim_targets, ndf, spectral_normed, update_collection, padding="VALID", n_layers=3,
           conv_type="conv2d", channel_multiplier=0):
    """Pix2Pix/networks.py:287-354 (n_layers = 3)"""
    _CONV_TYPE[:] = [conv_type, channel_multiplier]
    pad = lambda t: torch.nn.functional.pad(t, (0, 0, 1, 1, 1, 1))  # noqa: E731  tf.pad [[0,0],[1,1],[1,1],[0,0]]
    inputs = torch.cat([discrim_inputs, discrim_targets], dim=3)                                  # :293
    with g.variable_scope("layer_1"):                                                            # :296-307
        convolved = _conv(g, pad(inputs), ndf, 2, padding, spectral_normed, update_collection)
        layers = [rb.nonlinearity(convolved, "lrelu", 0.2)]
    for i in range(n_layers):                                                                    # :312-337
        with g.variable_scope("layer_%d" % (len(layers) + 1)):
            out_channels_ = ndf * min(2 ** (i + 1), 8)
            stride = 1 if i == n_layers - 1 else 2
            convolved = _conv(g, pad(layers[-1]), out_channels_, stride, padding, spectral_normed, update_collection)
            layers.append(rb.nonlinearity(convolved, "lrelu", 0.2))
    with g.variable_scope("layer_%d" % (len(layers) + 1)):                                       # :340-352
        layers.append(_conv(g, pad(layers[-1]), 1, 1, padding, spectral_normed, update_collection))
    return layers[-1]


class Pix2PixLosses:
    """create_model() of Pix2Pix/train.py:447-539 for the unet_g / unet_d topology: returns (cost, params, grads)."""

    def __init__(self, g, ngf, ndf, size, loss_type="HINGE", gan_weight=1.0, l1_weight=100.0):
        self.g, self.ngf, self.ndf, self.loss_type = g, ngf, ndf, loss_type
        self.gan_weight, self.l1_weight = gan_weight, l1_weight
        with torch.no_grad():
            g.draw_on_reuse = True
            x0 = torch.zeros(1, size, size, 3)
            out0 = self.G(x0)
            self.D(x0, x0, ops.NO_OPS)
            self.D(x0, out0, ops.NO_OPS)
            g.draw_on_reuse = False

    def G(self, inputs, keep_masks=None):
        with self.g.variable_scope("g_net"):
            return unet_g(self.g, inputs, 3, self.ngf, keep_masks=keep_masks)

    def D(self, inputs, targets, update_collection):
        with self.g.variable_scope("d_net"):
            return unet_d(self.g, inputs, targets, self.ndf, True, update_collection)

    def d_grads(self, inputs, targets, keep_masks=None, gp_alpha=None):
        from .acgan import get_loss
        with torch.no_grad():
            outputs = self.G(inputs, keep_masks)
        predict_real = self.D(inputs, targets, None)           # train.py:459-467, update_collection=None
        predict_fake = self.D(inputs, outputs, None)           # :470-478
        cost, _ = get_loss(predict_real, predict_fake, self.loss_type)                           # :486
        if self.loss_type == "WGAN-GP":                                                          # :489-503
            alpha = gp_alpha.reshape(-1, 1, 1, 1)              # tf.random_uniform([batch_size, 1, 1, 1])
            differences = outputs - targets
            interpolates = (targets + alpha * differences).detach().requires_grad_(True)
            d_hat = self.D(inputs, interpolates, None)
            gradients = torch.autograd.grad(d_hat.sum(), interpolates, create_graph=True)[0]
            slopes = torch.sqrt(torch.sum(gradients ** 2, dim=(1, 2, 3)) + 1e-10)
            self.last_penalty = 10 * torch.mean((slopes - 1.0) ** 2)
            cost = cost + self.last_penalty
        params = self.g.trainable_variables("d_net")
        return cost.detach(), params, torch.autograd.grad(cost, [p for _, p in params], allow_unused=True)

    def g_grads(self, inputs, targets, keep_masks=None):
        from .acgan import get_loss
        outputs = self.G(inputs, keep_masks)
        predict_fake = self.D(inputs, outputs, None)
        _, gen_loss_gan = get_loss(predict_fake, predict_fake, self.loss_type)                   # :510 (real unused)
        gen_loss_l1 = torch.mean(torch.abs(targets - outputs))                                   # :511
        cost = gen_loss_gan * self.gan_weight + gen_loss_l1 * self.l1_weight                     # :512
        params = self.g.trainable_variables("g_net")
        return cost.detach(), params, torch.autograd.grad(cost, [p for _, p in params], allow_unused=True)
