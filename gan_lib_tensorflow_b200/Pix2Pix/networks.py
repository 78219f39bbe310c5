"""Pix2Pix U-Net generator and PatchGAN discriminator on the B200 layer ops, with the reference's function names,
arguments and variable scopes (Pix2Pix/networks.py:25-43, 174-354; the 512x512 pair unet_generator /
unet_discriminator of :359-536 is the same graph one level deeper).

Graph kept from the reference:
  * unet_g: encoder_1 = 4x4 s2 SAME conv; encoder_2..8 = lrelu -> 4x4 s2 conv -> instance norm; decoder_8..2 =
    concat(skip) -> relu -> nearest 2x -> 4x4 s1 SAME conv (TF pads 1 before / 2 after) -> instance norm
    [-> dropout 0.5 on the first three]; decoder_1 = concat -> relu -> nearest 2x -> conv -> tanh;
  * unet_d: concat(inputs, targets) -> [tf.pad 1 -> 4x4 VALID conv (s2, s2, s2, s1) -> lrelu] x4 -> pad -> conv to 1.
The Self_Attn calls of the reference cannot run (Appendix B of SURVEY.md: `x.shape.as_list[-1]`, `out.view`) and are
left out, as in the `unet_g` / `unet_d` topology named for config 4.

Scheduling: the activation in front of every convolution (+ the nearest upsample of the decoders) is one
bandwidth-bound kernel that emits the bf16 tensor-core operand; dropout masks are inputs (TF's RNG is not
reproducible)."""
from __future__ import annotations

import torch

from .. import functional as F
from ..common.ops import conv2d as conv2d_ops
from ..common.ops import normalization as norm_ops
from ..framework import get_store

BF16 = torch.bfloat16


def norm_layer(inputs, decay=0.9, epsilon=1e-5, is_training=True, norm_type="BN"):
    """Pix2Pix/networks.py:25-43"""
    if norm_type == "BN":
        return norm_ops.batch_norm(inputs, decay=decay, epsilon=epsilon, is_training=True)
    if norm_type == "IN":
        return norm_ops.instance_norm(inputs, epsilon=epsilon)
    raise NotImplementedError('Normalization [%s] is not implemented!' % norm_type)


def _act(x, act, upsample=False):
    out, _ = F.norm_act(F.as_var(x), stats=None, act=act, upsample=upsample, out_dtype=BF16)
    return out


def _conv(inputs, out_channels, stride, padding, spectral_normed=False, update_collection=None, conv_type='conv2d',
          channel_multiplier=0):
    return conv2d_ops.Conv2D(inputs, inputs.shape[-1], out_channels, 4, stride, 'Conv2D', conv_type=conv_type,
                             channel_multiplier=channel_multiplier, padding=padding, spectral_normed=spectral_normed,
                             update_collection=update_collection, inputs_norm=False, he_init=True, biases=True)


def unet_g(generator_inputs, generator_outputs_channels, ngf, conv_type='conv2d', channel_multiplier=0, padding='SAME',
           upsampe_method='depth_to_space', keep_masks=None):
    """Pix2Pix/networks.py:174-284 (8 encoders: 256x256 -> 1x1).  keep_masks: three fp32 {0,1} tensors for the dropout
    of decoder_8/7/6 (None disables dropout, i.e. keep_prob = 1)."""
    return _unet(generator_inputs, generator_outputs_channels, ngf, conv_type, channel_multiplier, padding,
                 upsampe_method, keep_masks, deep=4)


def unet_generator(generator_inputs, generator_outputs_channels, ngf, conv_type='conv2d', channel_multiplier=0,
                   padding='SAME', upsampe_method='depth_to_space', keep_masks=None):
    """Pix2Pix/networks.py:359-472: the 512x512 U-Net -- one more ngf*8 encoder (encoder_9: 2x2 -> 1x1) and decoder
    (decoder_9 .. decoder_7 carry the dropout) than unet_g, otherwise the same graph."""
    return _unet(generator_inputs, generator_outputs_channels, ngf, conv_type, channel_multiplier, padding,
                 upsampe_method, keep_masks, deep=5)


def _unet(generator_inputs, generator_outputs_channels, ngf, conv_type, channel_multiplier, padding, upsampe_method,
          keep_masks, deep):
    """`deep` = number of ngf*8 encoders after encoder_4 (4: unet_g, 5: unet_generator)."""
    if upsampe_method not in ('depth_to_space', 'resize'):
        raise NotImplementedError('upsampe_method [%s] is not recognized' % upsampe_method)  # both are nearest 2x
    store = get_store()
    layers = []
    with store.variable_scope("encoder_1"):
        layers.append(_conv(F.as_var(generator_inputs), ngf, 2, padding, conv_type=conv_type,
                            channel_multiplier=channel_multiplier))
    for out_channels in (ngf * 2, ngf * 4, ngf * 8) + (ngf * 8,) * deep:
        with store.variable_scope("encoder_%d" % (len(layers) + 1)):
            rectified = _act(layers[-1], 'lrelu')
            convolved = _conv(rectified, out_channels, 2, padding, conv_type=conv_type,
                              channel_multiplier=channel_multiplier)
            layers.append(norm_layer(convolved, decay=0.9, epsilon=1e-5, is_training=True, norm_type="IN"))
    layer_specs = [(ngf * 8, 0.5)] * 3 + [(ngf * 8, 0.0)] * (deep - 3) + [(ngf * 4, 0.0), (ngf * 2, 0.0), (ngf, 0.0)]
    num_encoder_layers = len(layers)
    for decoder_layer, (out_channels, dropout) in enumerate(layer_specs):
        skip_layer = num_encoder_layers - decoder_layer - 1
        with store.variable_scope("decoder_%d" % (skip_layer + 1)):
            inputs = layers[-1] if decoder_layer == 0 else F.concat_channels(layers[-1], layers[skip_layer])
            resized = _act(inputs, 'relu', upsample=True)          # relu + nearest 2x + bf16 cast in one pass
            output = _conv(resized, out_channels, 1, padding, conv_type=conv_type,
                           channel_multiplier=channel_multiplier)
            output = norm_layer(output, decay=0.9, epsilon=1e-5, is_training=True, norm_type="IN")
            if dropout > 0.0 and keep_masks is not None:
                output = F.dropout(output, keep_masks[decoder_layer], 1.0 - dropout)
            layers.append(output)
    with store.variable_scope("decoder_1"):
        inputs = F.concat_channels(layers[-1], layers[0])
        resized = _act(inputs, 'relu', upsample=True)
        output = _conv(resized, generator_outputs_channels, 1, padding, conv_type=conv_type,
                       channel_multiplier=channel_multiplier)
        layers.append(F.activation(output, 'tanh'))
    unet_g.last_layers = layers   # kept for layer-by-layer parity probes
    return layers[-1]


def unet_d(discrim_inputs, discrim_targets, ndf, spectral_normed, update_collection, conv_type='conv2d',
           channel_multiplier=0, padding='VALID'):
    """Pix2Pix/networks.py:287-354: 70x70 PatchGAN; every convolution is tf.pad(1) + 4x4 `padding` conv."""
    return _patchgan(discrim_inputs, discrim_targets, ndf, spectral_normed, update_collection, conv_type,
                     channel_multiplier, padding, n_layers=3)


def unet_discriminator(discrim_inputs, discrim_targets, ndf, spectral_normed, update_collection, conv_type='conv2d',
                       channel_multiplier=0, padding='VALID'):
    """Pix2Pix/networks.py:475-536: the 512x512 PatchGAN -- n_layers = 4 (one more stride-2 layer, ndf*8 twice)."""
    return _patchgan(discrim_inputs, discrim_targets, ndf, spectral_normed, update_collection, conv_type,
                     channel_multiplier, padding, n_layers=4)


def _patchgan(discrim_inputs, discrim_targets, ndf, spectral_normed, update_collection, conv_type, channel_multiplier,
              padding, n_layers):
    if padding != 'VALID':
        raise NotImplementedError("the PatchGAN is built for padding='VALID' (Pix2Pix/train.py:466, 476, 499)")
    store = get_store()
    a, b = F.as_var(discrim_inputs), F.as_var(discrim_targets)
    inputs = F.concat_channels(a, b)
    pad1 = (1, 1, 1, 1)   # tf.pad [[0,0],[1,1],[1,1],[0,0]] folded into the convolution's zero fill
    with store.variable_scope("layer_1"):
        convolved = _conv(inputs, ndf, 2, pad1, spectral_normed, update_collection, conv_type, channel_multiplier)
        layers = [_act(convolved, 'lrelu')]
    for i in range(n_layers):
        with store.variable_scope("layer_%d" % (len(layers) + 1)):
            out_channels_ = ndf * min(2 ** (i + 1), 8)
            stride = 1 if i == n_layers - 1 else 2
            convolved = _conv(layers[-1], out_channels_, stride, pad1, spectral_normed, update_collection, conv_type,
                              channel_multiplier)
            layers.append(_act(convolved, 'lrelu'))
    with store.variable_scope("layer_%d" % (len(layers) + 1)):
        layers.append(_conv(layers[-1], 1, 1, pad1, spectral_normed, update_collection, conv_type,
                            channel_multiplier))
    return layers[-1]
