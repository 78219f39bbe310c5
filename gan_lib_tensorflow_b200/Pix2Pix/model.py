"""Pix2Pix/model.py:9-100: the generator / discriminator factory behind the reference's --net_type flag.

net_type 'UNet' is the nine-level 512x512 pair unet_generator / unet_discriminator, 'UNet_Attention' the eight-level
256x256 pair unet_g / unet_d (the topology BASELINE.json's config 4 names; its Self_Attn calls are commented out or
unrunnable in the reference).  'ResNet' and 'VGG' cannot run in the reference either (resnet_generator passes
resample='None' to ResidualBlock, resnet_discriminator is `pass`; the VGG pair needs /home/yhx/vgg19.npy) and raise."""
from __future__ import annotations

from ..framework import get_store
from . import networks


class Pix2Pix(object):
    def __init__(self):
        pass

    def get_generator(self, inputs, outputs_channels, ngf=64, conv_type='conv2d', channel_multiplier=None,
                      padding='SAME', net_type='UNet', reuse=False, upsampe_method='depth_to_space', keep_masks=None):
        """g-net (:15-59).  keep_masks: the three dropout keep masks (TF's RNG is not reproducible; None = no dropout)."""
        with get_store().variable_scope('g_net', reuse=reuse):
            kw = dict(conv_type=conv_type, channel_multiplier=channel_multiplier or 0, padding=padding,
                      upsampe_method=upsampe_method, keep_masks=keep_masks)
            if net_type == 'UNet':
                return networks.unet_generator(inputs, outputs_channels, ngf, **kw)
            if net_type == 'UNet_Attention':
                return networks.unet_g(inputs, outputs_channels, ngf, **kw)
            if net_type in ('ResNet', 'VGG'):
                raise NotImplementedError('Generator model [%s] cannot run in the reference either (DESIGN.md 7)' % net_type)
            raise NotImplementedError('Generator model name [%s] is not recognized' % net_type)

    def get_discriminator(self, inputs, targets, ndf=64, spectral_normed=True, update_collection=None,
                          conv_type='conv2d', channel_multiplier=None, padding='VALID', net_type='UNet',
                          reuse=False):
        """d-net (:61-100)."""
        with get_store().variable_scope('d_net', reuse=reuse):
            kw = dict(conv_type=conv_type, channel_multiplier=channel_multiplier or 0, padding=padding)
            if net_type == 'UNet':
                return networks.unet_discriminator(inputs, targets, ndf, spectral_normed, update_collection, **kw)
            if net_type == 'UNet_Attention':
                return networks.unet_d(inputs, targets, ndf, spectral_normed, update_collection, **kw)
            if net_type in ('ResNet', 'VGG'):
                raise NotImplementedError(
                    'Discriminator model [%s] cannot run in the reference either (DESIGN.md 7)' % net_type)
            raise NotImplementedError('Discriminator model name [%s] is not recognized' % net_type)
