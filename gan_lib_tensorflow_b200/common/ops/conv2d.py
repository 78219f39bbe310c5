"""Convolution for data in format of 'NHWC' with the reference's signature (common/ops/conv2d.py:31-218).

Initial values come from the NumPy global RNG exactly as in the reference (uniform +-stdev*sqrt(3), He or
Glorot stdev, conv2d.py:83-140), and are drawn on every call during graph construction.
"""
from __future__ import annotations

import numpy as np

from ... import functional as F
from ...framework import Var, get_store
from .sn import spectral_normed_weight

_default_weightnorm = False


def enable_default_weightnorm():
    global _default_weightnorm
    _default_weightnorm = True


_weights_stdev = None


def set_weights_stdev(weights_stdev):
    global _weights_stdev
    _weights_stdev = weights_stdev


def unset_weights_stdev():
    global _weights_stdev
    _weights_stdev = None


def _memo(fn):
    box = []

    def get():
        if not box:
            box.append(fn())
        return box[0]

    return get


def Conv2D(inputs, input_dim, output_dim, filter_size=3, stride=1, name='Conv2D',
           conv_type='conv2d', channel_multiplier=0, padding='SAME',
           spectral_normed=False, update_collection=None, inputs_norm=False, he_init=True,
           mask_type=None, weightnorm=None, biases=True, gain=1., reuse=None,
           residual=None, out_grad_dtype=None, residual_up2=False, out_dtype=None, subpixel_up2=False):
    """
    Args mirror common/ops/conv2d.py:31-55 (`reuse` is the extra keyword of conv2d_.py:33, accepted and ignored).
    `residual` (fp32 Var added in the GEMM epilogue; `residual_up2`: given at half resolution) and `out_grad_dtype`
    are extensions used by resnet_block.  `subpixel_up2`: the layer is UpsampleConv (nearest 2x in front of this 3x3
    convolution, common/resnet_block.py:83-97) and `inputs` is the LOW-resolution tensor: evaluated in sub-pixel
    form, output in quad layout (functional.upconv2d).

    Returns:
      Var of shape (batch_size, out_height, out_width, output_dim), fp32
    """
    store = get_store()
    inputs = F.as_var(inputs)
    with store.variable_scope(name):
        if conv_type != 'conv2d':
            raise NotImplementedError('{0} is not supported!'.format(conv_type))
        if mask_type is not None:
            raise NotImplementedError('PixelCNN masks are not built (SURVEY 8(f) rank 4)')
        if input_dim != inputs.shape[-1]:
            raise ValueError('input_dim={} but inputs have {} channels'.format(input_dim, inputs.shape[-1]))

        def uniform(stdev, size):
            return np.random.uniform(low=-stdev * np.sqrt(3), high=stdev * np.sqrt(3), size=size).astype('float32')

        fan_in = input_dim * filter_size ** 2
        fan_out = output_dim * filter_size ** 2 / (stride ** 2)
        if he_init:
            filters_stdev = np.sqrt(4. / (fan_in + fan_out))
        else:  # Normalized init (Glorot & Bengio)
            filters_stdev = np.sqrt(2. / (fan_in + fan_out))
        stdev = _weights_stdev if _weights_stdev is not None else filters_stdev
        filter_values = _memo(
            lambda: uniform(stdev, (filter_size, filter_size, input_dim, output_dim)) * np.float32(gain))
        filters = store.get_variable(name='Filters', initializer=lambda _s: filter_values())

        if weightnorm is None:
            weightnorm = _default_weightnorm
        if weightnorm:
            raise NotImplementedError('weight-norm is not built (SURVEY 8(f) rank 4)')

        sn_entry = None
        if spectral_normed:
            with store.variable_scope('filters'):
                sn_entry = spectral_normed_weight(filters, update_collection=update_collection).entry

        in_scale = float(np.sqrt(2.0 / fan_in)) if inputs_norm else None
        _biases = None
        if biases:
            _biases = store.get_variable(name='Biases', shape=[output_dim, ],
                                         initializer=lambda s: np.zeros(s, dtype='float32'))
        if subpixel_up2:
            if (filter_size != 3 or stride != 1 or padding != 'SAME' or sn_entry is not None or in_scale is not None
                    or residual is not None):
                raise NotImplementedError('sub-pixel UpsampleConv: plain 3x3 stride-1 SAME layers only')
            return F.upconv2d(inputs, filters, _biases, out_grad_dtype=out_grad_dtype,
                              **({'out_dtype': out_dtype} if out_dtype is not None else {}))
        return F.conv2d(inputs, filters, _biases, filter_size, filter_size, stride, padding, sn=sn_entry,
                        residual=residual, out_grad_dtype=out_grad_dtype, in_scale=in_scale,
                        residual_up2=residual_up2, **({'out_dtype': out_dtype} if out_dtype is not None else {}))
