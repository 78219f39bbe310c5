"""Convolution for data in format of 'NHWC' with the reference's signature (common/ops/conv2d.py:31-218).

Initial values come from the NumPy global RNG exactly as in the reference (uniform +-stdev*sqrt(3), He or
Glorot stdev, conv2d.py:83-140), and are drawn on every call during graph construction.
"""
from __future__ import annotations

import numpy as np
import torch

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


def pixelcnn_mask(mask_type, filter_size, input_dim, output_dim):
    """The constant filter mask of common/ops/conv2d.py:63-81: mask_type = ('a' | 'b', n_channels)."""
    kind, mask_n_channels = mask_type
    mask = np.ones((filter_size, filter_size, input_dim, output_dim), dtype='float32')
    center = filter_size // 2
    mask[center + 1:, :, :, :] = 0.          # future rows
    mask[center, center + 1:, :, :] = 0.     # future columns of the centre row
    for i in range(mask_n_channels):         # future channels at the centre tap
        for j in range(mask_n_channels):
            if (kind == 'a' and i >= j) or (kind == 'b' and i > j):
                mask[center, center, i::mask_n_channels, j::mask_n_channels] = 0.
    return mask


def _depthwise_types(store, inputs, conv_type, filters, depthwise_filters, pointwise_filters, input_dim, output_dim,
                     channel_multiplier, stride, padding, spectral_normed, update_collection, inputs_norm, mask_type,
                     weightnorm, biases, residual, subpixel_up2, out_grad_dtype, out_dtype):
    """conv_type 'depthwise_conv2d' / 'separable_conv2d' (common/ops/conv2d.py:188-208).

    Weight-norm and the PixelCNN mask act on `Filters` only (conv2d.py:153-167), which these two types never read.
    Spectral norm wraps ALL THREE filters (conv2d.py:169-178), each under its own scope -- `filters/spectral_norm/u`,
    `depthwise_filters/spectral_norm/u`, `pointwise_filters/spectral_norm/u` -- and the normalised depthwise / pointwise
    filters are what the op consumes.  A normalised filter that no op reads (`Filters` always; `pointwise_filters` of a
    plain depthwise layer) never runs its control-dependent u.assign in TensorFlow: its `u` variable is created and
    keeps its initial value, no power iteration is evaluated for it."""
    if depthwise_filters is None:
        raise ValueError('{0} needs channel_multiplier > 0 (the reference fails with a NameError)'.format(conv_type))
    if inputs_norm or residual is not None or subpixel_up2:
        raise NotImplementedError('{0} with inputs_norm / a fused residual (no reference call-site)'.format(conv_type))
    if weightnorm is None:
        weightnorm = _default_weightnorm
    if weightnorm or mask_type is not None:
        raise NotImplementedError('weight-norm / masks act on the unused `Filters` of {0}'.format(conv_type))
    sn_dw = sn_pw = None
    if spectral_normed:
        from ...framework import truncated_normal

        def dead_u(scope, c):
            with store.variable_scope(scope), store.variable_scope('spectral_norm'):
                store.get_variable('u', shape=[1, c], trainable=False,
                                   initializer=lambda s: truncated_normal(s, store.u_rng))

        dead_u('filters', output_dim)                                            # conv2d.py:170-171
        with store.variable_scope('depthwise_filters'):                          # conv2d.py:173-175
            sn_dw = spectral_normed_weight(depthwise_filters, update_collection=update_collection).entry
        if conv_type == 'separable_conv2d':                                      # conv2d.py:176-178
            with store.variable_scope('pointwise_filters'):
                sn_pw = spectral_normed_weight(pointwise_filters, update_collection=update_collection).entry
        else:
            dead_u('pointwise_filters', output_dim)
    _biases = None
    if biases:
        _biases = store.get_variable(name='Biases', shape=[output_dim, ],
                                     initializer=lambda s: np.zeros(s, dtype='float32'))
    if conv_type == 'depthwise_conv2d':
        if biases and output_dim != input_dim * channel_multiplier:
            raise ValueError('depthwise_conv2d yields input_dim * channel_multiplier = {} channels but Biases has '
                             'output_dim = {} (tf.nn.bias_add fails in the reference)'.format(
                                 input_dim * channel_multiplier, output_dim))
        return F.depthwise_conv2d(inputs, depthwise_filters, _biases, stride, padding,
                                  out_dtype=out_dtype or torch.float32, out_grad_dtype=out_grad_dtype, sn=sn_dw)
    mid = F.depthwise_conv2d(inputs, depthwise_filters, None, stride, padding, sn=sn_dw)
    return F.conv2d(mid, pointwise_filters, _biases, 1, 1, 1, 'VALID', sn=sn_pw, out_grad_dtype=out_grad_dtype,
                    **({'out_dtype': out_dtype} if out_dtype is not None else {}))


def Conv2D(inputs, input_dim, output_dim, filter_size=3, stride=1, name='Conv2D',
           conv_type='conv2d', channel_multiplier=0, padding='SAME',
           spectral_normed=False, update_collection=None, inputs_norm=False, he_init=True,
           mask_type=None, weightnorm=None, biases=True, gain=1., reuse=None,
           residual=None, out_grad_dtype=None, residual_up2=False, out_dtype=None, subpixel_up2=False,
           bn_stats=False, fused_act=None):
    """
    Args mirror common/ops/conv2d.py:31-55 (`reuse` is the extra keyword of conv2d_.py:33, accepted and ignored).
    `residual` (fp32 Var added in the GEMM epilogue; `residual_up2`: given at half resolution) and `out_grad_dtype`
    are extensions used by resnet_block.  `subpixel_up2`: the layer is UpsampleConv (nearest 2x in front of this 3x3
    convolution, common/resnet_block.py:83-97) and `inputs` is the LOW-resolution tensor: evaluated in sub-pixel
    form, output in quad layout (functional.upconv2d).  `bn_stats`: the output feeds a batch-statistics normalisation
    (the convolution epilogue then also produces its per-channel sums, functional.conv2d).  `fused_act`: the nonlinearity
    behind the layer, applied by the epilogue (functional.conv2d(act=...); the consumer must be a plain Conv2D).

    Returns:
      Var of shape (batch_size, out_height, out_width, output_dim), fp32
    """
    store = get_store()
    inputs = F.as_var(inputs)
    with store.variable_scope(name):
        if conv_type not in ('conv2d', 'depthwise_conv2d', 'separable_conv2d'):
            raise NotImplementedError('{0} is not supported!'.format(conv_type))      # conv2d.py:209-210
        if input_dim != inputs.shape[-1]:
            raise ValueError('input_dim={} but inputs have {} channels'.format(input_dim, inputs.shape[-1]))

        def uniform(stdev, size):
            return np.random.uniform(low=-stdev * np.sqrt(3), high=stdev * np.sqrt(3), size=size).astype('float32')

        fan_in = input_dim * filter_size ** 2
        fan_out = output_dim * filter_size ** 2 / (stride ** 2)
        in_scale = float(np.sqrt(2.0 / fan_in)) if inputs_norm else None   # conv2d.py:93-95, before the halving below
        if mask_type is not None:  # "only approximately correct" (conv2d.py:99-101)
            fan_in /= 2.
            fan_out /= 2.
        if he_init:
            filters_stdev = np.sqrt(4. / (fan_in + fan_out))
        else:  # Normalized init (Glorot & Bengio)
            filters_stdev = np.sqrt(2. / (fan_in + fan_out))
        stdev = _weights_stdev if _weights_stdev is not None else filters_stdev
        filter_values = _memo(
            lambda: uniform(stdev, (filter_size, filter_size, input_dim, output_dim)) * np.float32(gain))
        filters = store.get_variable(name='Filters', initializer=lambda _s: filter_values())
        depthwise_filters = pointwise_filters = None
        if channel_multiplier > 0:   # conv2d.py:117-126, 145-150: drawn after the filter values, never scaled by gain
            depthwise_filters = store.get_variable(name='depthwise_filters', initializer=lambda _s: uniform(
                stdev, (filter_size, filter_size, input_dim, channel_multiplier)))
            pointwise_filters = store.get_variable(name='pointwise_filters', initializer=lambda _s: uniform(
                stdev, (1, 1, input_dim * channel_multiplier, output_dim)))
        if fused_act is not None and (conv_type != 'conv2d' or subpixel_up2):
            raise NotImplementedError('fused_act: plain conv2d layers only')
        if conv_type != 'conv2d':
            return _depthwise_types(store, inputs, conv_type, filters, depthwise_filters, pointwise_filters, input_dim,
                                    output_dim, channel_multiplier, stride, padding, spectral_normed,
                                    update_collection, inputs_norm, mask_type, weightnorm, biases, residual,
                                    subpixel_up2, out_grad_dtype, out_dtype)

        if weightnorm is None:
            weightnorm = _default_weightnorm
        target_norms = None
        if weightnorm:   # conv2d.py:153-163: g initialised to the norms of the INITIAL values, filters * (g / norms)
            target_norms = store.get_variable(
                name='g', initializer=lambda _s: np.sqrt(np.sum(np.square(filter_values()), axis=(0, 1, 2))))
        mask_fn = None
        if mask_type is not None:   # conv2d.py:165-167
            mask_fn = lambda: pixelcnn_mask(mask_type, filter_size, input_dim, output_dim)  # noqa: E731
        filters = store.effective_weight(filters, target_norms, mask_fn,
                                         (filter_size * filter_size * input_dim, output_dim, 1))

        sn_entry = None
        if spectral_normed:
            with store.variable_scope('filters'):
                sn_entry = spectral_normed_weight(filters, update_collection=update_collection).entry

        _biases = None
        if biases:
            _biases = store.get_variable(name='Biases', shape=[output_dim, ],
                                         initializer=lambda s: np.zeros(s, dtype='float32'))
        if subpixel_up2:
            if (filter_size != 3 or stride != 1 or padding != 'SAME' or sn_entry is not None or in_scale is not None
                    or residual is not None):
                raise NotImplementedError('sub-pixel UpsampleConv: plain 3x3 stride-1 SAME layers only')
            return F.upconv2d(inputs, filters, _biases, out_grad_dtype=out_grad_dtype, bn_stats=bn_stats,
                              **({'out_dtype': out_dtype} if out_dtype is not None else {}))
        return F.conv2d(inputs, filters, _biases, filter_size, filter_size, stride, padding, sn=sn_entry,
                        residual=residual, out_grad_dtype=out_grad_dtype, in_scale=in_scale,
                        residual_up2=residual_up2, bn_stats=bn_stats, act=fused_act,
                        **({'out_dtype': out_dtype} if out_dtype is not None else {}))
