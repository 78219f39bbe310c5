"""Transposed convolution with the reference's signature (common/ops/deconv2d.py:29-118).

The reference has no call-site for it.  The op is the data gradient of a stride-2 convolution: it runs as the
stride-1 tensor-core kernel over the zero-dilated input (functional.conv2d_transpose); its own backward uses the
strided forward / filter-gradient kernels.  weight-norm (deconv2d.py:87-96) is not built."""
from __future__ import annotations

import numpy as np

from ... import functional as F
from ...framework import get_store

_default_weightnorm = False
_weights_stdev = None


def enable_default_weightnorm():
    global _default_weightnorm
    _default_weightnorm = True


def set_weights_stdev(weights_stdev):
    global _weights_stdev
    _weights_stdev = weights_stdev


def unset_weights_stdev():
    global _weights_stdev
    _weights_stdev = None


def Deconv2D(inputs, in_channels, output_channels, filter_size, stride=2, padding='SAME', he_init=True,
             weight_norm=None, gain=1., mask_type=None, biases=True, name='Deconv2D'):
    store = get_store()
    with store.variable_scope(name):
        if mask_type is not None:
            raise Exception('Unsupported configuration in Deconv2D!')
        if stride != 2:
            raise ValueError('Deconv2D always produces a 2x output (deconv2d.py:99-100); stride must be 2')
        fan_in = in_channels * filter_size ** 2 / (stride ** 2)
        fan_out = output_channels * filter_size ** 2
        stdev = np.sqrt((4. if he_init else 2.) / (fan_in + fan_out))
        if _weights_stdev is not None:
            stdev = _weights_stdev
        box = []

        def filter_values():
            if not box:
                box.append(np.random.uniform(
                    low=-stdev * np.sqrt(3), high=stdev * np.sqrt(3),
                    size=(filter_size, filter_size, output_channels, in_channels)).astype('float32') * np.float32(gain))
            return box[0]

        filters = store.get_variable(name='Filters', initializer=lambda _s: filter_values())
        if weight_norm is None:
            weight_norm = _default_weightnorm
        if weight_norm:   # deconv2d.py:87-96: one norm per OUTPUT channel (axis 2 of [k, k, Cout, Cin])
            target_norms = store.get_variable(
                name='g', initializer=lambda _s: np.sqrt(np.sum(np.square(filter_values()), axis=(0, 1, 3))))
            filters = store.effective_weight(filters, target_norms, None,
                                             (filter_size * filter_size, output_channels, in_channels))
        _biases = None
        if biases:
            _biases = store.get_variable(name='Biases', shape=[output_channels, ],
                                         initializer=lambda s: np.zeros(s, dtype='float32'))
        return F.conv2d_transpose(F.as_var(inputs), filters, _biases, filter_size, filter_size, stride, padding)
