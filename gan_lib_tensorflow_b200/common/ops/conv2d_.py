"""Older Conv2D signature (common/ops/conv2d_.py:32-160): extra `reuse` keyword, padding fixed to SAME and the
spectral-norm scope 'spectral_norm/u' WITHOUT the 'filters/' level.  PGGAN/model_nvidia.py is written against it."""
from __future__ import annotations

import numpy as np

from ... import functional as F
from ...framework import get_store
from . import conv2d as _new
from .sn import spectral_normed_weight


def Conv2D(inputs, input_dim, output_dim, filter_size=3, stride=1, name='Conv2D',
           spectral_normed=False, update_collection=None, reuse=False, inputs_norm=False, he_init=True,
           mask_type=None, weightnorm=None, biases=True, gain=1.):
    store = get_store()
    inputs = F.as_var(inputs)
    with store.variable_scope(name):
        fan_in = input_dim * filter_size ** 2
        fan_out = output_dim * filter_size ** 2 / (stride ** 2)
        in_scale = float(np.sqrt(2.0 / fan_in)) if inputs_norm else None
        if mask_type is not None:   # conv2d_.py halves the fans like conv2d.py:99-101
            fan_in /= 2.
            fan_out /= 2.
        stdev = np.sqrt((4. if he_init else 2.) / (fan_in + fan_out))
        if _new._weights_stdev is not None:
            stdev = _new._weights_stdev
        fv = _new._memo(lambda: np.random.uniform(low=-stdev * np.sqrt(3), high=stdev * np.sqrt(3),
                                                  size=(filter_size, filter_size, input_dim, output_dim)
                                                  ).astype('float32') * np.float32(gain))
        filters = store.get_variable(name='Filters', initializer=lambda _s: fv())
        if weightnorm is None:
            weightnorm = _new._default_weightnorm
        target_norms = None
        if weightnorm:
            target_norms = store.get_variable(
                name='g', initializer=lambda _s: np.sqrt(np.sum(np.square(fv()), axis=(0, 1, 2))))
        mask_fn = None
        if mask_type is not None:
            mask_fn = lambda: _new.pixelcnn_mask(mask_type, filter_size, input_dim, output_dim)  # noqa: E731
        filters = store.effective_weight(filters, target_norms, mask_fn,
                                         (filter_size * filter_size * input_dim, output_dim, 1))
        sn_entry = None
        if spectral_normed:
            sn_entry = spectral_normed_weight(filters, update_collection=update_collection).entry  # conv2d_.py:130
        _biases = None
        if biases:
            _biases = store.get_variable(name='Biases', shape=[output_dim, ],
                                         initializer=lambda s: np.zeros(s, dtype='float32'))
        return F.conv2d(inputs, filters, _biases, filter_size, filter_size, stride, 'SAME', sn=sn_entry,
                        in_scale=in_scale)
