"""Dense layer with the reference's signature (common/ops/linear.py:38-182)."""
from __future__ import annotations

import numpy as np

from ... import functional as F
from ...framework import get_store
from .conv2d import _memo
from .sn import spectral_normed_weight

_default_weightnorm = False


def disable_default_weightnorm():
    global _default_weightnorm
    _default_weightnorm = False


_weights_stdev = None


def unset_weights_stdev():
    global _weights_stdev
    _weights_stdev = None


def Linear(inputs, input_dim, output_dim, name,
           spectral_normed=False, update_collection=None, reuse=False, inputs_norm=False,
           biases=True, initialization=None, weightnorm=None, gain=1., out_dtype=None):
    """
    initialization: None, `lecun`, 'glorot', `he`, 'glorot_he', `orthogonal`, `("uniform", range)`
    """
    store = get_store()
    inputs = F.as_var(inputs)
    with store.variable_scope(name):
        in_scale = float(np.sqrt(2.0 / input_dim)) if inputs_norm else None   # linear.py:47-49
        kw = {'in_scale': in_scale} if in_scale is not None else {}

        shape = (input_dim, output_dim)
        # stdev of the uniform(+-stdev*sqrt(3)) draw per scheme (linear.py:63-136); None / 'glorot' / 'xavier' share
        # the Glorot value, which also makes the reference's "orthogonal if square" clause unreachable (:112-113)
        stdevs = {
            'lecun': np.sqrt(1. / input_dim),
            None: np.sqrt(2. / (input_dim + output_dim)),
            'glorot': np.sqrt(2. / (input_dim + output_dim)),
            'xavier': np.sqrt(2. / (input_dim + output_dim)),
            'he': np.sqrt(2. / input_dim),
            'glorot_he': np.sqrt(4. / (input_dim + output_dim)),
        }

        def draw():
            if isinstance(initialization, (tuple, list)) and initialization[0] == 'uniform':
                bound = initialization[1]
                wv = np.random.uniform(low=-bound, high=bound, size=shape).astype('float32')
            elif initialization == 'orthogonal':
                u, _, v = np.linalg.svd(np.random.normal(0.0, 1.0, shape), full_matrices=False)
                wv = (u if u.shape == shape else v).reshape(shape).astype('float32')
            elif initialization in stdevs:
                stdev = _weights_stdev if _weights_stdev is not None else stdevs[initialization]
                wv = np.random.uniform(low=-stdev * np.sqrt(3), high=stdev * np.sqrt(3), size=shape).astype('float32')
            else:
                raise Exception('Invalid initialization!')
            return wv * np.float32(gain)

        weight_values = _memo(draw)
        weight = store.get_variable(name='W', initializer=lambda _s: weight_values())
        if weightnorm is None:
            weightnorm = _default_weightnorm
        if weightnorm:   # linear.py:143-155: norms over axis 0
            target_norms = store.get_variable(
                name='g', initializer=lambda _s: np.sqrt(np.sum(np.square(weight_values()), axis=0)))
            weight = store.effective_weight(weight, target_norms, None, (input_dim, output_dim, 1))
        sn_entry = None
        if spectral_normed:
            sn_entry = spectral_normed_weight(weight, update_collection=update_collection).entry
        _biases = None
        if biases:
            _biases = store.get_variable(name='b', shape=[output_dim, ],
                                         initializer=lambda s: np.zeros(s, dtype='float32'))
        if len(inputs.shape) == 2:
            return F.linear(inputs, weight, _biases, sn=sn_entry, **kw,
                            **({'out_dtype': out_dtype} if out_dtype is not None else {}))
        lead = inputs.shape[:-1]
        flat = F.reshape(inputs, (-1, input_dim))
        result = F.linear(flat, weight, _biases, sn=sn_entry, **kw)
        return F.reshape(result, tuple(lead) + (output_dim,))
