"""Normalisation layers with the reference's signatures (common/ops/normalization.py).

Each accepts the extension keywords act / upsample / out_dtype / want_raw so that common/resnet_block.py can run
normalise + nonlinearity + nearest-upsample + bf16 cast as ONE kernel; with the defaults they behave like the
reference functions (fp32 in, fp32 out, no activation)."""
from __future__ import annotations

import numpy as np
import torch

from ... import functional as F
from ...framework import get_store


def _labels_i32(labels, store):
    t = labels.data if isinstance(labels, F.Var) else labels
    return t.to(device=store.device, dtype=torch.int32).reshape(-1).contiguous()


def batch_norm(inputs, decay=0.9, epsilon=1e-5, is_training=True, fused=True, act=None, upsample=False,
               out_dtype=torch.float32, want_raw=False):
    """common/ops/normalization.py:8-24: contrib fused batch norm, always in training mode (batch statistics).
    The moving averages are write-only state in every reference caller and are not maintained yet."""
    if not is_training:
        # inference mode would need the moving averages, which no reference caller ever reads (is_training=True everywhere)
        raise NotImplementedError('batch_norm(is_training=False): moving averages are not maintained')
    store = get_store()
    inputs = F.as_var(inputs)
    with store.variable_scope('BatchNorm'):
        c = inputs.shape[-1]
        beta = store.get_variable('beta', shape=[c], initializer=lambda s: np.zeros(s, dtype='float32'))
        gamma = store.get_variable('gamma', shape=[c], initializer=lambda s: np.ones(s, dtype='float32'))
        out, raw = F.norm_act(inputs, stats='batch', eps=epsilon, gamma=gamma, beta=beta, labels=None, act=act,
                              upsample=upsample, out_dtype=out_dtype, want_raw=want_raw)
        return (out, raw) if want_raw else out


def cond_batchnorm(name, axes, inputs, is_training=None, stats_iter=None, update_moving_stats=True, fused=True,
                   labels=None, n_labels=None, act=None, upsample=False, out_dtype=torch.float32, want_raw=False):
    """Conditional Batchnorm (dumoulin et al 2016) for BHWC conv filtermaps -- common/ops/normalization.py:27-59:
    batch moments (population variance), per-class offset/scale gathered by label, eps 1e-5, no moving averages."""
    store = get_store()
    inputs = F.as_var(inputs)
    with store.variable_scope('CondBatchNorm'):
        if axes != [0, 1, 2]:
            raise Exception('Axes is not supported in Conditional BatchNorm!')
        c = inputs.shape[3]
        offset_m = store.get_variable(name='offset', shape=[n_labels, c],
                                      initializer=lambda s: np.zeros(s, dtype='float32'))
        scale_m = store.get_variable(name='scale', shape=[n_labels, c],
                                     initializer=lambda s: np.ones(s, dtype='float32'))
        out, raw = F.norm_act(inputs, stats='batch', eps=1e-5, gamma=scale_m, beta=offset_m,
                              labels=_labels_i32(labels, store), act=act, upsample=upsample, out_dtype=out_dtype,
                              want_raw=want_raw)
        return (out, raw) if want_raw else out


def layer_norm(name, norm_axes, inputs, act=None, out_dtype=torch.float32):
    """common/ops/normalization.py:62-82: tf.contrib.layers.layer_norm(begin_norm_axis=1, begin_params_axis=-1,
    scope=name) -- moments over (h, w, c) per sample, `beta` then `gamma` of shape [c] under the scope `name`
    (reached with NORMALIZATION_D=True in the SNGAN scripts).  norm_axes is ignored by the reference as well.
    `act` fuses the nonlinearity that follows it in the residual blocks."""
    store = get_store()
    inputs = F.as_var(inputs)
    with store.variable_scope(name):
        c = inputs.shape[-1]
        beta = store.get_variable('beta', shape=[c], initializer=lambda s: np.zeros(s, dtype='float32'))
        gamma = store.get_variable('gamma', shape=[c], initializer=lambda s: np.ones(s, dtype='float32'))
        return F.layer_norm(inputs, gamma, beta, eps=1e-12, act=act, out_dtype=out_dtype)


def instance_norm(inputs, epsilon=1e-06, act=None, upsample=False, out_dtype=torch.float32):
    """common/ops/normalization.py:105-122: per-(n, c) moments over (h, w)."""
    store = get_store()
    inputs = F.as_var(inputs)
    with store.variable_scope('InstanceNorm'):
        c = inputs.shape[-1]
        beta = store.get_variable('beta', shape=[c], initializer=lambda s: np.zeros(s, dtype='float32'))
        gamma = store.get_variable('gamma', shape=[c], initializer=lambda s: np.ones(s, dtype='float32'))
        out, _ = F.norm_act(inputs, stats='instance', eps=epsilon, gamma=gamma, beta=beta, labels=None, act=act,
                            upsample=upsample, out_dtype=out_dtype)
        return out


def pixel_norm(inputs, eps=1e-8, act=None, out_dtype=None):
    """common/ops/normalization.py:125-140 (PGGAN): inputs * rsqrt(mean over channels of inputs^2 + eps).
    `act` fuses the leaky ReLU that follows it in PGGAN/model_nvidia.py:63-68."""
    return F.pixel_norm(F.as_var(inputs), eps=eps, act=act, out_dtype=out_dtype)
