"""PGGAN (Nvidia architecture) on the B200 layer ops, with the reference's class, method names, arguments and variable
scopes (PGGAN/model_nvidia.py:15-237; config 5 of BASELINE.json: block_count 6 = 256x256, z 512, fade-in on).

Graph kept from the reference: G = Linear -> [pixel-norm, lrelu] -> 3x3 conv -> [pn, lrelu] -> block_count up-blocks
(nearest 2x, two 3x3 convs each followed by pixel-norm + lrelu) -> 1x1 toRGB, with the fade-in
(1 - alpha) * toRGB(upsample(previous stage)) + alpha * toRGB(new block) while `trans`; D mirrors it (1x1 fromRGB,
blocks of two spectrally-normalised 3x3 convs + lrelu + 2x2 average pool, fade-in against fromRGB of the pooled image,
minibatch-stddev channel, 3x3 conv over C+1 channels, lrelu, spatial mean, Linear).

Scheduling: pixel-norm + lrelu is ONE bandwidth-bound kernel that emits the bf16 tensor-core operand of the next
convolution; `inputs_norm` (x * sqrt(2 / fan_in), conv2d.py:93-95) never touches x -- it is the alpha of the GEMM
epilogue; the (C+1)-channel convolution behind minibatch_std runs on zero-padded operands (functional._conv2d_ragged_cin).
`alpha` (a placeholder fed per step in the reference, PGGAN/train.py:83) is a Python float or a 1-element fp32 device
tensor; the blend kernel reads it from device memory, so a captured CUDA graph follows the tensor's value."""
from __future__ import annotations

import torch

from .. import functional as F
from ..common.ops import conv2d as conv2d_ops
from ..common.ops import linear as linear_ops
from ..framework import get_store

BF16 = torch.bfloat16
F32 = torch.float32


def lrelu(x, leakiness=0.2, out_dtype=BF16):
    """model_nvidia.py:15-17 as one pass that also casts to the next operand's dtype."""
    assert leakiness <= 1, "leakiness must be <= 1"
    if leakiness != 0.2:
        raise NotImplementedError('only leakiness=0.2 (the value every reference call-site uses) is built')
    out, _ = F.norm_act(F.as_var(x), stats=None, act='lrelu', out_dtype=out_dtype)
    return out


def minibatch_std(x):
    """model_nvidia.py:20-28"""
    return F.minibatch_std(F.as_var(x))


def _pn_lrelu(x):
    """lib.ops.pixelnorm.Pixelnorm / normalization.pixel_norm followed by lrelu (:63-64, :68-69, :89-90, :95-96)."""
    return F.pixel_norm(x, act='lrelu', out_dtype=BF16)


class PGGAN(object):
    def __init__(self, args=None, block_count=None, trans=None, inputs_norm=None):
        """args: an argparse-like namespace with block_count / trans / inputs_norm (model_nvidia.py:32-39); the
        keywords override it."""
        self.bc = block_count if block_count is not None else args.block_count
        self.trans = trans if trans is not None else args.trans
        self.inputs_norm = inputs_norm if inputs_norm is not None else args.inputs_norm

    def get_dim(self, stage):
        """:41-47 (the reference returns a float under Python 3)."""
        return int(min(2048 / (2 ** stage), 512))

    def _conv(self, x, out_dim, k, name, **kw):
        return conv2d_ops.Conv2D(x, x.shape[-1], out_dim, k, 1, name, he_init=True, biases=True, **kw)

    def generator_block(self, inputs, out_dim, name='generator_block'):
        """:49-71"""
        store = get_store()
        with store.variable_scope(name):
            output = F.upsample2(F.as_var(inputs), out_dtype=BF16)
            output = self._conv(output, out_dim, 3, 'Conv.1', inputs_norm=self.inputs_norm)
            output = _pn_lrelu(output)
            output = self._conv(output, out_dim, 3, 'Conv.2', inputs_norm=self.inputs_norm)
            output = _pn_lrelu(output)
        return output

    def get_generator(self, z_var, alpha, training=True, reuse=False):
        """:73-129.  Returns Var [n, 4 * 2^bc, 4 * 2^bc, 3] (no tanh in the reference)."""
        store = get_store()
        with store.variable_scope('g_net', reuse=reuse):
            z_var_ = F.as_var(z_var)
            z_var_ = F.reshape(z_var_, (z_var_.shape[0], -1))
            output = linear_ops.Linear(z_var_, z_var_.shape[-1], 4 * 4 * 512, 'G.Input', inputs_norm=self.inputs_norm)
            output = F.reshape(output, (-1, 4, 4, 512))
            output = _pn_lrelu(output)
            output = self._conv(output, 512, 3, 'G.Conv', inputs_norm=self.inputs_norm)
            output = _pn_lrelu(output)
            for i in range(self.bc - 1):
                output = self.generator_block(output, self.get_dim(i), 'G.UpBlock.{}'.format(i + 1))
            if self.trans:
                toRGB1 = self.generator_block(output, self.get_dim(self.bc - 1), 'G.UpBlock.{}'.format(self.bc))
                toRGB1 = self._conv(toRGB1, 3, 1, 'G.{}_toRGB1'.format(self.bc), inputs_norm=self.inputs_norm)
                # skip connection
                toRGB2 = F.upsample2(output, out_dtype=BF16)
                toRGB2 = self._conv(toRGB2, 3, 1, 'G.{}_toRGB2'.format(self.bc), inputs_norm=self.inputs_norm)
                toRGB = F.lerp(toRGB2, toRGB1, alpha)   # fade in: (1 - alpha) * toRGB2 + alpha * toRGB1
            else:
                if self.bc > 0:
                    toRGB = self.generator_block(output, self.get_dim(self.bc - 1), 'G.UpBlock.{}'.format(self.bc))
                else:
                    toRGB = output
                toRGB = self._conv(toRGB, 3, 1, 'G.{}_toRGB'.format(self.bc), inputs_norm=self.inputs_norm)
        return toRGB

    def discriminator_block(self, inputs, out_dim, name, spectral_normed=False, update_collection=None, reuse=False):
        """:131-162"""
        store = get_store()
        kw = dict(spectral_normed=spectral_normed, update_collection=update_collection, reuse=reuse)
        with store.variable_scope(name):
            inputs = F.as_var(inputs)
            output = self._conv(inputs, inputs.shape[-1], 3, 'Conv.1', **kw)
            output = lrelu(output)
            output = self._conv(output, out_dim, 3, 'Conv.2', **kw)
            output = lrelu(output, out_dtype=F32)
            output = F.meanpool2(output)            # tf.nn.avg_pool 2x2 / 2 VALID
        return output

    def get_discriminator(self, x_var, alpha, spectral_normed=True, update_collection=None, reuse=False):
        """:164-237.  Returns the logits Var [n]."""
        store = get_store()
        kw = dict(spectral_normed=spectral_normed, update_collection=update_collection, reuse=reuse)
        with store.variable_scope('d_net', reuse=reuse):
            x_var = F.as_var(x_var)
            if self.trans:
                fromRGB1 = self._conv(x_var, self.get_dim(self.bc - 1), 1, 'D.{}_fromRGB1'.format(self.bc), **kw)
                fromRGB1 = self.discriminator_block(fromRGB1, self.get_dim(self.bc - 1), 'D.Block.{}'.format(self.bc),
                                                    **kw)
                # skip connection
                fromRGB2 = F.meanpool2(x_var)
                fromRGB2 = self._conv(fromRGB2, self.get_dim(self.bc - 1), 1, 'D.{}_fromRGB2'.format(self.bc), **kw)
                x_code = F.lerp(fromRGB2, fromRGB1, alpha)   # fade in
            else:
                x_code = self._conv(x_var, self.get_dim(self.bc - 1), 1, 'D.{}_fromRGB'.format(self.bc), **kw)
                if self.bc > 0:
                    x_code = self.discriminator_block(x_code, self.get_dim(self.bc - 1),
                                                      'D.Block.{}'.format(self.bc), **kw)
            for i in range(1, self.bc):
                x_code = self.discriminator_block(x_code, self.get_dim(self.bc - 1 - i),
                                                  'D.Block.{}'.format(self.bc - i), **kw)
            output = minibatch_std(x_code)
            output = self._conv(output, self.get_dim(self.bc - 1), 3, 'D.Conv', **kw)
            output = F.act_mean_hw(output, 'lrelu')       # lrelu + tf.reduce_mean(axis=[1, 2])
            logits = linear_ops.Linear(output, output.shape[-1], 1, 'D.Output')
            return F.reshape(logits, (-1,))
