"""Residual-block library with the reference's names and arguments (common/resnet_block.py:20-184).

Same graph as the reference (pre-activation ResNet blocks; sub-layers named name+'.Shortcut' / '.N1' / '.Conv1' /
'.N2' / '.Conv2'), scheduled for B200:
  * Normalize + nonlinearity (+ nearest upsample) + bf16 cast run as one bandwidth-bound kernel whose output is
    directly the tensor-core operand of the next convolution;
  * 1x1 shortcut convolutions commute with nearest-upsampling / mean-pooling and are evaluated at the low
    resolution (4x fewer FLOPs, identical dot products);
  * the residual sum is folded into the epilogue of Conv2, and one mean-pool serves shortcut + main path.
"""
from __future__ import annotations

import functools

import torch

from .. import functional as F
from ..framework import Var, get_store
from . import ops as _ops  # noqa: F401  (keeps `lib.ops.<module>` attribute access working)
from .ops import conv2d as _conv2d
from .ops import linear as _linear
from .ops import normalization as _norm

NORMALIZATION_G = True
NORMALIZATION_D = True

BF16 = torch.bfloat16
F32 = torch.float32


def _adt():
    """Storage dtype of the activations between the layers of the network being built: bf16 (tensor-core operands of the
    default mode) or fp32 when the network runs in TF32 operand mode (VariableStore.set_precision)."""
    return F32 if get_store().is_tf32() else BF16


def nonlinearity(x, activation_fn='relu', leakiness=0.2):
    """common/resnet_block.py:24-29 (the reference silently returns None for unknown names; this raises)."""
    if activation_fn == 'relu':
        return F.activation(F.as_var(x), 'relu')
    if activation_fn == 'lrelu':
        assert 0 < leakiness <= 1, "leakiness must be <= 1"
        if leakiness != 0.2:
            raise NotImplementedError('only leakiness=0.2 (the value every reference call-site uses) is built')
        return F.activation(F.as_var(x), 'lrelu')
    raise ValueError('unknown activation_fn {!r}'.format(activation_fn))


def _normalize_kind(name, labels, spectral_normed):
    """Dispatch of common/resnet_block.py:32-50 on the SUBSTRING of the layer name."""
    if ('D.' in name) and NORMALIZATION_D:
        return None if spectral_normed else 'bn'
    elif ('G.' in name) and NORMALIZATION_G:
        return 'cbn' if labels is not None else 'bn'
    return None


def _norm_act(name, inputs, labels, kind, act, upsample=False, want_raw=False, out_dtype=None, n_labels=10):
    """Normalize(name, inputs) followed by nonlinearity(), fused. Returns (out, raw_bf16_or_None).
    n_labels: 10 is hard-wired in the library's Normalize (resnet_block.py:43) and the CIFAR script; the ImageNet
    script's copy uses 1000 (gan_imagNet_resnet.py:104)."""
    store = get_store()
    if out_dtype is None:
        out_dtype = _adt()
    with store.variable_scope(name):
        if kind == 'cbn':
            res = _norm.cond_batchnorm(name, [0, 1, 2], inputs, labels=labels, n_labels=n_labels, act=act,
                                       upsample=upsample, out_dtype=out_dtype, want_raw=want_raw)
        elif kind == 'bn':
            res = _norm.batch_norm(inputs, fused=True, act=act, upsample=upsample, out_dtype=out_dtype,
                                   want_raw=want_raw)
        elif kind == 'ln':
            # layer norm of the critic (the SNGAN scripts' NORMALIZATION_D switch); D blocks never upsample
            if upsample:
                raise NotImplementedError('layer_norm in front of an upsample (no reference call-site)')
            inputs = F.as_var(inputs)
            out = _norm.layer_norm(name, [1, 2, 3], inputs, act=act, out_dtype=out_dtype)
            raw = None
            if want_raw:
                raw = inputs if inputs.data.dtype == BF16 else F.cast(inputs, BF16)
            return out, raw
        else:
            return F.norm_act(inputs, stats=None, act=act, upsample=upsample, out_dtype=out_dtype, want_raw=want_raw)
    return res if want_raw else (res, None)


def Normalize(name, inputs, labels=None, spectral_normed=True):
    """common/resnet_block.py:32-50, un-fused (fp32 in, fp32 out)."""
    inputs = F.as_var(inputs)
    kind = _normalize_kind(name, labels, spectral_normed)
    if kind is None:
        with get_store().variable_scope(name):
            return inputs
    out, _ = _norm_act(name, inputs, labels, kind, None, out_dtype=F32)
    return out


def ConvMeanPool(inputs, output_dim, filter_size=3, stride=1, name=None,
                 spectral_normed=False, update_collection=None, inputs_norm=False,
                 he_init=True, biases=True):
    """common/resnet_block.py:53-64"""
    inputs = F.as_var(inputs)
    output = _conv2d.Conv2D(inputs, inputs.shape[-1], output_dim, filter_size, stride, name,
                            spectral_normed=spectral_normed, update_collection=update_collection,
                            inputs_norm=inputs_norm, he_init=he_init, biases=biases, out_grad_dtype=_adt())
    return F.meanpool2(output)


def MeanPoolConv(inputs, output_dim, filter_size=3, stride=1, name=None,
                 spectral_normed=False, update_collection=None, inputs_norm=False,
                 he_init=True, biases=True):
    """common/resnet_block.py:67-80"""
    inputs = F.as_var(inputs)
    output = F.meanpool2(inputs if inputs.dtype == F32 else F.cast(inputs, F32))
    return _conv2d.Conv2D(output, output.shape[-1], output_dim, filter_size, stride, name,
                          spectral_normed=spectral_normed, update_collection=update_collection,
                          inputs_norm=inputs_norm, he_init=he_init, biases=biases, out_grad_dtype=_adt())


def UpsampleConv(inputs, output_dim, filter_size=3, stride=1, name=None,
                 spectral_normed=False, update_collection=None, inputs_norm=False,
                 he_init=True, biases=True):
    """common/resnet_block.py:83-97"""
    inputs = F.as_var(inputs)
    output = F.upsample2(inputs, out_dtype=BF16 if inputs.shape[-1] > 8 else None)
    return _conv2d.Conv2D(output, output.shape[-1], output_dim, filter_size, stride, name,
                          spectral_normed=spectral_normed, update_collection=update_collection,
                          inputs_norm=inputs_norm, he_init=he_init, biases=biases)


def ResidualBlock(inputs, input_dim, output_dim, filter_size, name,
                  spectral_normed=False, update_collection=None, inputs_norm=False,
                  resample=None, labels=None, biases=True, activation_fn='relu',
                  normalize_kind=None, pre_activated=None, out_dtype=None, n_labels=10, out_bn_stats=False):
    """resample: None, 'down', or 'up' -- common/resnet_block.py:100-156.

    out_dtype=torch.bfloat16 stores the block output (the input of the next block's normalisation) in bf16; the sum
    conv2 + shortcut is formed in fp32 inside the GEMM epilogue and rounded once.

    normalize_kind overrides the name-based Normalize dispatch (used by the SNGAN scripts' own Normalize).
    pre_activated = (raw_bf16, act_bf16) lets a producer that already emitted both operands (the label-map
    concat of D) skip the block's first activation pass.
    out_bn_stats: the block output feeds a batch-statistics normalisation (the next block's N1 / the output norm of a
    generator): Conv2's epilogue then leaves the per-channel sums next to the output (functional.conv2d)."""
    if resample not in (None, 'down', 'up'):
        raise Exception('invalid resample value')
    if activation_fn not in ('relu', 'lrelu'):
        raise ValueError('unknown activation_fn {!r}'.format(activation_fn))
    conv = functools.partial(_conv2d.Conv2D, filter_size=filter_size, spectral_normed=spectral_normed,
                             update_collection=update_collection, inputs_norm=inputs_norm, biases=biases)
    kind = normalize_kind if normalize_kind is not None else (
        lambda nm: _normalize_kind(nm, labels, spectral_normed))
    identity_shortcut = (output_dim == input_dim and resample is None)

    # 'up' blocks: Conv1 = UpsampleConv runs in sub-pixel form (four 2x2 convolutions over the LOW-resolution
    # activation, 4/9 of the MMA work, no upsampled operand) wherever the CTA-pair kernel tiles the shape
    subpixel = False
    if resample == 'up' and pre_activated is None and not spectral_normed and not inputs_norm and _adt() == BF16:
        n_, h_, w_, _c = F.as_var(inputs).shape
        subpixel = F.upconv_eligible(n_, h_, w_, input_dim, output_dim, filter_size)

    # ---- N1 + nonlinearity (+ upsample), and the raw bf16 copy feeding the 1x1 shortcut
    if pre_activated is not None:
        raw, a1 = pre_activated
        x32 = None
    else:
        x32 = F.as_var(inputs)
        n1_in = x32
        if kind(name + '.N1') in ('cbn', 'bn') and x32.data.dtype == F32 and _adt() == BF16:
            # every input of a batch-statistics normalisation is stored in bf16 (DESIGN.md "Data layout"): an fp32
            # residual stream (ACGAN's batch-normed D) is rounded once here; statistics, normalise and the shortcut
            # operand then read 2-byte elements, the identity shortcut keeps the fp32 tensor
            n1_in = F.cast(x32, BF16)
        if kind(name + '.N1') is None and input_dim % 4 and resample != 'up':
            # an RGB-sided block (D.*_fromRGB of the ResNet PGGAN critic, resnet_block.py:283-311): the activation
            # runs on the flat fp32 tensor and both small-channel convolutions read fp32
            a1, raw = F.activation(x32 if x32.data.dtype == F32 else F.cast(x32, F32), activation_fn), x32
        else:
            a1, raw = _norm_act(name + '.N1', n1_in, labels, kind(name + '.N1'), activation_fn,
                                upsample=(resample == 'up' and not subpixel), want_raw=not identity_shortcut,
                                n_labels=n_labels)

    # ---- shortcut (reference order: the shortcut variables are created before Conv1's)
    if identity_shortcut:
        shortcut = x32
    else:
        # ConvMeanPool / UpsampleConv / Conv2D with a 1x1 filter, he_init=False (resnet_block.py:123-127)
        shortcut = _conv2d.Conv2D(raw, input_dim, output_dim, 1, 1, name + '.Shortcut',
                                  spectral_normed=spectral_normed, update_collection=update_collection,
                                  inputs_norm=inputs_norm, he_init=False, biases=biases, out_grad_dtype=_adt())
        # 'up': conv1x1(upsample(x)) == upsample(conv1x1(x)); the upsample itself happens inside Conv2's epilogue

    # ---- Conv1
    mid_dim = input_dim if resample == 'down' else output_dim
    # h1 is only consumed by N2 + nonlinearity, whose backward can emit the bf16 tensor-core operand directly
    # and which is stored in bf16: it is only read by that kernel (rounding commutes with relu / leaky relu, so
    # without a normalisation in between this is bit-identical to rounding after the activation)
    # without a normalisation between the two convolutions (critic blocks) the nonlinearity is applied by Conv1's
    # epilogue and its derivative by Conv2's data-gradient epilogue: no pass over h1 in either direction
    fuse_act = (kind(name + '.N2') is None and not subpixel and
                F.conv2d_act_fusable(input_dim, mid_dim, output_dim, filter_size))
    h1 = conv(a1, input_dim, mid_dim, name=name + '.Conv1', he_init=True, out_grad_dtype=_adt(), out_dtype=_adt(),
              subpixel_up2=subpixel, bn_stats=kind(name + '.N2') in ('cbn', 'bn'),
              fused_act=activation_fn if fuse_act else None)

    # ---- N2 + nonlinearity
    if fuse_act:
        a2 = h1
    else:
        a2, _ = _norm_act(name + '.N2', h1, labels, kind(name + '.N2'), activation_fn, n_labels=n_labels)

    # ---- Conv2 (+ residual in the epilogue) [+ mean-pool of the sum]
    if resample == 'down':
        t = conv(a2, mid_dim, output_dim, name=name + '.Conv2', he_init=True, residual=shortcut,
                 out_grad_dtype=_adt())
        # meanpool(conv2) + meanpool(shortcut) == meanpool(conv2 + shortcut)
        return F.meanpool2(t, out_grad_dtype=_adt())
    # the block output feeds the next block's normalise/activation kernel and (through 1x1 / identity shortcuts)
    # convolutions only: its gradient is a tensor-core operand, so it is produced in bf16 directly
    return conv(a2, mid_dim, output_dim, name=name + '.Conv2', he_init=True, residual=shortcut,
                residual_up2=(resample == 'up' and not identity_shortcut), out_grad_dtype=_adt(),
                bn_stats=bool(out_bn_stats) and out_dtype == BF16,
                **({'out_dtype': out_dtype} if out_dtype is not None else {}))


def OptimizedResBlockDisc1(inputs, DIM_D=128, activation_fn='relu',
                           spectral_normed=False, update_collection=None, inputs_norm=False,
                           biases=True, name_prefix='D.DownBlock.1'):
    """common/resnet_block.py:159-184.  name_prefix='D.Block.1' gives the SNGAN script's copy
    (SNGAN/gan_cifar_resnet.py:212-234)."""
    inputs = F.as_var(inputs)
    cin = inputs.shape[-1]
    shortcut = MeanPoolConv(inputs=inputs, output_dim=DIM_D, filter_size=1, name=name_prefix + '.Shortcut',
                            spectral_normed=spectral_normed, update_collection=update_collection,
                            inputs_norm=inputs_norm, he_init=False, biases=biases)
    fuse_act = F.conv2d_act_fusable(cin, DIM_D, DIM_D, 3)     # as in ResidualBlock: relu inside Conv1's epilogue
    output = _conv2d.Conv2D(inputs, cin, DIM_D, 3, 1, name_prefix + '.Conv1', spectral_normed=spectral_normed,
                            update_collection=update_collection, inputs_norm=inputs_norm, he_init=True,
                            biases=biases, out_grad_dtype=_adt(), out_dtype=_adt(),
                            fused_act=activation_fn if fuse_act else None)
    if not fuse_act:
        output, _ = F.norm_act(output, stats=None, act=activation_fn, out_dtype=_adt())
    output = _conv2d.Conv2D(output, DIM_D, DIM_D, 3, 1, name_prefix + '.Conv2', spectral_normed=spectral_normed,
                            update_collection=update_collection, inputs_norm=inputs_norm, he_init=True,
                            biases=biases, out_grad_dtype=_adt())
    return F.meanpool2(output, addend=shortcut, out_grad_dtype=_adt())


# ######## ######## PGGAN ######## ######## #
def get_dim(stage):
    """common/resnet_block.py:188-189 (returns a float under Python 3 in the reference; int here)."""
    return int(min(2048 / (2 ** stage), 512))


def Generator_PGGAN(noise, bc, trans=False, alpha=0.01, inputs_norm=False, labels=None, training=True):
    """common/resnet_block.py:192-263 -- the ResNet PGGAN generator (PGGAN/model_resnet.py:24-39): Linear to
    [n, 4, 4, 1024], batch norm + relu, 3x3 conv, bc 'up' residual blocks of get_dim(i) channels, a toRGB residual
    block, batch norm + relu, 3x3 conv to RGB, tanh.  With `trans` the last 'up' block + toRGB1 block is blended
    with toRGB2 = a residual block over the nearest-2x upsample of the previous stage (fade-in over FEATURE maps, :237).
    alpha: Python float or 1-element device tensor (functional.lerp)."""
    noise = F.as_var(noise)
    output = _linear.Linear(noise, noise.shape[-1], 4 * 4 * 1024, 'G.Input', inputs_norm=inputs_norm, biases=True,
                            initialization=None, out_dtype=BF16)
    output = F.reshape(output, (-1, 4, 4, 1024))
    output, _ = _norm_act('G.N0', output, labels, _normalize_kind('G.N0', labels, True), 'relu')
    output = _conv2d.Conv2D(output, output.shape[-1], 1024, 3, 1, 'G.Conv', he_init=True, biases=True)

    def block(x, out_dim, name, resample):
        return ResidualBlock(x, x.shape[-1], out_dim, 3, name, inputs_norm=inputs_norm, resample=resample, labels=labels)

    for i in range(bc - 1):
        output = block(output, get_dim(i), 'G.UpBlock.{}'.format(i + 1), 'up')
    if trans:
        toRGB1 = block(output, get_dim(bc - 1), 'G.UpBlock.{}'.format(bc), 'up')
        toRGB1 = block(toRGB1, get_dim(bc - 1), 'G.{}_toRGB1'.format(bc), None)
        # tf.image.resize_nearest_neighbor to toRGB1's size = nearest 2x
        toRGB2 = F.upsample2(F.as_var(output))
        toRGB2 = block(toRGB2, get_dim(bc - 1), 'G.{}_toRGB2'.format(bc), None)
        toRGB = F.lerp(toRGB2, toRGB1, alpha)       # fade in
    else:
        toRGB = block(output, get_dim(bc - 1), 'G.UpBlock.{}'.format(bc), 'up') if bc > 0 else output
        toRGB = block(toRGB, get_dim(bc - 1), 'G.{}_toRGB'.format(bc), None)
    toRGB = F.as_var(toRGB)
    if toRGB.data.dtype == F32:
        toRGB = F.cast(toRGB, BF16)       # input of a batch-statistics normalisation (DESIGN.md "Data layout")
    output, _ = _norm_act('G.Output_Normalize', toRGB, labels, _normalize_kind('G.Output_Normalize', labels, True),
                          'relu')
    output = _conv2d.Conv2D(output, output.shape[-1], 3, 3, 1, 'G.Output', he_init=False)
    return F.activation(output, 'tanh')


def Discriminator_PGGAN(x_var, c_var, bc, trans=False, alpha=0.01, inputs_norm=False, labels=None,
                        update_collection=None, reuse=False):
    """common/resnet_block.py:266-349 -- the ResNet PGGAN critic: a fromRGB residual block (3 -> get_dim(bc-1), no
    resampling), bc 'down' blocks, D.NoneBlock, relu, spatial mean, spectrally-normalised Linear.  With `trans` the
    skip path runs a second fromRGB block over the image resized (nearest) to half its size and the two FEATURE maps
    are blended (:294).  Every layer is spectrally normalised, so Normalize is the identity (:34-36).  c_var is unused
    by the reference as well."""
    del c_var, labels
    kw = dict(spectral_normed=True, update_collection=update_collection, inputs_norm=inputs_norm, biases=True)
    x_var = F.as_var(x_var)
    if trans:
        fromRGB1 = ResidualBlock(x_var, 3, get_dim(bc - 1), 3, 'D.{}_fromRGB1'.format(bc), resample=None, **kw)
        fromRGB1 = ResidualBlock(fromRGB1, get_dim(bc - 1), get_dim(bc - 1), 3, 'D.DownBlock.{}'.format(bc),
                                 resample='down', **kw)
        fromRGB2 = F.subsample2(x_var)         # tf.image.resize_nearest_neighbor to fromRGB1's size: x[:, ::2, ::2]
        fromRGB2 = ResidualBlock(fromRGB2, 3, get_dim(bc - 1), 3, 'D.{}_fromRGB2'.format(bc), resample=None, **kw)
        x_code = F.lerp(fromRGB2, fromRGB1, alpha)
    else:
        x_code = ResidualBlock(x_var, 3, get_dim(bc - 1), 3, 'D.{}_fromRGB'.format(bc), resample=None, **kw)
        if bc > 0:
            x_code = ResidualBlock(x_code, get_dim(bc - 1), get_dim(bc - 1), 3, 'D.DownBlock.{}'.format(bc),
                                   resample='down', **kw)
    for i in range(1, bc):
        x_code = ResidualBlock(x_code, x_code.shape[-1], get_dim(bc - 1 - i), 3, 'D.DownBlock.{}'.format(bc - i),
                               resample='down', **kw)
    output = ResidualBlock(x_code, x_code.shape[-1], get_dim(0), 3, 'D.NoneBlock', resample=None, **kw)
    output = F.act_mean_hw(output, 'relu')      # nonlinearity + tf.reduce_mean(axis=[1, 2])
    logits = _linear.Linear(output, output.shape[-1], 1, 'D.Output', spectral_normed=True,
                            update_collection=update_collection, inputs_norm=inputs_norm, biases=True,
                            initialization=None)
    return F.reshape(logits, (-1,))
