"""Oracle restatement of common/resnet_block.py:24-184 (and the private copy in
SNGAN/gan_cifar_resnet.py:80-234, which differs only in Normalize and in dropping inputs_norm)."""
from __future__ import annotations

import functools

import torch

from . import ops

NORMALIZATION_G = True   # common/resnet_block.py:20
NORMALIZATION_D = True   # common/resnet_block.py:21


def nonlinearity(x, activation_fn="relu", leakiness=0.2):
    """common/resnet_block.py:24-29.  relu: slope 0 at 0; maximum(x, 0.2x): slope 1 at 0 (TF MaximumGrad)."""
    if activation_fn == "relu":
        return torch.relu(x)
    if activation_fn == "lrelu":
        assert 0 < leakiness <= 1, "leakiness must be <= 1"
        return torch.where(x >= 0, x, leakiness * x)
    raise ValueError("unknown activation %r (the reference silently returns None)" % (activation_fn,))


def Normalize(g, name, inputs, labels=None, spectral_normed=True):
    """common/resnet_block.py:32-50: dispatch on the substring of the layer name."""
    with g.variable_scope(name):
        if ("D." in name) and NORMALIZATION_D:
            if spectral_normed:
                return inputs
            return ops.batch_norm(g, inputs, fused=True)
        elif ("G." in name) and NORMALIZATION_G:
            if labels is not None:
                return ops.cond_batchnorm(g, name, [0, 1, 2], inputs, labels=labels, n_labels=10)
            return ops.batch_norm(g, inputs, fused=True)
        else:
            return inputs


def mean_pool2(x):
    """tf.add_n of the four strided slices / 4 (common/resnet_block.py:62-63, 71-72)."""
    return (x[:, ::2, ::2, :] + x[:, 1::2, ::2, :] + x[:, ::2, 1::2, :] + x[:, 1::2, 1::2, :]) / 4.0


def upsample2(x):
    """tf.depth_to_space(tf.concat([x,x,x,x], 3), 2) == nearest-neighbour 2x (common/resnet_block.py:87-88)."""
    n, h, w, c = x.shape
    cat = torch.cat([x, x, x, x], dim=3)                 # [n,h,w,4c], block (i,j) at channels (2i+j)c..
    y = cat.reshape(n, h, w, 2, 2, c).permute(0, 1, 3, 2, 4, 5)
    return y.reshape(n, 2 * h, 2 * w, c)


def ConvMeanPool(g, inputs, output_dim, filter_size=3, stride=1, name=None, spectral_normed=False,
                 update_collection=None, inputs_norm=False, he_init=True, biases=True):
    """common/resnet_block.py:53-64"""
    output = ops.Conv2D(g, inputs, inputs.shape[-1], output_dim, filter_size, stride, name,
                        spectral_normed=spectral_normed, update_collection=update_collection,
                        inputs_norm=inputs_norm, he_init=he_init, biases=biases)
    return mean_pool2(output)


def MeanPoolConv(g, inputs, output_dim, filter_size=3, stride=1, name=None, spectral_normed=False,
                 update_collection=None, inputs_norm=False, he_init=True, biases=True):
    """common/resnet_block.py:67-80"""
    output = mean_pool2(inputs)
    return ops.Conv2D(g, output, output.shape[-1], output_dim, filter_size, stride, name,
                      spectral_normed=spectral_normed, update_collection=update_collection,
                      inputs_norm=inputs_norm, he_init=he_init, biases=biases)


# Set by tests to the product's rule (functional.upconv_eligible) so that the bf16-operand oracle rounds at the same
# points as the product: (n, h, w, cin, cout, k) -> bool.  None: UpsampleConv is always upsample + conv.
SUBPIXEL_RULE = None


def UpsampleConv(g, inputs, output_dim, filter_size=3, stride=1, name=None, spectral_normed=False,
                 update_collection=None, inputs_norm=False, he_init=True, biases=True):
    """common/resnet_block.py:83-97"""
    n, h, w, c = inputs.shape
    if (SUBPIXEL_RULE is not None and not spectral_normed and not inputs_norm and stride == 1
            and SUBPIXEL_RULE(n, h, w, c, output_dim, filter_size)):
        return ops.Conv2D(g, inputs, c, output_dim, filter_size, stride, name, spectral_normed=False,
                          update_collection=update_collection, inputs_norm=False, he_init=he_init, biases=biases,
                          subpixel_up2=True)
    output = upsample2(inputs)
    return ops.Conv2D(g, output, output.shape[-1], output_dim, filter_size, stride, name,
                      spectral_normed=spectral_normed, update_collection=update_collection,
                      inputs_norm=inputs_norm, he_init=he_init, biases=biases)


def ResidualBlock(g, inputs, input_dim, output_dim, filter_size, name, spectral_normed=False,
                  update_collection=None, inputs_norm=False, resample=None, labels=None, biases=True,
                  activation_fn="relu", normalize=None):
    """common/resnet_block.py:100-156.  `normalize` lets the SNGAN script's own Normalize be plugged in."""
    norm = normalize or (lambda nm, x, labels=None: Normalize(g, nm, x, labels=labels, spectral_normed=spectral_normed))
    conv = functools.partial(ops.Conv2D, g)
    if resample == "down":
        conv_1 = functools.partial(conv, input_dim=input_dim, output_dim=input_dim)
        conv_2 = functools.partial(ConvMeanPool, g, output_dim=output_dim)
        conv_shortcut = functools.partial(ConvMeanPool, g)
    elif resample == "up":
        conv_1 = functools.partial(UpsampleConv, g, output_dim=output_dim)
        conv_shortcut = functools.partial(UpsampleConv, g)
        conv_2 = functools.partial(conv, input_dim=output_dim, output_dim=output_dim)
    elif resample is None:
        conv_shortcut = functools.partial(conv, input_dim=input_dim)
        conv_1 = functools.partial(conv, input_dim=input_dim, output_dim=output_dim)
        conv_2 = functools.partial(conv, input_dim=output_dim, output_dim=output_dim)
    else:
        raise Exception("invalid resample value")

    if output_dim == input_dim and resample is None:
        shortcut = inputs
    else:
        shortcut = conv_shortcut(inputs=inputs, output_dim=output_dim, filter_size=1, name=name + ".Shortcut",
                                 spectral_normed=spectral_normed, update_collection=update_collection,
                                 inputs_norm=inputs_norm, he_init=False, biases=biases)
    output = inputs
    output = norm(name + ".N1", output, labels=labels)
    output = nonlinearity(output, activation_fn=activation_fn)
    output = conv_1(inputs=output, filter_size=filter_size, name=name + ".Conv1", spectral_normed=spectral_normed,
                    update_collection=update_collection, inputs_norm=inputs_norm, he_init=True, biases=biases)
    output = norm(name + ".N2", output, labels=labels)
    output = nonlinearity(output, activation_fn=activation_fn)
    output = conv_2(inputs=output, filter_size=filter_size, name=name + ".Conv2", spectral_normed=spectral_normed,
                    update_collection=update_collection, inputs_norm=inputs_norm, he_init=True, biases=biases)
    return shortcut + output


def OptimizedResBlockDisc1(g, inputs, DIM_D=128, activation_fn="relu", spectral_normed=False,
                           update_collection=None, inputs_norm=False, biases=True, prefix="D.DownBlock.1"):
    """common/resnet_block.py:159-184 (names 'D.DownBlock.1.*'); the SNGAN script's copy
    (gan_cifar_resnet.py:212-234) uses prefix 'D.Block.1'."""
    conv_1 = functools.partial(ops.Conv2D, g, input_dim=inputs.shape[-1], output_dim=DIM_D)
    conv_2 = functools.partial(ConvMeanPool, g, output_dim=DIM_D)
    shortcut = MeanPoolConv(g, inputs=inputs, output_dim=DIM_D, filter_size=1, name=prefix + ".Shortcut",
                            spectral_normed=spectral_normed, update_collection=update_collection,
                            inputs_norm=inputs_norm, he_init=False, biases=biases)
    output = conv_1(inputs=inputs, filter_size=3, name=prefix + ".Conv1", spectral_normed=spectral_normed,
                    update_collection=update_collection, inputs_norm=inputs_norm, he_init=True, biases=biases)
    output = nonlinearity(output, activation_fn=activation_fn)
    output = conv_2(inputs=output, filter_size=3, name=prefix + ".Conv2", spectral_normed=spectral_normed,
                    update_collection=update_collection, inputs_norm=inputs_norm, he_init=True, biases=biases)
    return shortcut + output
