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


# ---------------------------------------------------------------------------------------------- ResNet PGGAN
def get_dim(stage):
    """common/resnet_block.py:188-189 (a float under Python 3 in the reference; used as a channel count)."""
    return int(min(2048 / (2 ** stage), 512))


def resize_nearest(x, out_h, out_w):
    """tf.image.resize_nearest_neighbor(align_corners=False): source index = floor(i * in / out)."""
    n, h, w, c = x.shape
    ii = torch.div(torch.arange(out_h) * h, out_h, rounding_mode="floor")
    jj = torch.div(torch.arange(out_w) * w, out_w, rounding_mode="floor")
    return x[:, ii][:, :, jj]


def Generator_PGGAN(g, noise, bc, trans=False, alpha=0.01, inputs_norm=False, labels=None, training=True):
    """common/resnet_block.py:192-263"""
    output = ops.Linear(g, noise, noise.shape[-1], 4 * 4 * 1024, "G.Input", inputs_norm=inputs_norm, biases=True,
                        initialization=None)                                                        # :206-207
    output = output.reshape(-1, 4, 4, 1024)                                                         # :208
    output = Normalize(g, "G.N0", output, labels=labels, spectral_normed=True)                      # :211
    output = nonlinearity(output, activation_fn="relu")                                             # :212
    output = ops.Conv2D(g, output, output.shape[-1], 1024, 3, 1, "G.Conv", he_init=True, biases=True)   # :215-216

    def block(x, out_dim, name, resample):
        return ResidualBlock(g, x, x.shape[-1], out_dim, 3, name, inputs_norm=inputs_norm, resample=resample,
                             labels=labels)

    for i in range(bc - 1):                                                                         # :222-225
        output = block(output, get_dim(i), "G.UpBlock.{}".format(i + 1), "up")
    if trans:                                                                                       # :227-241
        toRGB1 = block(output, get_dim(bc - 1), "G.UpBlock.{}".format(bc), "up")
        toRGB1 = block(toRGB1, get_dim(bc - 1), "G.{}_toRGB1".format(bc), None)
        toRGB2 = resize_nearest(output, toRGB1.shape[1], toRGB1.shape[2])
        toRGB2 = block(toRGB2, get_dim(bc - 1), "G.{}_toRGB2".format(bc), None)
        toRGB = (1.0 - alpha) * toRGB2 + alpha * toRGB1
    else:                                                                                           # :242-250
        toRGB = block(output, get_dim(bc - 1), "G.UpBlock.{}".format(bc), "up") if bc > 0 else output
        toRGB = block(toRGB, get_dim(bc - 1), "G.{}_toRGB".format(bc), None)
    output = Normalize(g, "G.Output_Normalize", toRGB, labels=labels, spectral_normed=True)         # :253
    output = nonlinearity(output, activation_fn="relu")
    output = ops.Conv2D(g, output, output.shape[-1], 3, 3, 1, "G.Output", he_init=False)            # :255
    return torch.tanh(output)                                                                       # :259


def Discriminator_PGGAN(g, x_var, c_var, bc, trans=False, alpha=0.01, inputs_norm=False, labels=None,
                        update_collection=None, reuse=False):
    """common/resnet_block.py:266-349"""
    kw = dict(spectral_normed=True, update_collection=update_collection, inputs_norm=inputs_norm, biases=True)
    if trans:                                                                                       # :282-296
        fromRGB1 = ResidualBlock(g, x_var, 3, get_dim(bc - 1), 3, "D.{}_fromRGB1".format(bc), resample=None, **kw)
        fromRGB1 = ResidualBlock(g, fromRGB1, get_dim(bc - 1), get_dim(bc - 1), 3, "D.DownBlock.{}".format(bc),
                                 resample="down", **kw)
        fromRGB2 = resize_nearest(x_var, fromRGB1.shape[1], fromRGB1.shape[2])
        fromRGB2 = ResidualBlock(g, fromRGB2, 3, get_dim(bc - 1), 3, "D.{}_fromRGB2".format(bc), resample=None, **kw)
        x_code = (1.0 - alpha) * fromRGB2 + alpha * fromRGB1
    else:                                                                                           # :297-311
        x_code = ResidualBlock(g, x_var, 3, get_dim(bc - 1), 3, "D.{}_fromRGB".format(bc), resample=None, **kw)
        if bc > 0:
            x_code = ResidualBlock(g, x_code, get_dim(bc - 1), get_dim(bc - 1), 3, "D.DownBlock.{}".format(bc),
                                   resample="down", **kw)
    for i in range(1, bc):                                                                          # :313-320
        x_code = ResidualBlock(g, x_code, x_code.shape[-1], get_dim(bc - 1 - i), 3, "D.DownBlock.{}".format(bc - i),
                               resample="down", **kw)
    output = ResidualBlock(g, x_code, x_code.shape[-1], get_dim(0), 3, "D.NoneBlock", resample=None, **kw)  # :322-328
    output = nonlinearity(output, activation_fn="relu")
    output = output.mean(dim=(1, 2))                                                                # :332
    logits = ops.Linear(g, output, output.shape[-1], 1, "D.Output", spectral_normed=True,
                        update_collection=update_collection, inputs_norm=inputs_norm, biases=True,
                        initialization=None)                                                        # :333-337
    return logits.reshape(-1)
