"""Oracle restatement of the reference layer ops (common/ops/*.py) with torch CPU tensors.

Tensors are NHWC like the reference (conv2d.py:186); filters HWIO (conv2d.py:111); linear weights
[in,out] (linear.py:67).  `g` is an oracle.tfshim.Graph that plays the role of the TF default graph.
TF-1.x library semantics that the repository does not show are encoded here once:

  conv2d SAME padding ....... out=ceil(in/s), pad_total=max((out-1)s+k-in,0), before=total//2, rest after
  tf.nn.moments ............. population (biased) variance
  tf.nn.batch_normalization . inv = rsqrt(var+eps)*gamma ; y = x*inv + (beta - mean*inv)
  depth_to_space(concat x4) . nearest-neighbour 2x upsample
  relu / maximum(x,0.2x) .... slope 0 at exactly 0 for relu; slope 1 at exactly 0 for the leaky form
"""
from __future__ import annotations

import warnings

import numpy as np
import torch
import torch.nn.functional as F

from . import tfshim

NO_OPS = "NO_OPS"

# When True the oracle rounds the operands of every convolution / large matmul to bf16 at exactly the points
# where the B200 kernels do (activations, raw filters, output gradients; fp32 accumulation everywhere).  This
# separates "is the implementation right" (bf16-operand oracle vs product: ~1e-5) from "how far is bf16 from
# fp32" (fp32 oracle vs product: the <=1e-2 per-layer tolerance of the north star).
BF16_OPERANDS = False


def _r16(t):
    return t.to(torch.bfloat16).to(t.dtype)


def _ste_r16(t):
    """bf16 rounding with a straight-through gradient."""
    return t + (_r16(t) - t).detach()


class _ConvRoundedOperands(torch.autograd.Function):
    """y = conv(r16(x), w): forward rounds the activation; backward rounds dy and uses the rounded x:
    dw = wgrad(r16(x), r16(dy)), dx = dgrad(r16(dy), w) -- the operand roundings of conv_tc.cu."""

    @staticmethod
    def forward(ctx, x, w, stride, padding, scale=1.0):
        """scale: a constant the product applies in the GEMM epilogue (inputs_norm), i.e. AFTER the rounded
        contraction in the forward pass and after the rounded-gradient contractions in the backward pass."""
        xr = _r16(x)
        ctx.save_for_backward(xr, w)
        ctx.cfg = (stride, padding, scale)
        y = conv2d_nhwc(xr, w, stride, padding)
        return y if scale == 1.0 else y * scale

    @staticmethod
    def backward(ctx, gy):
        """Written with differentiable ops (no once_differentiable): the WGAN-GP oracle differentiates THROUGH this
        backward pass (torch.autograd.grad(..., create_graph=True))."""
        xr, w = ctx.saved_tensors
        stride, padding, scale = ctx.cfg
        gyr = _ste_r16(gy)
        dx, dw = _conv_grads_nhwc(xr, w, gyr, stride, padding)
        if scale != 1.0:
            dx, dw = dx * scale, dw * scale
        return dx, dw, None, None, None


def _conv_grads_nhwc(x, w, gy, stride, padding):
    """(d/dx, d/dw) of sum(gy * conv2d_nhwc(x, w)) as differentiable expressions (TF padding handled explicitly)."""
    kh, kw = w.shape[0], w.shape[1]
    if padding == "SAME":
        pt, pb, _ = same_pads(x.shape[1], kh, stride)
        pl, pr, _ = same_pads(x.shape[2], kw, stride)
    elif isinstance(padding, (tuple, list)):
        pt, pb, pl, pr = padding
    else:
        pt = pb = pl = pr = 0
    xc = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    wc = w.permute(3, 2, 0, 1)
    gc = gy.permute(0, 3, 1, 2)
    dxp = torch.nn.grad.conv2d_input(xc.shape, wc, gc, stride=stride)
    dwc = torch.nn.grad.conv2d_weight(xc, wc.shape, gc, stride=stride)
    dx = dxp[:, :, pt:dxp.shape[2] - pb, pl:dxp.shape[3] - pr].permute(0, 2, 3, 1)
    return dx, dwc.permute(2, 3, 1, 0)

# module-level switches of conv2d.py:10-28 / linear.py:12-35 / deconv2d.py:8-26
_default_weightnorm = False
_weights_stdev = None


def enable_default_weightnorm():
    global _default_weightnorm
    _default_weightnorm = True


def disable_default_weightnorm():
    global _default_weightnorm
    _default_weightnorm = False


def set_weights_stdev(weights_stdev):
    global _weights_stdev
    _weights_stdev = weights_stdev


def unset_weights_stdev():
    global _weights_stdev
    _weights_stdev = None


# ---------------------------------------------------------------------------------------------- helpers
def same_pads(in_size: int, k: int, s: int):
    """TF 'SAME' padding for one spatial dim -> (before, after, out)."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    before = total // 2
    return before, total - before, out


def conv2d_nhwc(x, filt, stride=1, padding="SAME"):
    """tf.nn.conv2d(NHWC, HWIO): cross-correlation, TF padding rules (conv2d.py:181-187)."""
    kh, kw = filt.shape[0], filt.shape[1]
    xc = x.permute(0, 3, 1, 2)
    if padding == "SAME":
        pt, pb, _ = same_pads(x.shape[1], kh, stride)
        pl, pr, _ = same_pads(x.shape[2], kw, stride)
        xc = F.pad(xc, (pl, pr, pt, pb))
    elif isinstance(padding, (tuple, list)):      # explicit (top, bottom, left, right) zero padding
        pt, pb, pl, pr = padding
        xc = F.pad(xc, (pl, pr, pt, pb))
    elif padding != "VALID":
        raise ValueError(padding)
    y = F.conv2d(xc, filt.permute(3, 2, 0, 1), stride=stride)
    return y.permute(0, 2, 3, 1)


def depthwise_conv2d_nhwc(x, filt, stride=1, padding="SAME"):
    """tf.nn.depthwise_conv2d(NHWC, [kh, kw, c, cm]): output channel ci*cm + m = the grouped convolution with one group
    per input channel (conv2d.py:188-197)."""
    kh, kw, c, cm = filt.shape
    xc = x.permute(0, 3, 1, 2)
    if padding == "SAME":
        pt, pb, _ = same_pads(x.shape[1], kh, stride)
        pl, pr, _ = same_pads(x.shape[2], kw, stride)
        xc = F.pad(xc, (pl, pr, pt, pb))
    elif isinstance(padding, (tuple, list)):
        pt, pb, pl, pr = padding
        xc = F.pad(xc, (pl, pr, pt, pb))
    elif padding != "VALID":
        raise ValueError(padding)
    wg = filt.permute(2, 3, 0, 1).reshape(c * cm, 1, kh, kw)          # torch grouped layout [c*cm, 1, kh, kw]
    y = F.conv2d(xc, wg, stride=stride, groups=c)
    return y.permute(0, 2, 3, 1)


def subpixel_upconv_rounded(x_low, w, r16_filters=True):
    """conv3x3_SAME(nearest2x(x_low), w) evaluated the way the B200 path does (ganb_upconv_*): four 2x2 convolutions
    over the low-resolution tensor with effective filters E_ij[p][q] = sum_{r in R_i[p], s in R_j[q]} w[r][s],
    R_0 = ({0}, {1,2}), R_1 = ({0,1}, {2}); identical to the reference formula in exact arithmetic -- the bf16 rounding
    point moves from w to E (used only with BF16_OPERANDS)."""
    n, h, wd, _ = x_low.shape
    cout = w.shape[-1]
    rows = {0: ([0], [1, 2]), 1: ([0, 1], [2])}
    out = x_low.new_zeros((n, h, 2, wd, 2, cout))
    for i in (0, 1):
        for j in (0, 1):
            e = torch.stack([torch.stack([sum(w[r, s] for r in rows[i][p_] for s in rows[j][q_]) for q_ in (0, 1)])
                             for p_ in (0, 1)])                                 # [2, 2, cin, cout]
            eq = _ste_r16(e) if r16_filters else e
            y = _ConvRoundedOperands.apply(x_low, eq, 1, (1 - i, i, 1 - j, j))
            out[:, :, i, :, j, :] = y
    return out.reshape(n, 2 * h, 2 * wd, cout)


def conv2d_transpose_nhwc(x, filt, stride=2, padding="SAME"):
    """tf.nn.conv2d_transpose with output [N, 2H, 2W, Cout] (deconv2d.py:99-109).

    filt is [k, k, Cout, Cin]; the op is the input-gradient of conv2d([N,2H,2W,Cout] -> [N,H,W,Cin])."""
    k = filt.shape[0]
    n, h, w, _ = x.shape
    oh, ow = 2 * h, 2 * w
    full = F.conv_transpose2d(x.permute(0, 3, 1, 2), filt.permute(3, 2, 0, 1), stride=stride)
    if padding == "SAME":
        pt, _, _ = same_pads(oh, k, stride)
        pl, _, _ = same_pads(ow, k, stride)
    else:
        pt = pl = 0
    fh, fw = full.shape[2], full.shape[3]
    # the forward conv may leave trailing rows unused; pad so the crop below is always in range
    full = F.pad(full, (0, max(0, pl + ow - fw), 0, max(0, pt + oh - fh)))
    return full[:, :, pt:pt + oh, pl:pl + ow].permute(0, 2, 3, 1)


def _memo(fn):
    """Evaluates fn at most once (initial values are drawn lazily, see tfshim.Graph.draw_on_reuse)."""
    box = []

    def get():
        if not box:
            box.append(fn())
        return box[0]

    return get


def _uniform(stdev, size):
    """np.random.uniform(+-stdev*sqrt(3)) as float32 -- conv2d.py:83-88, linear.py:53-60."""
    return np.random.uniform(low=-stdev * np.sqrt(3), high=stdev * np.sqrt(3), size=size).astype("float32")


# ---------------------------------------------------------------------------------------------- sn.py
def _l2normalize(v, eps=1e-12):
    """common/ops/sn.py:11-12"""
    return v / (torch.sum(v ** 2) ** 0.5 + eps)


def spectral_normed_weight(g, W, u=None, num_iters=1, update_collection=None, with_sigma=False, reuse=False):
    """common/ops/sn.py:15-69.  No stop_gradient anywhere: gradients flow through the power iteration."""
    with g.variable_scope("spectral_norm"):
        W_shape = list(W.shape)
        W_reshaped = W.reshape(-1, W_shape[-1])                      # sn.py:30
        if u is None:                                                 # sn.py:31-32
            u = g.get_variable("u", initializer=lambda s: tfshim.truncated_normal(s, g.u_rng),
                               shape=[1, W_shape[-1]], trainable=False)
        u_i = u.detach().clone()  # u is non-trainable state; the clone keeps u.assign() out of the tape
        v_i = torch.zeros(1, W_reshaped.shape[0], dtype=W.dtype)
        for _ in range(num_iters):                                    # sn.py:34-47 (tf.while_loop)
            v_i = _l2normalize(u_i @ W_reshaped.t())
            u_i = _l2normalize(v_i @ W_reshaped)
        u_final, v_final = u_i, v_i
        sigma = ((v_final @ W_reshaped) @ u_final.t())[0, 0]          # sn.py:52 / :58
        W_bar = (W_reshaped / sigma).reshape(W_shape)
        if update_collection is None:                                 # sn.py:48-56
            g.assign(u, u_final)
        elif update_collection != NO_OPS:                             # sn.py:64-65
            g.add_to_collection(update_collection, (u, u_final.detach().clone()))
    if with_sigma:
        return W_bar, sigma
    return W_bar


# ---------------------------------------------------------------------------------------------- conv2d.py
def Conv2D(g, inputs, input_dim, output_dim, filter_size=3, stride=1, name="Conv2D", conv_type="conv2d",
           channel_multiplier=0, padding="SAME", spectral_normed=False, update_collection=None,
           inputs_norm=False, he_init=True, mask_type=None, weightnorm=None, biases=True, gain=1.0, reuse=None,
           subpixel_up2=False):
    """common/ops/conv2d.py:31-218 (conv2d_.py adds the ignored `reuse` keyword).
    subpixel_up2 (oracle-only plumbing for UpsampleConv, resnet_block.py:83-97): `inputs` is the tensor BEFORE the
    nearest 2x upsample; the fp32 oracle upsamples and convolves exactly like the reference, the bf16-operand oracle
    mirrors the product's sub-pixel evaluation (subpixel_upconv_rounded)."""
    if conv_type not in ("conv2d", "depthwise_conv2d", "separable_conv2d"):
        raise NotImplementedError("{0} is not supported!".format(conv_type))   # conv2d.py:209-210
    with g.variable_scope(name):
        mask = None
        if mask_type is not None:                                     # conv2d.py:63-81
            kind, mask_n_channels = mask_type
            mask = np.ones((filter_size, filter_size, input_dim, output_dim), dtype="float32")
            center = filter_size // 2
            mask[center + 1:, :, :, :] = 0.0
            mask[center, center + 1:, :, :] = 0.0
            for i in range(mask_n_channels):
                for j in range(mask_n_channels):
                    if (kind == "a" and i >= j) or (kind == "b" and i > j):
                        mask[center, center, i::mask_n_channels, j::mask_n_channels] = 0.0
        fan_in = input_dim * filter_size ** 2                         # conv2d.py:90
        fan_out = output_dim * filter_size ** 2 / (stride ** 2)       # conv2d.py:91
        inv_c = float(np.sqrt(2.0 / fan_in))
        if inputs_norm:                                               # conv2d.py:93-97
            inputs_ = inputs * inv_c
        else:
            inputs_ = inputs
        if mask_type is not None:                                     # conv2d.py:99-101
            fan_in /= 2.0
            fan_out /= 2.0
        if he_init:                                                   # conv2d.py:103-106
            filters_stdev = np.sqrt(4.0 / (fan_in + fan_out))
        else:
            filters_stdev = np.sqrt(2.0 / (fan_in + fan_out))
        stdev = _weights_stdev if _weights_stdev is not None else filters_stdev
        fv = _memo(lambda: _uniform(stdev, (filter_size, filter_size, input_dim, output_dim)) * np.float32(gain))
        filters = g.get_variable("Filters", initializer=lambda _s: fv())   # conv2d.py:124-144
        if channel_multiplier > 0:                                    # conv2d.py:117-126, 145-150 (no gain)
            depthwise_filters = g.get_variable("depthwise_filters", initializer=lambda _s: _uniform(
                stdev, (filter_size, filter_size, input_dim, channel_multiplier)))
            pointwise_filters = g.get_variable("pointwise_filters", initializer=lambda _s: _uniform(
                stdev, (1, 1, input_dim * channel_multiplier, output_dim)))
        if conv_type != "conv2d":
            # weight-norm and the mask act on `Filters` only, which these types never read (conv2d.py:153-167).
            # Spectral norm wraps ALL THREE filters (conv2d.py:169-178), each under its own scope: `filters` (W_bar
            # unused, so in TF its u.assign never runs: the variable exists and keeps its initial value),
            # `depthwise_filters` and `pointwise_filters` -- and the normalised ones are what the op consumes.  A
            # W_bar that no op reads never runs its control-dependent u.assign in TF: evaluated with NO_OPS here.
            pw_raw, pw_sigma = pointwise_filters, None
            if spectral_normed:
                with g.variable_scope("filters"):
                    spectral_normed_weight(g, filters, update_collection=NO_OPS)
                with g.variable_scope("depthwise_filters"):           # conv2d.py:173-175
                    depthwise_filters = spectral_normed_weight(g, depthwise_filters, update_collection=update_collection)
                with g.variable_scope("pointwise_filters"):           # conv2d.py:176-178
                    uc = update_collection if conv_type == "separable_conv2d" else NO_OPS
                    pointwise_filters, pw_sigma = spectral_normed_weight(g, pointwise_filters, update_collection=uc,
                                                                         with_sigma=True)
            x_ = _ste_r16(inputs_) if BF16_OPERANDS else inputs_      # the B200 path reads conv operands in bf16
            result = depthwise_conv2d_nhwc(x_, depthwise_filters, stride, padding)        # conv2d.py:188-197
            if conv_type == "separable_conv2d":                       # conv2d.py:198-208: then the 1x1 pointwise conv
                if BF16_OPERANDS:
                    wq = _ste_r16(pw_raw)                             # bf16 operand copy of W, 1/sigma in the epilogue
                    if pw_sigma is not None:
                        wq = wq / pw_sigma
                    result = _ConvRoundedOperands.apply(result, wq, 1, "VALID")
                else:
                    result = conv2d_nhwc(result, pointwise_filters, 1, "VALID")
            if biases:
                b = g.get_variable("Biases", initializer=tfshim.constant_initializer(0.0), shape=[output_dim])
                result = result + b                                   # fails in TF too unless the channels agree
            return result
        if weightnorm is None:
            weightnorm = _default_weightnorm
        if weightnorm:                                                # conv2d.py:153-163
            norm_values = np.sqrt(np.sum(np.square(fv()), axis=(0, 1, 2)))
            target_norms = g.get_variable("g", initializer=norm_values)
            norms = torch.sqrt(torch.sum(filters ** 2, dim=(0, 1, 2)))
            filters = filters * (target_norms / norms)
        if mask is not None:                                          # conv2d.py:165-167
            filters = filters * torch.as_tensor(mask, dtype=filters.dtype)
        raw_filters = filters
        if spectral_normed:                                           # conv2d.py:169-171
            with g.variable_scope("filters"):
                filters, sigma = spectral_normed_weight(g, filters, update_collection=update_collection,
                                                        with_sigma=True)
        if subpixel_up2:
            assert filter_size == 3 and stride == 1 and padding == "SAME" and not spectral_normed and not inputs_norm
            if BF16_OPERANDS:
                result = subpixel_upconv_rounded(inputs_, raw_filters)
            else:
                n_, h_, w_, c_ = inputs_.shape
                up = torch.cat([inputs_] * 4, dim=3).reshape(n_, h_, w_, 2, 2, c_).permute(0, 1, 3, 2, 4, 5)
                result = conv2d_nhwc(up.reshape(n_, 2 * h_, 2 * w_, c_), filters, stride, padding)
        elif BF16_OPERANDS:
            wq = _ste_r16(raw_filters)
            if spectral_normed:
                wq = wq / sigma
            if inputs_norm:
                # the product keeps the constant outside the tensor-core contraction (the epilogue's alpha):
                # c * conv(r16(x), r16(W)) -- same value in exact arithmetic, the rounding point moves
                result = _ConvRoundedOperands.apply(inputs, wq, stride, padding, inv_c)
            else:
                result = _ConvRoundedOperands.apply(inputs_, wq, stride, padding)
        else:
            result = conv2d_nhwc(inputs_, filters, stride, padding)   # conv2d.py:181-187
        if biases:                                                    # conv2d.py:212-216
            b = g.get_variable("Biases", initializer=tfshim.constant_initializer(0.0), shape=[output_dim])
            result = result + b
        return result


# ---------------------------------------------------------------------------------------------- deconv2d.py
def Deconv2D(g, inputs, in_channels, output_channels, filter_size, stride=2, padding="SAME", he_init=True,
             weight_norm=None, gain=1.0, mask_type=None, biases=True, name="Deconv2D"):
    """common/ops/deconv2d.py:29-118"""
    with g.variable_scope(name):
        if mask_type is not None:
            raise Exception("Unsupported configuration in Deconv2D!")
        fan_in = in_channels * filter_size ** 2 / (stride ** 2)       # deconv2d.py:62
        fan_out = output_channels * filter_size ** 2                  # deconv2d.py:63
        if he_init:
            filters_stdev = np.sqrt(4.0 / (fan_in + fan_out))
        else:
            filters_stdev = np.sqrt(2.0 / (fan_in + fan_out))
        stdev = _weights_stdev if _weights_stdev is not None else filters_stdev
        fv = _memo(lambda: _uniform(stdev, (filter_size, filter_size, output_channels, in_channels))
                   * np.float32(gain))
        filters = g.get_variable("Filters", initializer=lambda _s: fv())
        if weight_norm is None:
            weight_norm = _default_weightnorm
        if weight_norm:                                               # deconv2d.py:87-96
            norm_values = np.sqrt(np.sum(np.square(fv()), axis=(0, 1, 3)))
            target_norms = g.get_variable("g", initializer=norm_values)
            norms = torch.sqrt(torch.sum(filters ** 2, dim=(0, 1, 3)))
            filters = filters * (target_norms / norms).unsqueeze(1)
        result = conv2d_transpose_nhwc(inputs, filters, stride, padding)  # deconv2d.py:99-109
        if biases:
            b = g.get_variable("Biases", initializer=tfshim.constant_initializer(0.0), shape=[output_channels])
            result = result + b
        return result


# ---------------------------------------------------------------------------------------------- linear.py
def Linear(g, inputs, input_dim, output_dim, name, spectral_normed=False, update_collection=None, reuse=False,
           inputs_norm=False, biases=True, initialization=None, weightnorm=None, gain=1.0):
    """common/ops/linear.py:38-182"""
    with g.variable_scope(name):
        if inputs_norm:                                               # linear.py:47-51
            inputs_ = inputs * float(np.sqrt(2.0 / input_dim))
        else:
            inputs_ = inputs

        def uniform(stdev, size):                                     # linear.py:53-60
            if _weights_stdev is not None:
                stdev = _weights_stdev
            return _uniform(stdev, size)

        def draw():
            if initialization == "lecun":
                wv = uniform(np.sqrt(1.0 / input_dim), (input_dim, output_dim))
            elif initialization in ("glorot", "xavier") or initialization is None:  # linear.py:76 (None lands here)
                wv = uniform(np.sqrt(2.0 / (input_dim + output_dim)), (input_dim, output_dim))
            elif initialization == "he":
                wv = uniform(np.sqrt(2.0 / input_dim), (input_dim, output_dim))
            elif initialization == "glorot_he":
                wv = uniform(np.sqrt(4.0 / (input_dim + output_dim)), (input_dim, output_dim))
            elif initialization == "orthogonal":                      # linear.py:112-128
                a = np.random.normal(0.0, 1.0, (input_dim, output_dim))
                uu, _, vv = np.linalg.svd(a, full_matrices=False)
                q = uu if uu.shape == (input_dim, output_dim) else vv
                wv = q.reshape((input_dim, output_dim)).astype("float32")
            elif initialization[0] == "uniform":
                wv = np.random.uniform(low=-initialization[1], high=initialization[1],
                                       size=(input_dim, output_dim)).astype("float32")
            else:
                raise Exception("Invalid initialization!")
            return wv * np.float32(gain)                              # linear.py:138

        wv_memo = _memo(draw)
        weight = g.get_variable("W", initializer=lambda _s: wv_memo())
        if weightnorm is None:
            weightnorm = _default_weightnorm
        if weightnorm:                                                # linear.py:143-155
            norm_values = np.sqrt(np.sum(np.square(wv_memo()), axis=0))
            target_norms = g.get_variable("g", initializer=norm_values)
            norms = torch.sqrt(torch.sum(weight ** 2, dim=0))
            weight = weight * (target_norms / norms)
        w_eff = spectral_normed_weight(g, weight, update_collection=update_collection) if spectral_normed else weight
        if (BF16_OPERANDS and inputs_.dim() == 2 and input_dim % 8 == 0 and output_dim % 8 == 0
                and input_dim * output_dim >= 65536 and not spectral_normed):
            # the product runs this layer as a 1x1 convolution on the tensor cores
            x4 = inputs.reshape(-1, 1, 1, input_dim)
            w4 = _ste_r16(weight).reshape(1, 1, input_dim, output_dim)
            # inputs_norm: constant applied as the GEMM's alpha, outside the rounded contraction
            c = float(np.sqrt(2.0 / input_dim)) if inputs_norm else 1.0
            result = _ConvRoundedOperands.apply(x4, w4, 1, "VALID", c).reshape(-1, output_dim)
        elif inputs_.dim() == 2:                                      # linear.py:161-165
            result = inputs_ @ w_eff
        else:                                                         # linear.py:166-174
            result = (inputs_.reshape(-1, input_dim) @ w_eff).reshape(*inputs_.shape[:-1], output_dim)
        if biases:                                                    # linear.py:176-180
            b = g.get_variable("b", initializer=tfshim.constant_initializer(0.0), shape=[output_dim])
            result = result + b
        return result


# ---------------------------------------------------------------------------------------------- normalization.py
def batch_norm(g, inputs, decay=0.9, epsilon=1e-5, is_training=True, fused=True):
    """common/ops/normalization.py:8-24: contrib fused BN, always training mode -> batch statistics.

    Moving statistics are write-only state in every shipped caller; they are updated here with the plain
    EMA (population mean, Bessel-corrected variance) and no zero-debias slots."""
    with g.variable_scope("BatchNorm"):
        if BF16_OPERANDS:
            inputs = _ste_r16(inputs)   # the B200 path stores every batch-norm input in bf16
        c = inputs.shape[-1]
        beta = g.get_variable("beta", initializer=tfshim.constant_initializer(0.0), shape=[c])
        gamma = g.get_variable("gamma", initializer=tfshim.constant_initializer(1.0), shape=[c])
        mm = g.get_variable("moving_mean", initializer=tfshim.constant_initializer(0.0), shape=[c], trainable=False)
        mv = g.get_variable("moving_variance", initializer=tfshim.constant_initializer(1.0), shape=[c],
                            trainable=False)
        red = tuple(range(inputs.dim() - 1))
        mean = inputs.mean(dim=red)
        var = inputs.var(dim=red, unbiased=False)
        y = (inputs - mean) * torch.rsqrt(var + epsilon) * gamma + beta
        if is_training:
            cnt = inputs.numel() // c
            with torch.no_grad():
                mm.mul_(decay).add_((1 - decay) * mean.detach())
                mv.mul_(decay).add_((1 - decay) * var.detach() * (cnt / max(cnt - 1, 1)))
        return y


def cond_batchnorm(g, name, axes, inputs, is_training=None, stats_iter=None, update_moving_stats=True, fused=True,
                   labels=None, n_labels=None):
    """common/ops/normalization.py:27-59"""
    with g.variable_scope("CondBatchNorm"):
        if axes != [0, 1, 2]:
            raise Exception("Axes is not supported in Conditional BatchNorm!")
        if BF16_OPERANDS:
            inputs = _ste_r16(inputs)   # the B200 path stores every batch-norm input in bf16
        mean = inputs.mean(dim=(0, 1, 2), keepdim=True)               # tf.nn.moments -> population variance
        var = inputs.var(dim=(0, 1, 2), unbiased=False, keepdim=True)
        c = inputs.shape[3]
        offset_m = g.get_variable("offset", initializer=tfshim.constant_initializer(0.0), shape=[n_labels, c])
        scale_m = g.get_variable("scale", initializer=tfshim.constant_initializer(1.0), shape=[n_labels, c])
        offset = offset_m[labels.long()]                              # embedding_lookup, normalization.py:54-55
        scale = scale_m[labels.long()]
        inv = torch.rsqrt(var + 1e-5) * scale[:, None, None, :]       # tf.nn.batch_normalization
        return inputs * inv + (offset[:, None, None, :] - mean * inv)


def layer_norm(g, name, norm_axes, inputs):
    """common/ops/normalization.py:62-82: contrib layer_norm, begin_norm_axis=1, begin_params_axis=-1."""
    with g.variable_scope(name):
        c = inputs.shape[-1]
        beta = g.get_variable("beta", initializer=tfshim.constant_initializer(0.0), shape=[c])
        gamma = g.get_variable("gamma", initializer=tfshim.constant_initializer(1.0), shape=[c])
        if BF16_OPERANDS:
            inputs = _ste_r16(inputs)   # the B200 path keeps Conv1's output (the input of N2) in bf16
        red = tuple(range(1, inputs.dim()))
        mean = inputs.mean(dim=red, keepdim=True)
        var = inputs.var(dim=red, unbiased=False, keepdim=True)
        return (inputs - mean) * torch.rsqrt(var + 1e-12) * gamma + beta


def instance_norm(g, inputs, epsilon=1e-06):
    """common/ops/normalization.py:105-122: per-(n,c) moments over H,W."""
    with g.variable_scope("InstanceNorm"):
        c = inputs.shape[-1]
        beta = g.get_variable("beta", initializer=tfshim.constant_initializer(0.0), shape=[c])
        gamma = g.get_variable("gamma", initializer=tfshim.constant_initializer(1.0), shape=[c])
        mean = inputs.mean(dim=(1, 2), keepdim=True)
        var = inputs.var(dim=(1, 2), unbiased=False, keepdim=True)
        return (inputs - mean) * torch.rsqrt(var + epsilon) * gamma + beta


def pixel_norm(inputs, eps=1e-8):
    """common/ops/normalization.py:125-140"""
    alpha = 1.0 / torch.sqrt(torch.mean(inputs * inputs, dim=3, keepdim=True) + eps)
    return alpha * inputs


# ---------------------------------------------------------------------------------------------- embedding.py
def embed_y(g, inputs, vocab_size=1000, embedding_dim=300, word2vec_file=None, spectral_normed=False,
            update_collection=None, reuse=False):
    """common/ops/embedding.py:12-51"""
    with g.variable_scope("Embedding.Label"):
        if word2vec_file is None:
            embedding_map = g.get_variable(
                "embedding_map", trainable=True,
                initializer=lambda _s: np.random.uniform(low=-0.08, high=0.08,
                                                         size=(vocab_size, embedding_dim)).astype("float32"))
        else:
            embedding_map = g.get_variable("embedding_map", initializer=word2vec_file, trainable=False)
        return embedding_map[inputs.long()]


def _silence():
    warnings.filterwarnings("ignore")
