"""The WGAN-GP term of Pix2Pix/train.py:489-503 for the spectrally-normalised PatchGAN (networks.unet_d /
unet_discriminator):

    alpha ~ U(0, 1) [n, 1, 1, 1];  interpolates = targets + alpha * (outputs - targets)
    gradients = tf.gradients(D(inputs, interpolates, update_collection=None, reuse=True), [interpolates])[0]
    slopes = sqrt(sum_{hwc} gradients^2 + 1e-10);  gradient_penalty = 10 * mean((slopes - 1)^2)

The critic is a chain  conv_1 -> lrelu -> ... -> conv_L  of spectrally-normalised 4x4 convolutions, so its input
gradient  g = T_1(m_1 * T_2(... m_{L-1} * T_L(1)))  is a chain of data-gradient operators T_l (linear in W_l / sigma_l)
and leaky-ReLU masks m_l (piecewise constant).  d(penalty)/dW needs no new kernels:

  1. D(inputs, interpolates) runs on an inner tape with d_net frozen and is back-propagated from a cotangent of ones:
     that IS the chain above; it leaves g (the gradient of the image pair) and gy_l (the gradient at every conv output).
  2. the penalty and c_1 = d(penalty)/dg come from the loss kernel on a second inner tape.
  3. the vector-Jacobian product of the chain is evaluated forwards: y_l = conv_l(c_l) (no bias), c_{l+1} = m_l * y_l,
     and d(penalty)/d(W_l / sigma_l) = wgrad(c_l, gy_l) -- which is exactly what functional.conv2d's own backward
     computes when y_l's gradient is set to gy_l.  The y_l live on the critic-loss tape, so these filter gradients join
     the spectral-norm backward (sigma, u, v of THIS evaluation: the third update_collection=None pass of the step).
"""
from __future__ import annotations

import torch

from .. import functional as F
from .. import kernels as K
from ..common.ops.sn import spectral_normed_weight
from ..framework import Var, get_store

F32 = torch.float32
BF16 = torch.bfloat16
PAD1 = (1, 1, 1, 1)      # tf.pad [[0,0],[1,1],[1,1],[0,0]] in front of every VALID convolution (networks.py:287-354)


def _layers(n_layers: int):
    """(scope, stride) of the PatchGAN convolutions: layer_1, n_layers middle layers (the last with stride 1), the head."""
    strides = [2] + [1 if i == n_layers - 1 else 2 for i in range(n_layers)] + [1]
    return [("layer_%d" % (i + 1), s) for i, s in enumerate(strides)]


def gradient_penalty(inputs: torch.Tensor, targets: torch.Tensor, outputs: torch.Tensor, alpha: torch.Tensor,
                     n_layers: int = 3, scale: float = 10.0) -> Var:
    """Returns the penalty as a loss Var on the recording (critic-loss) tape.  inputs / targets / outputs: NHWC fp32;
    alpha: fp32 [n] (the tf.random_uniform draw).  Must be called after D(real) and D(fake) of the same step
    (reference order of the three update_collection=None evaluations)."""
    st = get_store()
    main = st.tape
    if main is None:
        raise ValueError("gradient_penalty must be built inside the critic's gradient tape")
    interp = K.interpolate(targets.contiguous(), outputs.contiguous(), alpha)
    x_hat = torch.cat([inputs, interp], dim=3).contiguous()           # tensor plumbing: the discriminator's concat
    n = x_hat.shape[0]
    c_in = inputs.shape[3]
    layers = _layers(n_layers)
    # ---- spectral-norm state of this (third) evaluation, acquired on the critic-loss tape: u <- u' once more
    ws, bs, entries = [], [], []
    for scope, _ in layers:
        key = "d_net/%s/Conv2D/" % scope
        w = st.vars[key + "Filters"]
        with st.variable_scope("d_net", reuse=True), st.variable_scope(scope), st.variable_scope("Conv2D"), \
                st.variable_scope("filters"):
            entries.append(spectral_normed_weight(w, update_collection=None).entry)
        ws.append(w)
        bs.append(st.vars[key + "Biases"])
    token, gens = st.tape_token, st._sn_gen
    try:
        # ---- 1. D(x_hat) and its input gradient on an inner tape (parameters frozen: data gradients only)
        xv = Var(x_hat, requires_grad=True, grad_dtype=F32)
        with st.gradient_tape() as t1, st.frozen_scopes("d_net"):
            hs, a = [], xv
            for (scope, stride), w, b, e in zip(layers, ws, bs, entries):
                h = F.conv2d(a, w, b, 4, 4, stride, PAD1, sn=e, out_grad_dtype=F32)
                hs.append(h)
                if len(hs) < len(layers):
                    a, _ = F.norm_act(h, stats=None, act='lrelu', out_dtype=BF16, out_grad_dtype=F32)
            t1.backward(hs[-1], grad=torch.ones_like(hs[-1].data))
        gys = [h.grad if h.grad.dtype == F32 else K.cast(h.grad, F32) for h in hs]
        g_t = xv.grad[..., c_in:].contiguous()                         # gradient w.r.t. the interpolates
        # ---- 2. penalty value and its cotangent c = d(penalty)/dg
        gv = Var(g_t, requires_grad=True, grad_dtype=F32)
        with st.gradient_tape() as t2:
            pen = F.gradient_penalty_loss(gv, scale)
            t2.backward(pen)
        c = torch.zeros_like(x_hat)
        c[..., c_in:].copy_(gv.grad)
    finally:
        st.tape, st.tape_token, st._sn_gen = main, token, gens
    # ---- 3. forward evaluation of the chain's VJP on the critic-loss tape
    ys = []
    cv = Var(c)
    for i, ((scope, stride), w, e) in enumerate(zip(layers, ws, entries)):
        y = F.conv2d(cv, w, None, 4, 4, stride, PAD1, sn=e)
        ys.append(y)
        if i + 1 < len(layers):
            nh, hh, wh, ch = hs[i].shape
            nxt = K.norm_act_bwd(hs[i].data, y.data, 0, nh, hh, wh, ch, None, None, 1, None, None, None, 'lrelu', False,
                                 None, None, None, F32)                 # c_{l+1} = lrelu'(h_l) * conv_l(c_l)
            cv = Var(nxt)
    out = Var(pen.data)
    out.requires_grad = True

    def seed():
        # runs first when the critic-loss tape unwinds (recorded last): the "output gradients" of the y_l are the gy_l,
        # so each conv's own backward emits wgrad(c_l, gy_l) into the spectral-norm G buffer of this evaluation
        for y, gy in zip(ys, gys):
            y.accum(gy)
    main.record(seed)
    main.keep.extend(gys)
    return out
