"""The WGAN-GP term of ACGAN/train.py:97-105 for the ACGAN discriminator (ACGAN/model.py:59-90):

    alpha ~ U(0, 1) per sample;  interpolates = real + alpha * (x_fake - real)
    gradients = tf.gradients(D(interpolates, real_labels, 'NO_OPS', reuse=True)[0], [interpolates])[0]
    slopes = sqrt(sum_{hwc} gradients^2 + 1e-10);  gradient_penalty = 10 * mean((slopes - 1)^2)

D is run a third time on the interpolates, layer by layer with fp32 activations kept as tape values, then its input
gradient is built by hand from the differentiable "*_input_grad" ops of functional.py (the backward pass as a forward
graph), so that Tape.backward() of the penalty reaches D's parameters through both passes -- including the grad-grad
of the six batch norms.  Variables are found by their reference names under d_net/ (reuse)."""
from __future__ import annotations

import torch

from .. import functional as F
from .. import kernels as K
from ..framework import Var, get_store

F32 = torch.float32
EPS = 1e-5


def _v(name):
    return get_store().vars['d_net/' + name]


def _conv(x: Var, name: str, k: int) -> Var:
    return F.conv2d(x, _v(name + '/Filters'), _v(name + '/Biases'), k, k, 1, "SAME")


def _bn_lrelu(x: Var, name: str):
    """fp32 batch norm + leaky relu of the block's Normalize + nonlinearity; returns (activation, mean, rstd)."""
    n, h, w, c = x.shape
    gamma, beta = _v(name + '/BatchNorm/gamma'), _v(name + '/BatchNorm/beta')
    mean, rstd = K.bn_stats(x.data, n, h * w, c, 1, EPS)
    a, _ = F.norm_act(x, stats="batch", eps=EPS, gamma=gamma, beta=beta, act='lrelu', out_dtype=torch.bfloat16, groups=1)
    return a, (gamma, beta, mean, rstd)


def discriminator_input_gradient(x_hat: torch.Tensor) -> Var:
    """g = d sum_n D(x_hat)[0] / d x_hat as a tape value (shape of x_hat)."""
    x0 = Var(x_hat)
    n = x0.shape[0]
    # ---------------------------------------------------------------- forward (ACGAN/model.py:69-87, fp32 tape values)
    p = 'D.DownBlock.1'
    h = _conv(x0, p + '.Conv1', 3)
    a = F.activation(h, 'lrelu')
    h2 = _conv(a, p + '.Conv2', 3)
    s0 = _conv(F.meanpool2(x0), p + '.Shortcut', 1)
    o1 = F.add(s0, F.meanpool2(h2))
    blocks = []
    o = o1
    for name, down in (('D.DownBlock.2', True), ('D.NoneBlock.3', False), ('D.NoneBlock.4', False)):
        a1, n1 = _bn_lrelu(o, name + '.N1')
        h1 = _conv(a1, name + '.Conv1', 3)
        a2, n2 = _bn_lrelu(h1, name + '.N2')
        c2 = _conv(a2, name + '.Conv2', 3)
        if down:
            sc = _conv(o, name + '.Shortcut', 1)
            out = F.add(F.meanpool2(sc), F.meanpool2(c2))
        else:
            out = F.add(o, c2)
        blocks.append((name, down, o, n1, h1, n2))
        o = out
    w_out = _v('D.Output/W')
    # ---------------------------------------------------------------- backward pass as a forward graph
    ones = torch.ones((n, 1), dtype=F32, device=x_hat.device)
    g_feat = F.linear_input_grad(ones, w_out)                                   # [n, 128]
    g = F.act_mean_hw_input_grad(o, g_feat, 'lrelu')                            # gradient w.r.t. the last block's output
    for name, down, x_in, n1, h1, n2 in reversed(blocks):
        if down:
            g_c2 = F.meanpool2_input_grad(g)
            g_sc = F.conv2d_input_grad(F.meanpool2_input_grad(g), _v(name + '.Shortcut/Filters'), x_in.shape, 1, 1)
        else:
            g_c2, g_sc = g, g
        g_a2 = F.conv2d_input_grad(g_c2, _v(name + '.Conv2/Filters'), h1.shape, 3, 3)
        g_h1 = F.bn_act_input_grad(h1, g_a2, *n2, 'lrelu')
        g_a1 = F.conv2d_input_grad(g_h1, _v(name + '.Conv1/Filters'), x_in.shape, 3, 3)
        g_main = F.bn_act_input_grad(x_in, g_a1, *n1, 'lrelu')
        g = F.add(g_main, g_sc)
    # first block: o1 = Shortcut(meanpool(x)) + meanpool(Conv2(lrelu(Conv1(x))))
    g_a = F.conv2d_input_grad(F.meanpool2_input_grad(g), _v(p + '.Conv2/Filters'), a.shape, 3, 3)
    g_h = F.act_input_grad(h, g_a, 'lrelu')
    g_x_main = F.conv2d_input_grad(g_h, _v(p + '.Conv1/Filters'), x0.shape, 3, 3)
    g_s = F.conv2d_input_grad(g, _v(p + '.Shortcut/Filters'), (n, x0.shape[1] // 2, x0.shape[2] // 2, x0.shape[3]), 1, 1)
    return F.add(g_x_main, F.meanpool2_input_grad(g_s))


def gradient_penalty(real: torch.Tensor, fake: torch.Tensor, alpha: torch.Tensor, scale: float = 10.0) -> Var:
    """ACGAN/train.py:97-104.  real / fake: NHWC fp32 [n, 32, 32, 3]; alpha: fp32 [n] (the tf.random_uniform draw)."""
    x_hat = K.interpolate(real, fake, alpha)
    return F.gradient_penalty_loss(discriminator_input_gradient(x_hat), scale)
