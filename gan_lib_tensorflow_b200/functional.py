"""Differentiable building blocks over framework.Var: each forward launches libganb200 kernels and records a
backward closure on the current tape.  These are the fused primitives that the reference-compatible layer
functions (common/ops/*.py, common/resnet_block.py) are assembled from.

Dtype convention: residual-stream / pre-normalisation tensors are fp32, tensor-core operands are bf16,
accumulation is always fp32.  A value's gradient uses Var.gdtype.
"""
from __future__ import annotations

import torch

from . import kernels as K
from .framework import Var, Variable, get_store, small_k

BF16 = torch.bfloat16
F32 = torch.float32


def _tape():
    return get_store().tape


def _rg(*xs) -> bool:
    tape = _tape()
    if tape is None:
        return False
    for x in xs:
        if x is None:
            continue
        if isinstance(x, Variable):
            if x.needs_grad:
                return True
        elif x.requires_grad:
            return True
    return False


def as_var(x) -> Var:
    return x if isinstance(x, Var) else Var(x)


# ------------------------------------------------------------------------------------------------ basics
def reshape(x: Var, shape) -> Var:
    out = Var(x.data.reshape(shape), grad_dtype=x.grad_dtype)
    if _rg(x):
        out.requires_grad = True

        def bwd():
            if out.grad is not None:
                x.accum(out.grad.reshape(x.data.shape))
        _tape().record(bwd)
    return out


def cast(x: Var, dtype) -> Var:
    if x.data.dtype == dtype:
        return x
    out = Var(K.cast(x.data, dtype))
    if _rg(x):
        out.requires_grad = True

        def bwd():
            if out.grad is not None:
                g = out.grad
                x.accum(g if g.dtype == x.gdtype else K.cast(g, x.gdtype))
        _tape().record(bwd)
    return out


def add(a: Var, b: Var) -> Var:
    """fp32 a + b (gradient accumulation / explicit residual sums)."""
    assert a.data.dtype == F32 and b.data.dtype == F32 and a.shape == b.shape
    data = K.cast(a.data, F32)  # copy
    K.axpby(b.data, data, 1.0, 1.0)
    out = Var(data)
    if _rg(a, b):
        out.requires_grad = True

        def bwd():
            g = out.grad
            if g is None:
                return
            if a.requires_grad:
                a.accum(g if g.dtype == a.gdtype else K.cast(g, a.gdtype))
            if b.requires_grad:
                gb = K.cast(g, b.gdtype) if (g.dtype != b.gdtype or a.requires_grad) else g
                b.accum(gb)
        _tape().record(bwd)
    return out


def split_rows(x: Var, sizes) -> list:
    """Row blocks of a 2-D (or n-D, first axis) value as separate Vars (views, no copy): x[0:s0], x[s0:s0+s1], ...
    Each block may then go through its own branch of the tape (framework.Tape.branch); the gradient of x is the
    concatenation of the blocks' gradients, formed when the backward pass reaches this node."""
    assert sum(sizes) == x.shape[0]
    outs, start = [], 0
    for n in sizes:
        outs.append(Var(x.data[start:start + n], grad_dtype=x.grad_dtype))
        start += n
    if _rg(x):
        for o in outs:
            o.requires_grad = True

        def bwd():
            gs = [o.grad for o in outs]
            if any(g is None for g in gs):
                if all(g is None for g in gs):
                    return
                raise RuntimeError("split_rows: some blocks received no gradient")
            x.accum(torch.cat([g.reshape((g.shape[0],) + tuple(x.data.shape[1:])) for g in gs], dim=0))   # tensor plumbing
        _tape().record(bwd)
    return outs


def concat_rows(a: Var, b: Var) -> Var:
    """tf.concat([a, b], axis=0) of two fp32 logit vectors / matrices (real | fake halves of a loss)."""
    assert a.data.dtype == F32 and b.data.dtype == F32 and a.shape[1:] == b.shape[1:]
    na = a.shape[0]
    data = torch.empty((na + b.shape[0],) + tuple(a.shape[1:]), dtype=F32, device=a.data.device)
    data[:na].copy_(a.data)      # device-to-device memcpy (allocator / tensor plumbing, no arithmetic)
    data[na:].copy_(b.data)
    out = Var(data)
    if _rg(a, b):
        out.requires_grad = True

        def bwd():
            g = out.grad
            if g is None:
                return
            if a.requires_grad:
                a.accum(g[:na].clone())
            if b.requires_grad:
                b.accum(g[na:].clone())
        _tape().record(bwd)
    return out


def add_scalars(a: Var, b: Var) -> Var:
    """Sum of two loss scalars (d_loss_gan + d_loss_acgan).  Loss Vars push their pre-computed dlogits when the tape
    unwinds (gan_loss / softmax_xent), so the sum itself has nothing to propagate."""
    data = K.cast(a.data, F32)
    K.axpby(b.data, data, 1.0, 1.0)
    out = Var(data)
    out.requires_grad = _rg(a, b)
    return out


def lerp(a: Var, b: Var, alpha) -> Var:
    """(1 - alpha) * a + alpha * b in fp32: the fade-in of PGGAN (PGGAN/model_nvidia.py:118, :200).  alpha: a Python
    float, or a 1-element fp32 device tensor (the placeholder of PGGAN/train.py:83 -- a captured CUDA graph then
    follows the value written into it)."""
    if not torch.is_tensor(alpha):
        alpha = torch.full((1,), float(alpha), dtype=F32, device=a.data.device)
    a32 = a if a.data.dtype == F32 else cast(a, F32)
    b32 = b if b.data.dtype == F32 else cast(b, F32)
    assert a32.shape == b32.shape
    out = Var(K.lerp_fwd(a32.data, b32.data, alpha))
    if _rg(a32, b32):
        out.requires_grad = True

        def bwd():
            g = out.grad
            if g is None:
                return
            g = g if g.dtype == F32 else K.cast(g, F32)
            da, db = K.lerp_bwd(g, alpha, a32.gdtype if a32.requires_grad else None,
                                b32.gdtype if b32.requires_grad else None)
            if da is not None:
                a32.accum(da)
            if db is not None:
                b32.accum(db)
        _tape().record(bwd)
    return out


# ------------------------------------------------------------------------------------------------ convolution
def _pads(padding, h, w, kh, kw, stride):
    if padding == "SAME":
        pt, _, ho = K.same_pads(h, kh, stride)
        pl, _, wo = K.same_pads(w, kw, stride)
    elif padding == "VALID":
        pt = pl = 0
        ho = (h - kh) // stride + 1
        wo = (w - kw) // stride + 1
    elif isinstance(padding, (tuple, list)) and len(padding) == 4:
        # explicit (top, bottom, left, right): tf.pad followed by a VALID convolution (Pix2Pix/networks.py:287-354)
        pt, pb, pl, pr = (int(v) for v in padding)
        ho = (h + pt + pb - kh) // stride + 1
        wo = (w + pl + pr - kw) // stride + 1
    else:
        raise ValueError(f"unknown padding {padding!r}")
    return pt, pl, ho, wo


# Batch statistics of a layer output accumulated in the convolution epilogue (ganb_conv2d_igemm_stats).  Built, tested
# (tests/test_gpu_fused.py) and measured: alone, conv + fused statistics beats conv + ganb_bn_stats (108.7 vs 97.2 + 26.9
# us on the dominant layer), but inside the D+G pair schedule the separate statistics kernel runs for free next to the
# other stream's tensor-core kernel while the longer epilogue holds the tensor pipe: 3.30 vs 3.20 ms per pair, same box
# (profiles/r02_fused_stats_ab.txt).  Default therefore OFF; GANB_FUSED_STATS=1 turns it on.
# relu / leaky relu between two convolutions without a normalisation in between (critic blocks) fused into the first
# one's epilogue, its derivative into the second one's data-gradient epilogue (GANB_FUSED_ACT=0: separate passes)
FUSED_CONV_ACT = __import__("os").environ.get("GANB_FUSED_ACT", "1") != "0"
FUSED_BN_STATS = __import__("os").environ.get("GANB_FUSED_STATS", "0") == "1"


def _stat_groups(n: int, groups: int | None = None) -> int:
    """Statistic towers of a batch of n (the rule of norm_act)."""
    g = groups if groups is not None else get_store().stat_groups
    return g if n % g == 0 else 1


def conv2d(x: Var, W: Variable, b: Variable | None, kh: int, kw: int, stride: int = 1, padding: str = "SAME",
           sn=None, residual: Var | None = None, out_grad_dtype=None, in_scale: float | None = None,
           residual_up2: bool = False, out_dtype=F32, bn_stats: bool = False, act: str | None = None) -> Var:
    """NHWC x HWIO cross-correlation (tf.nn.conv2d, common/ops/conv2d.py:181-187) + bias (+ residual), fp32 out.

    `sn` is a framework.SNEntry whose 1/sigma multiplies the accumulator (W/sigma is never materialised).
    `residual_up2`: the residual is given at half resolution and added through a nearest-2x upsample inside the
    GEMM epilogue (shortcut of an 'up' block); its gradient is the 2x2 block sum of the output gradient.
    Every shape runs on the tensor cores: layers with <= 8 channels on one side go through a bf16 im2col of the
    small tensor (kh*kw*c <= 32 columns) and become 1x1 GEMMs; larger small-channel filters fall back to the
    CUDA-core kernels of smallconv.cu.

    stride > 1 (Pix2Pix encoders / PatchGAN): the forward and filter-gradient kernels gather every stride-th pixel
    through TMA element strides; the data gradient is the stride-1 kernel applied to the zero-dilated output
    gradient (correct for any TF padding; the structural zeros cost stride^2 more MMA work than necessary).

    `act` ('relu' / 'lrelu'): the nonlinearity behind the layer is applied by the GEMM epilogue and only act(y) is stored
    (bf16).  The returned Var is marked (`fused_act`): its gradient is the gradient wrt the PRE-activation, delivered by
    the consuming conv2d, whose data-gradient epilogue multiplies by act' (read off the sign of its own input).  Used
    where no normalisation sits between two convolutions (the critic's residual blocks): saves the activation pass and
    its backward pass; bit-identical for relu (rounding to bf16 commutes with it).

    `bn_stats`: the output feeds a batch-statistics normalisation; where the kernel supports it the epilogue also leaves
    the per-tile column sums of y and y^2 (out.stats, consumed by norm_act instead of a separate pass over y)."""
    store = get_store()
    n, h, w, cin = x.shape
    cout = W.data.shape[-1]
    gate_act = x.fused_act
    if act is not None or gate_act is not None:
        # the fused pair is built for the bf16 tensor-core routes only (callers ask conv2d_act_fusable first)
        ok = act in (None, 'relu', 'lrelu') and FUSED_CONV_ACT and not store.is_tf32(W.root) and stride == 1
        if act is not None:
            ok = ok and out_dtype == BF16 and residual is None and not bn_stats and cout % 8 == 0 and \
                (cin % 8 == 0 or (cin < 8 and small_k(kh * kw, cin) is not None))
        if gate_act is not None:
            ok = ok and x.data.dtype == BF16 and cin % 8 == 0 and cout >= 8
        if not ok:
            raise NotImplementedError(f"conv2d: fused activation act={act} / gated input {gate_act} on this route "
                                      f"(cin={cin}, cout={cout}, stride={stride}, out {out_dtype})")
    if store.is_tf32(W.root):
        return _conv2d_tf32(x, W, b, kh, kw, stride, padding, sn, residual, out_grad_dtype, in_scale, residual_up2,
                            out_dtype)
    if cin > 8 and cin % 8:
        return _conv2d_ragged_cin(x, W, b, kh, kw, stride, padding, sn, residual, out_grad_dtype, in_scale,
                                  residual_up2, out_dtype)
    taps = kh * kw
    pt, pl, ho, wo = _pads(padding, h, w, kh, kw, stride)
    # inputs_norm (conv2d.py:93-95): conv(c * x, W) = c * conv(x, W), so the constant rides on the epilogue's alpha
    # (and on the filter gradient's scale) instead of costing a pass over x
    wscale = store.const(in_scale) if in_scale is not None else None
    if in_scale is not None and sn is not None:
        # the ResNet PGGAN critic (common/resnet_block.py:276-337): y = (c / sigma) * conv(x, W); the filter gradient
        # wrt W / sigma carries c (wscale below), the data gradient c / sigma of THIS evaluation
        alpha = K.cast(sn.inv_sigma, F32, scale=in_scale)
    else:
        alpha = sn.inv_sigma if sn is not None else wscale
    bias = b.data if b is not None else None
    res = residual.data if residual is not None else None
    small_in, small_out = cin < 8, cout < 8   # 8 channels already satisfy the 16-byte rows of the TMA path
    kp_in, kp_out = small_k(taps, cin), small_k(taps, cout)
    route_in = small_in and kp_in is not None and cout % 8 == 0
    route_out = (not small_in) and small_out and kp_out is not None and cin % 8 == 0 and stride == 1
    if stride != 1 and not (route_in or (cin % 8 == 0 and not small_out)):
        raise NotImplementedError(f"stride {stride} with cin={cin}, cout={cout}: only the tensor-core routes are strided")
    group = store.pack_group(W.root)
    pack = group.entry(W)
    group.refresh()
    xcol = None
    fused = None
    if small_in:
        xin = x if x.data.dtype == F32 else cast(x, F32)
        if route_in:
            xcol = K.im2col_small(xin.data, n, h, w, cin, ho, wo, kh, kw, pt, pl, +1, kp_in, stride=stride)
            y = K.conv_igemm(xcol, pack.ws, n, ho, wo, kp_in, ho, wo, cout, 1, 1, 0, 0, False, alpha, bias, res,
                             act, out_dtype, residual_up2=residual_up2)
        else:
            if res is not None or act is not None:
                raise NotImplementedError("residual / fused activation on the CUDA-core small-channel path")
            y = K.conv_smallcin(xin.data, W.data, n, h, w, cin, ho, wo, cout, kh, kw, pt, pl, False, False, alpha,
                                bias, None, F32)
    else:
        if cin % 8:
            raise NotImplementedError(f"cin={cin}: tensor-core path needs cin % 8 == 0")
        xin = x if x.data.dtype == BF16 else cast(x, BF16)
        g_stats = _stat_groups(n) if (bn_stats and FUSED_BN_STATS) else 0
        if g_stats and K.conv_stats_rows(n, ho, wo, cout, kh, kw, stride, g_stats):
            y, fused = K.conv_igemm_stats(xin.data, pack.wt, n, h, w, cin, ho, wo, cout, kh, kw, pt, pl, False, alpha,
                                          bias, res, None, out_dtype, g_stats, residual_up2=residual_up2, stride=stride)
        else:
            y = K.conv_igemm(xin.data, pack.wt, n, h, w, cin, ho, wo, cout, kh, kw, pt, pl, False, alpha, bias, res,
                             act, out_dtype, residual_up2=residual_up2, stride=stride)
    out = Var(y, grad_dtype=out_grad_dtype)
    out.stats = fused
    out.fused_act = act
    need_w = W.needs_grad and _tape() is not None
    need_b = b is not None and b.needs_grad and _tape() is not None
    if _rg(xin, residual) or need_w or need_b:
        out.requires_grad = True
        tape = _tape()

        def bwd():
            gy = out.grad
            if gy is None:
                return
            if residual is not None and residual.requires_grad:
                if residual_up2:
                    residual.accum(K.sum2x2(gy, 1.0, residual.gdtype))
                else:
                    residual.accum(gy if gy.dtype == residual.gdtype else K.cast(gy, residual.gdtype))
            need_x = xin.requires_grad
            if not (need_w or need_x or need_b):
                return
            gy16 = None
            if not small_out and (need_w or need_x):
                gy16 = gy if gy.dtype == BF16 else K.cast(gy, BF16)
            dycol = None
            if route_out and (need_w or need_x):
                gy32 = gy if gy.dtype == F32 else K.cast(gy, F32)
                dycol = K.im2col_small(gy32, n, ho, wo, cout, h, w, kh, kw, pt, pl, -1, kp_out)

            def _conv_wgrad():
                if sn is not None:
                    dst, beta = sn.g, (1.0 if sn.g_written else 0.0)
                else:
                    dst, beta = W.grad, 1.0
                if route_in:
                    r = torch.empty((kp_in, cout), dtype=F32, device=gy.device)
                    K.conv_wgrad(xcol, gy16, r, n, ho, wo, kp_in, ho, wo, cout, 1, 1, 0, 0, None, 0.0)
                    K.small_wgrad_scatter(r, dst, taps, cin, cout, False, wscale, beta)
                elif route_out:
                    r = torch.empty((kp_out, cin), dtype=F32, device=gy.device)
                    K.conv_wgrad(dycol, xin.data, r, n, h, w, kp_out, h, w, cin, 1, 1, 0, 0, None, 0.0)
                    K.small_wgrad_scatter(r, dst, taps, cout, cin, True, wscale, beta)
                elif small_in:
                    K.conv_small_wgrad(xin.data, gy16, dst, n, h, w, cin, ho, wo, cout, kh, kw, pt, pl, +1, False,
                                       wscale, beta)
                elif small_out:
                    gy32 = gy if gy.dtype == F32 else K.cast(gy, F32)
                    K.conv_small_wgrad(gy32, xin.data, dst, n, ho, wo, cout, h, w, cin, kh, kw, pt, pl, -1, True,
                                       wscale, beta)
                else:
                    K.conv_wgrad(xin.data, gy16, dst, n, h, w, cin, ho, wo, cout, kh, kw, pt, pl, wscale, beta,
                                 stride=stride)
                if sn is not None:
                    sn.g_written = True
                    lst = tape.pending_sn.setdefault(W.root, [])
                    if sn not in lst:
                        lst.append(sn)

            # ---- off the critical chain: bias and filter gradients (side stream; joined at the end of backward)
            if need_b or need_w:
                tape.keep.extend(t for t in (gy, gy16, dycol) if t is not None)
                if need_w:
                    with tape.offchain():
                        _conv_wgrad()
                if need_b:
                    with tape.offchain(1):
                        K.colsum(gy, n * ho * wo, cout, b.grad, 1.0)
            if need_x:
                gdt = xin.gdtype
                if route_out:
                    dx = K.conv_igemm(dycol, pack.ws, n, h, w, kp_out, h, w, cin, 1, 1, 0, 0, False, alpha, None,
                                      None, None, gdt)
                elif small_out:
                    gy32 = gy if gy.dtype == F32 else K.cast(gy, F32)
                    dx = K.conv_smallcin(gy32, W.data, n, ho, wo, cout, h, w, cin, kh, kw, kh - 1 - pt, kw - 1 - pl,
                                         True, True, alpha, None, None, gdt)
                elif gate_act is not None:
                    dx = K.conv_igemm_gated(gy16, pack.wn, n, ho, wo, cout, h, w, cin, kh, kw, kh - 1 - pt, kw - 1 - pl,
                                            True, alpha, xin.data, gate_act, gdt)
                elif stride == 1:
                    dx = K.conv_igemm(gy16, pack.wn, n, ho, wo, cout, h, w, cin, kh, kw, kh - 1 - pt, kw - 1 - pl, True,
                                      alpha, None, None, None, gdt)
                else:
                    hd, wd = (ho - 1) * stride + 1, (wo - 1) * stride + 1
                    gyd = K.dilate2d(gy16, stride, hd, wd)
                    dx = K.conv_igemm(gyd, pack.wn, n, hd, wd, cout, h, w, cin, kh, kw, kh - 1 - pt, kw - 1 - pl, True,
                                      alpha, None, None, None, gdt)
                xin.accum(dx, gated=gate_act is not None)
        tape.record(bwd)
    return out


def conv2d_act_fusable(cin1: int, mid: int, cout2: int, k: int, stride: int = 1, root: str | None = None) -> bool:
    """Can the activation between Conv1 (cin1 -> mid, k x k) and Conv2 (mid -> cout2), both stride-1 plain convolutions of
    the network being built, be fused into Conv1's epilogue and Conv2's data-gradient epilogue (conv2d(act=...))?  Both
    must take the bf16 tensor-core routes: Conv1 the TMA route or the im2col route of an RGB-sided layer, Conv2 the
    TMA route."""
    if not FUSED_CONV_ACT or get_store().is_tf32(root) or stride != 1:
        return False
    producer = cin1 % 8 == 0 or (cin1 < 8 and small_k(k * k, cin1) is not None)
    return producer and mid % 8 == 0 and mid >= 8 and cout2 >= 8


def _conv2d_tf32(x, W, b, kh, kw, stride, padding, sn, residual, out_grad_dtype, in_scale, residual_up2, out_dtype):
    """conv2d of a network in TF32 operand mode (VariableStore.set_precision(root, 'tf32')): fp32 activations and filters,
    rounded to TF32 (round-to-nearest) and contracted by the kind::tf32 tensor-core kernels with fp32 accumulation;
    gradients the same way (fp32 storage).  Sides with fewer than 8 channels (RGB) take the fp32 CUDA-core kernels of
    smallconv.cu, which are exact.  Per-layer error vs fp32: ~3e-4 (tests/test_gpu_tf32.py; tolerance 1e-3)."""
    store = get_store()
    n, h, w, cin = x.shape
    cout = W.data.shape[-1]
    taps = kh * kw
    pt, pl, ho, wo = _pads(padding, h, w, kh, kw, stride)
    wscale = store.const(in_scale) if in_scale is not None else None
    if in_scale is not None and sn is not None:
        alpha = K.cast(sn.inv_sigma, F32, scale=in_scale)
    else:
        alpha = sn.inv_sigma if sn is not None else wscale
    bias = b.data if b is not None else None
    xin = x if x.data.dtype == F32 else cast(x, F32)
    res = residual.data if residual is not None else None
    if res is not None and res.dtype != F32:
        res = K.cast(res, F32)
    small_in = cin < 8
    small_out = (not small_in) and cout < 8
    small = small_in or small_out
    if small and (stride != 1 or res is not None):
        raise NotImplementedError("TF32 mode: small-channel layers are plain stride-1 convolutions (RGB sides)")
    if not small_in and (cin % 4 or (cout % 4 and not small_out)):
        raise NotImplementedError(f"TF32 mode: cin={cin}, cout={cout} must be multiples of 4")
    wr = None
    if small_in:
        y = K.conv_smallcin(xin.data, W.data, n, h, w, cin, ho, wo, cout, kh, kw, pt, pl, False, False, alpha, bias,
                            None, F32)
        xr = xin.data
    else:
        wt, wr = store.tf32_operands(W)
        xr = K.round_tf32(xin.data)
        y = K.conv_igemm_tf32(xr, wt, n, h, w, cin, ho, wo, cout, kh, kw, pt, pl, False, alpha, bias, res, None,
                              out_dtype, residual_up2=residual_up2, stride=stride)
    out = Var(y, grad_dtype=out_grad_dtype if out_grad_dtype in (None, F32) else F32)
    need_w = W.needs_grad and _tape() is not None
    need_b = b is not None and b.needs_grad and _tape() is not None
    if _rg(xin, residual) or need_w or need_b:
        out.requires_grad = True
        tape = _tape()

        def bwd():
            gy = out.grad
            if gy is None:
                return
            gy32 = gy if gy.dtype == F32 else K.cast(gy, F32)
            if residual is not None and residual.requires_grad:
                if residual_up2:
                    residual.accum(K.sum2x2(gy32, 1.0, residual.gdtype))
                else:
                    residual.accum(gy32 if residual.gdtype == F32 else K.cast(gy32, residual.gdtype))
            need_x = xin.requires_grad
            if not (need_w or need_x or need_b):
                return
            if sn is not None:
                dst, beta = sn.g, (1.0 if sn.g_written else 0.0)
            else:
                dst, beta = W.grad, 1.0
            if need_b:
                K.colsum(gy32, n * ho * wo, cout, b.grad, 1.0)
            if small:
                if small_in:
                    if need_w:
                        K.conv_small_wgrad(xin.data, gy32, dst, n, h, w, cin, ho, wo, cout, kh, kw, pt, pl, +1, False,
                                           wscale, beta)
                    if need_x:   # narrow output (3 channels): the tensor-core data-gradient kernel with a 16-wide tile
                        _, wr_in = store.tf32_operands(W)
                        xin.accum(K.conv_igemm_tf32(K.round_tf32(gy32), wr_in, n, ho, wo, cout, h, w, cin, kh, kw,
                                                    kh - 1 - pt, kw - 1 - pl, True, alpha, None, None, None, F32))
                else:
                    if need_w:
                        K.conv_small_wgrad(gy32, xin.data, dst, n, ho, wo, cout, h, w, cin, kh, kw, pt, pl, -1, True,
                                           wscale, beta)
                    if need_x:
                        xin.accum(K.conv_smallcin(gy32, W.data, n, ho, wo, cout, h, w, cin, kh, kw, kh - 1 - pt,
                                                  kw - 1 - pl, True, True, alpha, None, None, F32))
            else:
                gyr = K.round_tf32(gy32)
                if need_w:
                    K.conv_wgrad_tf32(xr, gyr, dst, n, h, w, cin, ho, wo, cout, kh, kw, pt, pl, wscale, beta, stride=stride)
                if need_x:
                    if stride == 1:
                        dx = K.conv_igemm_tf32(gyr, wr, n, ho, wo, cout, h, w, cin, kh, kw, kh - 1 - pt, kw - 1 - pl, True,
                                               alpha, None, None, None, F32)
                    else:
                        hd, wd = (ho - 1) * stride + 1, (wo - 1) * stride + 1
                        gyd = K.dilate2d(gyr, stride, hd, wd)
                        dx = K.conv_igemm_tf32(gyd, wr, n, hd, wd, cout, h, w, cin, kh, kw, kh - 1 - pt, kw - 1 - pl, True,
                                               alpha, None, None, None, F32)
                    xin.accum(dx)
            if need_w and sn is not None:
                sn.g_written = True
                lst = tape.pending_sn.setdefault(W.root, [])
                if sn not in lst:
                    lst.append(sn)
        tape.record(bwd)
    return out


def depthwise_conv2d(x: Var, Wd: Variable, b: Variable | None, stride: int = 1, padding="SAME", out_dtype=BF16,
                     out_grad_dtype=F32, sn=None) -> Var:
    """tf.nn.depthwise_conv2d(x, Wd [kh, kw, c, cm]) (+ bias): the `depthwise_conv2d` conv_type and the first half of
    `separable_conv2d` (common/ops/conv2d.py:188-208).  Bandwidth-bound CUDA-core kernels in fp32 arithmetic; the input
    is read in bf16 (like every convolution operand of this library) and the output is the bf16 operand of the pointwise
    1x1 convolution that follows, with an fp32 gradient.

    `sn`: the framework.SNEntry of `depthwise_filters` (conv2d.py:173-175).  These kernels have no GEMM epilogue to carry
    1/sigma, and the filter is tiny, so W_d / sigma is formed explicitly; the filter gradient is dL/d(W_d / sigma),
    handed to the spectral-norm backward of the network through sn.g like every other normalised weight."""
    n, h, w, c = x.shape
    kh, kw, c2, cm = Wd.data.shape
    assert c2 == c, (c2, c)
    pt, pl, ho, wo = _pads(padding, h, w, kh, kw, stride)
    xin = x if x.data.dtype == BF16 else cast(x, BF16)
    w_eff = Wd.data if sn is None else K.scale_dev(Wd.data, sn.inv_sigma)
    y = K.depthwise_fwd(xin.data, w_eff, b.data if b is not None else None, ho, wo, stride, pt, pl, out_dtype)
    out = Var(y, grad_dtype=out_grad_dtype)
    need_w = Wd.needs_grad and _tape() is not None
    need_b = b is not None and b.needs_grad and _tape() is not None
    if _rg(xin) or need_w or need_b:
        out.requires_grad = True
        tape = _tape()

        def bwd():
            gy = out.grad
            if gy is None:
                return
            if need_b:
                K.colsum(gy, n * ho * wo, c * cm, b.grad, 1.0)
            if need_w and sn is None:
                K.depthwise_bwd_filter(xin.data, gy, Wd.grad, kh, kw, cm, stride, pt, pl)
            elif need_w:
                if not sn.g_written:
                    sn.g.zero_()                    # depthwise_bwd_filter adds into its destination
                K.depthwise_bwd_filter(xin.data, gy, sn.g.view(Wd.data.shape), kh, kw, cm, stride, pt, pl)
                sn.g_written = True
                lst = tape.pending_sn.setdefault(Wd.root, [])
                if sn not in lst:
                    lst.append(sn)
            if xin.requires_grad:
                xin.accum(K.depthwise_bwd_input(gy, w_eff, h, w, stride, pt, pl, xin.gdtype))
        tape.record(bwd)
    return out


SUBPIXEL_UPCONV = True   # sub-pixel form of UpsampleConv where ganb_upconv_supported() and the layer is large enough


def upconv_eligible(n: int, h: int, w: int, cin: int, cout: int, k: int = 3) -> bool:
    """True when UpsampleConv(k x k) over a [n, h, w, cin] input runs in its sub-pixel form: the shapes the CTA-pair
    kernel tiles (ganb_upconv_supported) and enough pixel tiles to fill the machine in the data-gradient pass."""
    return (SUBPIXEL_UPCONV and k == 3 and n * h * w >= 16384 and K.upconv_supported(n, h, w, cin, cout))


def upconv2d(x: Var, W: Variable, b: Variable | None, out_grad_dtype=None, out_dtype=BF16,
             bn_stats: bool = False) -> Var:
    """UpsampleConv (common/resnet_block.py:83-97) = nearest 2x + 3x3 SAME Conv2D + bias, computed as four 2x2
    convolutions over the LOW-resolution x (4/9 of the MMA work, the upsampled tensor is never written).

    Returns a Var of logical shape [n, 2h, 2w, cout] whose storage is in QUAD LAYOUT (`quad = True`): only batch
    statistics / norm_act (which translate the pixel order) and column sums may consume it."""
    store = get_store()
    n, h, w, cin = x.shape
    cout = W.data.shape[-1]
    group = store.pack_group(W.root)
    pack = group.entry(W)
    if pack.enable_upconv():
        group.valid_for = None
    group.refresh()
    xin = x if x.data.dtype == BF16 else cast(x, BF16)
    g_stats = _stat_groups(n) if (bn_stats and FUSED_BN_STATS) else 0
    fused = None
    if g_stats and K.upconv_stats_rows(n, h, w, cin, cout, g_stats):
        y, fused = K.upconv_fprop_stats(xin.data, pack.we_t, n, h, w, cin, cout, None,
                                        b.data if b is not None else None, None, out_dtype, g_stats)
    else:
        y = K.upconv_fprop(xin.data, pack.we_t, n, h, w, cin, cout, None, b.data if b is not None else None, None,
                           out_dtype)
    out = Var(y, grad_dtype=out_grad_dtype)
    out.stats = fused
    out.quad = True
    need_w = W.needs_grad and _tape() is not None
    need_b = b is not None and b.needs_grad and _tape() is not None
    if _rg(xin) or need_w or need_b:
        out.requires_grad = True
        tape = _tape()

        def bwd():
            gy = out.grad          # quad layout, like the forward value
            if gy is None:
                return
            gy16 = gy if gy.dtype == BF16 else K.cast(gy, BF16)
            if need_b or need_w:
                tape.keep.extend((gy, gy16))
                if need_w:
                    with tape.offchain():
                        K.upconv_wgrad(xin.data, gy16, W.grad, n, h, w, cin, cout, None, 1.0)
                if need_b:
                    with tape.offchain(1):
                        K.colsum(gy, n * 4 * h * w, cout, b.grad, 1.0)
            if xin.requires_grad:
                xin.accum(K.upconv_dgrad(gy16, pack.we_n, n, h, w, cin, cout, None, xin.gdtype))
        tape.record(bwd)
    return out


def _conv2d_ragged_cin(x, W, b, kh, kw, stride, padding, sn, residual, out_grad_dtype, in_scale, residual_up2,
                       out_dtype):
    """Convolution whose input-channel count is above 8 and not a multiple of 8 (the 513 channels behind
    minibatch_std, PGGAN/model_nvidia.py:223-229): x and the bf16 operand copies of the filter are zero-padded to the
    next multiple of 8 channels; the padded rows of the filter gradient / channels of the data gradient are dropped.
    Such layers sit at 4x4 resolution, so the two extra copies are noise."""
    if stride != 1 or residual is not None or in_scale is not None:
        raise NotImplementedError("ragged input-channel counts: only plain stride-1 convolutions are built")
    store = get_store()
    n, h, w, cin = x.shape
    cout = W.data.shape[-1]
    taps = kh * kw
    pt, pl, ho, wo = _pads(padding, h, w, kh, kw, 1)
    group = store.pack_group(W.root)
    pack = group.entry(W)
    group.refresh()
    cp = pack.ci_pad
    xp = torch.zeros((n, h, w, cp), dtype=BF16, device=x.data.device)
    K.copy_channels(x.data, 0, xp, 0, cin)
    alpha = sn.inv_sigma if sn is not None else None
    y = K.conv_igemm(xp, pack.wt, n, h, w, cp, ho, wo, cout, kh, kw, pt, pl, False, alpha,
                     b.data if b is not None else None, None, None, out_dtype)
    out = Var(y, grad_dtype=out_grad_dtype)
    need_w = W.needs_grad and _tape() is not None
    need_b = b is not None and b.needs_grad and _tape() is not None
    if _rg(x) or need_w or need_b:
        out.requires_grad = True
        tape = _tape()

        def bwd():
            gy = out.grad
            if gy is None:
                return
            gy16 = gy if gy.dtype == BF16 else K.cast(gy, BF16)
            if need_b:
                K.colsum(gy, n * ho * wo, cout, b.grad, 1.0)
            if need_w:
                dwp = torch.empty((taps, cp, cout), dtype=F32, device=gy.device)
                K.conv_wgrad(xp, gy16, dwp, n, h, w, cp, ho, wo, cout, kh, kw, pt, pl, None, 0.0)
                dw = torch.empty((taps, cin * cout), dtype=F32, device=gy.device)
                K.copy_channels(dwp.reshape(taps, cp * cout), 0, dw, 0, cin * cout)
                if sn is not None:
                    dst, beta = sn.g, (1.0 if sn.g_written else 0.0)
                else:
                    dst, beta = W.grad, 1.0
                K.axpby(dw, dst, 1.0, beta)
                if sn is not None:
                    sn.g_written = True
                    lst = tape.pending_sn.setdefault(W.root, [])
                    if sn not in lst:
                        lst.append(sn)
            if x.requires_grad:
                dxp = K.conv_igemm(gy16, pack.wn, n, ho, wo, cout, h, w, cp, kh, kw, kh - 1 - pt, kw - 1 - pl, True,
                                   alpha, None, None, None, F32)
                dx = torch.empty((n, h, w, cin), dtype=x.gdtype, device=gy.device)
                K.copy_channels(dxp, 0, dx, 0, cin)
                x.accum(dx)
        tape.record(bwd)
    return out


def conv2d_transpose(x: Var, W: Variable, b: Variable | None, kh: int, kw: int, stride: int = 2,
                     padding: str = "SAME", out_grad_dtype=None) -> Var:
    """tf.nn.conv2d_transpose to [n, stride*h, stride*w, cout] (common/ops/deconv2d.py:99-109) + bias, fp32 out.

    W is [kh, kw, cout, cin] (the HWIO filter of the forward convolution [n, s*h, s*w, cout] -> [n, h, w, cin] whose
    input gradient this op is).  Forward = stride-1 tensor-core kernel over the zero-dilated input with the filter
    taps flipped; backward: dx = the strided forward convolution of the output gradient, dW = its filter gradient
    with the roles of activation and gradient exchanged."""
    n, h, w, cin = x.shape
    cout = W.data.shape[-2]
    if W.data.shape[-1] != cin:
        raise ValueError(f"conv2d_transpose: filter expects {W.data.shape[-1]} input channels, got {cin}")
    if cin % 8 or cout % 8:
        raise NotImplementedError("conv2d_transpose: channel counts must be multiples of 8")
    oh, ow = stride * h, stride * w
    pt, pl, ho_chk, wo_chk = _pads(padding, oh, ow, kh, kw, stride)   # pads of the forward conv on the big side
    if (ho_chk, wo_chk) != (h, w):
        raise ValueError("conv2d_transpose: padding does not map the 2x output back onto the input size")
    store = get_store()
    group = store.pack_group(W.root)
    pack = group.entry(W)
    group.refresh()
    xin = x if x.data.dtype == BF16 else cast(x, BF16)
    hd, wd = (h - 1) * stride + 1, (w - 1) * stride + 1
    xd = K.dilate2d(xin.data, stride, hd, wd)
    y = K.conv_igemm(xd, pack.wn, n, hd, wd, cin, oh, ow, cout, kh, kw, kh - 1 - pt, kw - 1 - pl, True, None,
                     b.data if b is not None else None, None, None, F32)
    out = Var(y, grad_dtype=out_grad_dtype)
    need_w = W.needs_grad and _tape() is not None
    need_b = b is not None and b.needs_grad and _tape() is not None
    if _rg(xin) or need_w or need_b:
        out.requires_grad = True
        tape = _tape()

        def bwd():
            gy = out.grad
            if gy is None:
                return
            gy16 = gy if gy.dtype == BF16 else K.cast(gy, BF16)
            if need_b or need_w:
                tape.keep.extend((gy, gy16))
                if need_w:   # filter gradient of the forward conv: "input" = gy (cout channels), "dy" = x
                    with tape.offchain():
                        K.conv_wgrad(gy16, xin.data, W.grad, n, oh, ow, cout, h, w, cin, kh, kw, pt, pl, None, 1.0,
                                     stride=stride)
                if need_b:
                    with tape.offchain(1):
                        K.colsum(gy, n * oh * ow, cout, b.grad, 1.0)
            if xin.requires_grad:
                dx = K.conv_igemm(gy16, pack.wt, n, oh, ow, cout, h, w, cin, kh, kw, pt, pl, False, None, None, None,
                                  None, xin.gdtype, stride=stride)
                xin.accum(dx)
        tape.record(bwd)
    return out


def linear(x: Var, W: Variable, b: Variable | None, sn=None, out_dtype=F32, in_scale: float | None = None) -> Var:
    """tf.matmul(x, W) + b (common/ops/linear.py:161-180) for 2-D x; large layers run as 1x1 convolutions on the
    tensor cores, tiny ones (in % 8 != 0 or out < 8) on CUDA cores.  in_scale = the inputs_norm constant
    sqrt(2 / in) (linear.py:47-49), applied as the GEMM's alpha."""
    m, kin = x.shape
    kout = W.data.shape[1]
    if kin % 8 == 0 and kout % 8 == 0 and kin * kout >= 65536:
        x4 = reshape(x, (m, 1, 1, kin))
        y4 = conv2d(x4, W, b, 1, 1, 1, "VALID", sn=sn, out_dtype=out_dtype, in_scale=in_scale)
        return reshape(y4, (m, kout))
    if out_dtype != F32:
        raise NotImplementedError("bf16 output is only built for the tensor-core linear path")
    xin = x if x.data.dtype == F32 else cast(x, F32)
    wscale = get_store().const(in_scale) if in_scale is not None else None     # D.Output of the ResNet PGGAN critic
    if in_scale is not None and sn is not None:
        alpha = K.cast(sn.inv_sigma, F32, scale=in_scale)
    else:
        alpha = sn.inv_sigma if sn is not None else wscale
    y = torch.empty((m, kout), dtype=F32, device=x.data.device)
    K.sgemm_small(xin.data, W.data, y, m, kout, kin, False, False, alpha, b.data if b is not None else None, 0.0)
    out = Var(y)
    need_w = W.needs_grad and _tape() is not None
    need_b = b is not None and b.needs_grad and _tape() is not None
    if _rg(xin) or need_w or need_b:
        out.requires_grad = True
        tape = _tape()

        def bwd():
            gy = out.grad
            if gy is None:
                return
            gy = gy if gy.dtype == F32 else K.cast(gy, F32)
            if need_b:
                K.colsum(gy, m, kout, b.grad, 1.0)
            if need_w:
                if sn is not None:
                    K.sgemm_small(xin.data, gy, sn.g, kin, kout, m, True, False, wscale, None,
                                  1.0 if sn.g_written else 0.0)
                    sn.g_written = True
                    lst = tape.pending_sn.setdefault(W.root, [])
                    if sn not in lst:
                        lst.append(sn)
                else:
                    K.sgemm_small(xin.data, gy, W.grad, kin, kout, m, True, False, wscale, None, 1.0)
            if xin.requires_grad:
                dx = torch.empty((m, kin), dtype=F32, device=gy.device)
                K.sgemm_small(gy, W.data, dx, m, kin, kout, False, True, alpha, None, 0.0)
                xin.accum(dx if xin.gdtype == F32 else K.cast(dx, xin.gdtype))
        tape.record(bwd)
    return out


# ------------------------------------------------------------------------------------------------ norm + act
def layer_norm(x: Var, gamma: Variable, beta: Variable, eps: float = 1e-12, act=None, out_dtype=BF16,
               out_grad_dtype=None) -> Var:
    """act(tf.contrib.layers.layer_norm(x, begin_norm_axis=1, begin_params_axis=-1)): per-sample moments over (h, w, c),
    per-channel gamma / beta (common/ops/normalization.py:62-82; variance_epsilon 1e-12 is contrib's constant)."""
    y, mr = K.layer_norm_fwd(x.data, gamma.data, beta.data, eps, act, out_dtype)
    out = Var(y, grad_dtype=out_grad_dtype)
    need_p = gamma.needs_grad and _tape() is not None
    if _rg(x) or need_p:
        out.requires_grad = True

        def bwd():
            gz = out.grad
            if gz is None or not (x.requires_grad or need_p):
                return
            dx = K.layer_norm_bwd(x.data, gz, mr, gamma.data, beta.data, act, x.gdtype if x.requires_grad else None,
                                  gamma.grad if need_p else None, beta.grad if need_p else None)
            if dx is not None:
                x.accum(dx)
        _tape().record(bwd)
    return out


def norm_act(x: Var, *, stats: str | None, eps: float = 1e-5, gamma: Variable | None = None,
             beta: Variable | None = None, labels: torch.Tensor | None = None, act=None, upsample: bool = False,
             out_dtype=BF16, want_raw: bool = False, groups: int | None = None, out_grad_dtype=None):
    """act(normalise(x)) [-> nearest 2x upsample], optionally also the raw bf16 copy of x (shortcut operand).

    stats: None (identity), 'batch' (moments over n,h,w per statistic group; cond. BN when labels given) or
    'instance' (per-sample moments).  Returns (out, raw_or_None)."""
    store = get_store()
    n, h, w, c = x.shape
    mean = rstd = None
    sync_key = None
    g = 1
    if stats == "batch":
        g = groups if groups is not None else store.stat_groups
        if n % g:
            g = 1
    elif stats == "instance":
        g = n
    elif stats is not None:
        raise ValueError(stats)
    # cross-GPU batch statistics (store.bn_sync = (allreduce_sum, world)): only for statistics over the BATCH
    sync = store.bn_sync if stats == "batch" else None
    if stats is not None:
        fused = x.stats if stats == "batch" else None
        if fused is not None and fused.groups == g and fused.c == c:
            mean, rstd = fused.finalize((n // g) * h * w, eps)     # sums left by the producing conv's epilogue
        else:
            mean, rstd = K.bn_stats(x.data, n, h * w, c, g, eps)
        if sync is not None:
            # one exchange site per (layer, batch size): the generator forward of the critic step and of the generator
            # step may be in flight at the same time on two streams (Trainer.pair_step)
            sync_key = f"{store.scope_name()}/n{n}x{h}x{w}x{c}"
            K.bn_stats_sync(mean, rstd, eps, sync, sync_key)
    gam = gamma.data if gamma is not None else None
    bet = beta.data if beta is not None else None
    # the raw bf16 copy (1x1-shortcut operand) is x itself when x is already stored in bf16
    # (in TF32 operand mode the activations are fp32 and the shortcut convolution reads x itself)
    share_raw = want_raw and (x.data.dtype == BF16 or (out_dtype == F32 and x.data.dtype == F32))
    raw = torch.empty((n, h, w, c), dtype=BF16, device=x.data.device) if (want_raw and not share_raw) else None
    quad = bool(getattr(x, "quad", False))   # x is the quad-layout output of upconv2d; the result is plain NHWC
    if quad and (upsample or want_raw):
        raise NotImplementedError("a quad-layout tensor can only be normalised in place (no upsample / raw copy)")
    ups = 2 if quad else upsample
    y = K.norm_act_fwd(x.data, n, h, w, c, mean, rstd, g, gam, bet, labels, act, ups, out_dtype, out_raw=raw)
    out = Var(y, grad_dtype=out_grad_dtype)
    raw_var = x if share_raw else (Var(raw) if want_raw else None)  # bf16 value, bf16 gradient
    need_p = gamma is not None and gamma.needs_grad and _tape() is not None
    if _rg(x) or need_p:
        out.requires_grad = True
        if raw_var is not None and not share_raw:
            raw_var.requires_grad = x.requires_grad

        def bwd():
            gz = out.grad
            extra = raw_var.grad if (raw_var is not None and not share_raw) else None
            if gz is None:
                if extra is not None and x.requires_grad:
                    x.accum(extra)
                return
            dgam = gamma.grad if need_p else None
            dbet = beta.grad if need_p else None
            if not x.requires_grad and not need_p:
                return
            # a second gradient path into x (shortcut operand, or an identity-shortcut residual already accumulated
            # in x.grad) is added inside the same kernel instead of a separate read-modify-write pass
            if (extra is None and x.grad is not None and x.requires_grad and x.grad.shape == x.data.shape
                    and not quad):
                extra, x.grad = x.grad, None
            dx = K.norm_act_bwd(x.data, gz, 0, n, h, w, c, mean, rstd, g, gam, bet, labels, act, ups, dgam, dbet,
                                extra, x.gdtype, sync=sync, sync_key=sync_key)
            if x.requires_grad:
                x.accum(dx)
        _tape().record(bwd)
    return out, raw_var


def activation(x: Var, act, out_dtype=None) -> Var:
    """Stand-alone nonlinearity on a tensor of any shape (viewed as rows of 4)."""
    shape = x.shape
    total = x.data.numel()
    if total % 4:
        raise NotImplementedError("activation: element count must be a multiple of 4")
    xin = x if x.data.dtype == F32 else cast(x, F32)
    flat = reshape(xin, (1, 1, total // 4, 4))
    y, _ = norm_act(flat, stats=None, act=act, out_dtype=out_dtype or F32)
    return reshape(y, shape)


def meanpool2(x: Var, addend: Var | None = None, in_grad_dtype=None, out_grad_dtype=None) -> Var:
    """2x2 mean pool (+ addend), fp32 out (common/resnet_block.py:62-63)."""
    n, h, w, c = x.shape
    y = K.meanpool2(x.data, addend.data if addend is not None else None, F32)
    out = Var(y, grad_dtype=out_grad_dtype)
    if _rg(x, addend):
        out.requires_grad = True

        def bwd():
            g = out.grad
            if g is None:
                return
            if addend is not None and addend.requires_grad:
                addend.accum(g if g.dtype == addend.gdtype else K.cast(g, addend.gdtype))
            if x.requires_grad:
                x.accum(K.expand2(g, 0.25, in_grad_dtype or x.gdtype))
        _tape().record(bwd)
    return out


def upsample2(x: Var, out_dtype=None) -> Var:
    """Nearest-neighbour 2x (tf.depth_to_space of four concatenated copies, common/resnet_block.py:87-88)."""
    y = K.expand2(x.data, 1.0, out_dtype or x.data.dtype)
    out = Var(y)
    if _rg(x):
        out.requires_grad = True

        def bwd():
            if out.grad is not None:
                x.accum(K.sum2x2(out.grad, 1.0, x.gdtype))
        _tape().record(bwd)
    return out


def subsample2(x: Var, out_dtype=None) -> Var:
    """tf.image.resize_nearest_neighbor to half the size: x[:, ::2, ::2, :] (common/resnet_block.py:286-287)."""
    n, h, w, c = x.shape
    if h % 2 or w % 2:   # TF picks floor(i * h / out_h): only for even sizes is that every second pixel
        raise NotImplementedError("subsample2: odd sizes (no reference call-site; PGGAN stages are powers of two)")
    out = Var(K.subsample2d(x.data, 2, out_dtype or x.data.dtype))
    if _rg(x):
        out.requires_grad = True

        def bwd():
            if out.grad is not None:
                x.accum(K.subsample2d_bwd(out.grad, 2, h, w, x.gdtype))
        _tape().record(bwd)
    return out


def act_mean_hw(x: Var, act) -> Var:
    """mean over (h, w) of act(x): nonlinearity + tf.reduce_mean(axis=[1,2])."""
    assert x.data.dtype == F32
    out = Var(K.act_mean_hw_fwd(x.data, act))
    if _rg(x):
        out.requires_grad = True

        def bwd():
            if out.grad is not None:
                x.accum(K.act_mean_hw_bwd(x.data, out.grad, act, x.gdtype))
        _tape().record(bwd)
    return out


def embedding(table: Variable, labels: torch.Tensor) -> Var:
    """tf.nn.embedding_lookup (common/ops/embedding.py:51)."""
    vocab, dim = table.data.shape
    n = labels.numel()
    out = Var(K.embedding_fwd(table.data, labels, n, dim))
    if table.needs_grad and _tape() is not None:
        out.requires_grad = True

        def bwd():
            if out.grad is not None:
                K.embedding_bwd(out.grad, labels, n, dim, vocab, table.grad)
        _tape().record(bwd)
    return out


def tile_hw(e: Var, h: int, w: int) -> Var:
    """tf.tile(e[:, None, None, :], [1, h, w, 1]) for fp32 e [n, c]; the gradient is the sum over (h, w)."""
    n, c = e.shape
    out = Var(e.data.reshape(n, 1, 1, c).expand(n, h, w, c).contiguous())      # tensor plumbing, no arithmetic
    if _rg(e):
        out.requires_grad = True

        def bwd():
            if out.grad is not None:
                g = out.grad if out.grad.dtype == F32 else K.cast(out.grad, F32)
                e.accum(K.cast(K.act_mean_hw_fwd(g, None), F32, scale=float(h * w)))
        _tape().record(bwd)
    return out


def concat_label_map(x: Var, e: Var, act="relu"):
    """tf.concat([x, tile(e[:,None,None,:])], axis=3) (SNGAN/gan_cifar_resnet.py:282-284), emitted directly as the
    two bf16 operands the next residual block needs: the raw concat (shortcut) and act(concat) (Conv1 input)."""
    n, h, w, c1 = x.shape
    c2 = e.shape[1]
    ct = c1 + c2
    dev = x.data.device
    if get_store().is_tf32():
        # TF32 operand mode: fp32 operands, assembled from the generic ops (concat + stand-alone activation)
        x32 = x if x.data.dtype == F32 else cast(x, F32)
        raw32 = concat_channels(x32, tile_hw(e, h, w), out_dtype=F32)
        act32, _ = norm_act(raw32, stats=None, act=act, out_dtype=F32)
        return raw32, act32
    raw = torch.empty((n, h, w, ct), dtype=BF16, device=dev)
    actv = torch.empty((n, h, w, ct), dtype=BF16, device=dev)
    K.norm_act_fwd(x.data, n, h, w, c1, None, None, 1, None, None, None, act, False, BF16, out=actv, out_cstride=ct,
                   out_raw=raw, raw_cstride=ct)
    K.bcast_channels_fwd(e.data, n, h * w, c2, c1, ct, act, raw, actv)
    raw_v, act_v = Var(raw), Var(actv)
    if _rg(x, e):
        raw_v.requires_grad = act_v.requires_grad = True

        def bwd():
            d_raw, d_act = raw_v.grad, act_v.grad
            if d_raw is None and d_act is None:
                return
            if d_raw is not None and d_act is not None and d_raw.dtype != d_act.dtype:
                d_raw, d_act = K.cast(d_raw, F32), K.cast(d_act, F32)
            if e.requires_grad:
                e.accum(K.bcast_channels_bwd(e.data, n, h * w, c2, c1, ct, act, d_raw, d_act))
            if x.requires_grad:
                x.accum(K.concat_bwd_x(x.data, n * h * w, c1, ct, act, d_raw, d_act, x.gdtype))
        _tape().record(bwd)
    return raw_v, act_v


# ------------------------------------------------------------------------------------------------ PGGAN / Pix2Pix
def pixel_norm(x: Var, eps: float = 1e-8, act=None, out_dtype=None) -> Var:
    """act(x * rsqrt(mean_c(x^2) + eps)) -- common/ops/normalization.py:125-140 (+ the lrelu that follows it in
    PGGAN/model_nvidia.py:63-68) as one kernel."""
    out = Var(K.pixel_norm_fwd(x.data, eps, act, out_dtype or x.data.dtype))
    if _rg(x):
        out.requires_grad = True

        def bwd():
            if out.grad is not None:
                x.accum(K.pixel_norm_bwd(x.data, out.grad, eps, act, x.gdtype))
        _tape().record(bwd)
    return out


def minibatch_std(x: Var) -> Var:
    """tf.concat([x, tile(mean(sqrt(var_batch(x) + 1e-8)))], axis=3) -- PGGAN/model_nvidia.py:20-28."""
    xin = x if x.data.dtype == F32 else cast(x, F32)
    store = get_store()
    sync = store.bn_sync if (store.bn_sync is not None and hasattr(store.bn_sync, "allreduce")) else None
    if sync is not None:
        # data-parallel run with synced batch statistics: the standard deviation is taken over the GLOBAL batch
        # ([sum x | sum x^2] per position all-reduced forward, the scalar g backward; SURVEY 8(e) collective 3)
        key = f"{store.scope_name()}/mbstd/{'x'.join(str(d) for d in xin.shape)}"
        y, ws = K.minibatch_std_sync_fwd(xin.data, sync, key)
    else:
        y, ws = K.minibatch_std_fwd(xin.data)
    out = Var(y)
    if _rg(xin):
        out.requires_grad = True

        def bwd():
            if out.grad is not None:
                g = out.grad if out.grad.dtype == F32 else K.cast(out.grad, F32)
                if sync is not None:
                    xin.accum(K.minibatch_std_sync_bwd(xin.data, g, ws, sync, key))
                else:
                    xin.accum(K.minibatch_std_bwd(xin.data, g, ws))
        _tape().record(bwd)
    return out


def concat_channels(a: Var, b: Var, out_dtype=None) -> Var:
    """tf.concat([a, b], axis=3) (U-Net skip connections, Pix2Pix/networks.py:268-270)."""
    ca, cb = a.shape[-1], b.shape[-1]
    assert a.shape[:-1] == b.shape[:-1]
    dtype = out_dtype or a.data.dtype
    y = torch.empty(tuple(a.shape[:-1]) + (ca + cb,), dtype=dtype, device=a.data.device)
    K.copy_channels(a.data, 0, y, 0, ca)
    K.copy_channels(b.data, 0, y, ca, cb)
    out = Var(y)
    if _rg(a, b):
        out.requires_grad = True

        def bwd():
            g = out.grad
            if g is None:
                return
            for v, off, c in ((a, 0, ca), (b, ca, cb)):
                if v.requires_grad:
                    dv = torch.empty(v.data.shape, dtype=v.gdtype, device=g.device)
                    K.copy_channels(g, off, dv, 0, c)
                    v.accum(dv)
        _tape().record(bwd)
    return out


def dropout(x: Var, keep_mask: torch.Tensor, keep_prob: float) -> Var:
    """tf.nn.dropout with an explicit fp32 keep mask in {0, 1} (Pix2Pix/networks.py:263-264): x * mask / keep_prob.
    The mask is an input because TF's op-level RNG cannot be reproduced."""
    c = x.shape[-1]
    y = torch.empty_like(x.data)
    K.copy_channels(x.data, 0, y, 0, c, mask=keep_mask, scale=1.0 / keep_prob)
    out = Var(y, grad_dtype=x.grad_dtype)
    if _rg(x):
        out.requires_grad = True

        def bwd():
            if out.grad is not None:
                dx = torch.empty(x.data.shape, dtype=x.gdtype, device=y.device)
                K.copy_channels(out.grad, 0, dx, 0, c, mask=keep_mask, scale=1.0 / keep_prob)
                x.accum(dx)
        _tape().record(bwd)
    return out


LOSS_TYPES = {"HINGE": 0, "WGAN": 1, "WGAN-GP": 1, "LSGAN": 2, "CGAN": 3, "Modified_MiniMax": 4, "MiniMax": 5}


def gan_loss(logits: Var, mode: str, n_real: int = 0, scale: float = 1.0, loss_out: torch.Tensor | None = None,
             loss_type: str = "HINGE"):
    """mode 'hinge_d' / 'd': discriminator loss over logits = [disc_real (n_real) | disc_fake]; 'gen' / 'g': generator
    loss over disc_fake -- lib.misc.get_loss (common/misc.py:310-394) for every loss_type it knows; the default is the
    hinge pair of gan_cifar_resnet.py:376-378, 492.  Returns the device scalar (accumulated into loss_out when given)."""
    if loss_type not in LOSS_TYPES:
        raise ValueError(f"unknown loss_type {loss_type!r}")
    side = {"hinge_d": 0, "d": 0, "gen": 1, "g": 1}[mode]
    code = 2 * LOSS_TYPES[loss_type] + side
    accumulate = loss_out is not None
    if loss_out is None:
        loss_out = torch.zeros(1, dtype=F32, device=logits.data.device)
    dlogits = K.gan_loss(logits.data.reshape(-1), n_real, code, scale, loss_out, accumulate)
    out = Var(loss_out)
    if _rg(logits):
        out.requires_grad = True

        def bwd():
            logits.accum(dlogits.reshape(logits.data.shape))
        _tape().record(bwd)
    return out


def softmax_xent(logits: Var, labels: torch.Tensor, scale: float = 1.0, loss_out: torch.Tensor | None = None):
    """scale * tf.reduce_mean(tf.nn.sparse_softmax_cross_entropy_with_logits(logits, labels)) (ACGAN/train.py:110-121)."""
    assert logits.data.dtype == F32 and logits.data.dim() == 2
    accumulate = loss_out is not None
    if loss_out is None:
        loss_out = torch.zeros(1, dtype=F32, device=logits.data.device)
    dlogits = K.softmax_xent(logits.data, labels, scale, loss_out, accumulate)
    out = Var(loss_out)
    if _rg(logits):
        out.requires_grad = True

        def bwd():
            logits.accum(dlogits)
        _tape().record(bwd)
    return out


def l1_loss(targets: torch.Tensor, outputs: Var, scale: float = 1.0, loss_out: torch.Tensor | None = None):
    """scale * tf.reduce_mean(tf.abs(targets - outputs)) (Pix2Pix/train.py:511); `targets` is a constant."""
    assert outputs.data.dtype == F32
    accumulate = loss_out is not None
    if loss_out is None:
        loss_out = torch.zeros(1, dtype=F32, device=outputs.data.device)
    d = K.l1_loss(targets, outputs.data, scale, loss_out, accumulate)
    out = Var(loss_out)
    if _rg(outputs):
        out.requires_grad = True

        def bwd():
            outputs.accum(d)
        _tape().record(bwd)
    return out


# ------------------------------------------------------------------------------------------------ second order
# The WGAN-GP penalty (ACGAN/train.py:97-105) is a function of g = d sum(D(x_hat)) / d x_hat and is differentiated
# w.r.t. D's parameters, i.e. through the backward pass of D.  The ops below are the steps of that backward pass written
# as differentiable forward ops ("*_input_grad"): each computes the input gradient of one layer with the same kernels
# the tape uses, and records the vector-Jacobian product of THAT computation.  Convolution and pooling backward are
# linear (their VJPs are the layers' own forward / filter-gradient kernels), activation masks are piecewise constant;
# only the backward of a training-mode batch norm depends on the forward activations (ganb_bn_bwd_vjp).
def _f32(t: torch.Tensor) -> torch.Tensor:
    return t if t.dtype == F32 else K.cast(t, F32)


def conv2d_input_grad(gy: Var, W: Variable, x_shape, kh: int, kw: int, padding: str = "SAME") -> Var:
    """gx = d<gy, conv2d(x, W)>/dx for a stride-1 convolution without spectral norm (tf.nn.conv2d's
    Conv2DBackpropInput), differentiable w.r.t. gy and W."""
    store = get_store()
    n, h, w, cin = x_shape
    cout = W.data.shape[-1]
    taps = kh * kw
    pt, pl, ho, wo = _pads(padding, h, w, kh, kw, 1)
    if cout % 8 or (cin % 8 and not (cin < 8 and small_k(taps, cin) is not None)):
        raise NotImplementedError(f"conv2d_input_grad: cin={cin}, cout={cout}")
    small_in = cin < 8
    kp_in = small_k(taps, cin) if small_in else None
    group = store.pack_group(W.root)
    pack = group.entry(W)
    group.refresh()
    gy16 = gy.data if gy.data.dtype == BF16 else K.cast(gy.data, BF16)
    gx = K.conv_igemm(gy16, pack.wn, n, ho, wo, cout, h, w, cin, kh, kw, kh - 1 - pt, kw - 1 - pl, True, None, None,
                      None, None, F32)
    out = Var(gx, grad_dtype=F32)
    need_w = W.needs_grad and _tape() is not None
    if _rg(gy) or need_w:
        out.requires_grad = True

        def bwd():
            c = out.grad
            if c is None:
                return
            c32 = _f32(c)
            if small_in:
                ccol = K.im2col_small(c32, n, h, w, cin, ho, wo, kh, kw, pt, pl, +1, kp_in)
                if gy.requires_grad:   # forward convolution of the cotangent
                    gy.accum(K.conv_igemm(ccol, pack.ws, n, ho, wo, kp_in, ho, wo, cout, 1, 1, 0, 0, False, None, None,
                                          None, None, gy.gdtype))
                if need_w:
                    r = torch.empty((kp_in, cout), dtype=F32, device=c.device)
                    K.conv_wgrad(ccol, gy16, r, n, ho, wo, kp_in, ho, wo, cout, 1, 1, 0, 0, None, 0.0)
                    K.small_wgrad_scatter(r, W.grad, taps, cin, cout, False, None, 1.0)
            else:
                c16 = K.cast(c32, BF16)
                if gy.requires_grad:
                    gy.accum(K.conv_igemm(c16, pack.wt, n, h, w, cin, ho, wo, cout, kh, kw, pt, pl, False, None, None,
                                          None, None, gy.gdtype))
                if need_w:             # gx is linear in W: dW = wgrad with the cotangent in the activation's place
                    K.conv_wgrad(c16, gy16, W.grad, n, h, w, cin, ho, wo, cout, kh, kw, pt, pl, None, 1.0)
        _tape().record(bwd)
    return out


def act_input_grad(x: Var, gy: Var, act) -> Var:
    """gx = gy * act'(x) (backward of a stand-alone relu / leaky relu on fp32 x); the mask is piecewise constant."""
    n, h, w, c = x.shape
    gx = K.norm_act_bwd(_f32(x.data), _f32(gy.data), 0, n, h, w, c, None, None, 1, None, None, None, act, False, None,
                        None, None, F32)
    out = Var(gx, grad_dtype=F32)
    if _rg(gy):
        out.requires_grad = True

        def bwd():
            if out.grad is not None:
                gy.accum(K.norm_act_bwd(_f32(x.data), _f32(out.grad), 0, n, h, w, c, None, None, 1, None, None, None, act,
                                        False, None, None, None, gy.gdtype))
        _tape().record(bwd)
    return out


def bn_act_input_grad(x: Var, gy: Var, gamma: Variable, beta: Variable, mean: torch.Tensor, rstd: torch.Tensor, act,
                      eps: float = 1e-5) -> Var:
    """gx = d<gy, act(batch_norm(x))>/dx in training mode (one statistic group) for fp32 x; differentiable w.r.t. x
    (the statistics and x_hat), gy and gamma -- the grad-grad of fused batch norm (ganb_bn_bwd_vjp)."""
    n, h, w, c = x.shape
    assert x.data.dtype == F32
    gy32 = _f32(gy.data)
    gx = K.norm_act_bwd(x.data, gy32, 0, n, h, w, c, mean, rstd, 1, gamma.data, beta.data, None, act, False, None, None,
                        None, F32)
    out = Var(gx, grad_dtype=F32)
    need_p = gamma.needs_grad and _tape() is not None
    if _rg(x, gy) or need_p:
        out.requires_grad = True

        def bwd():
            c_ = out.grad
            if c_ is None:
                return
            dx, dgy = K.bn_bwd_vjp(x.data, gy32, _f32(c_), mean.reshape(-1), rstd.reshape(-1), gamma.data, beta.data, act,
                                   gamma.grad if need_p else None)
            if x.requires_grad:
                x.accum(dx if x.gdtype == F32 else K.cast(dx, x.gdtype))
            if gy.requires_grad:
                gy.accum(dgy if gy.gdtype == F32 else K.cast(dgy, gy.gdtype))
        _tape().record(bwd)
    return out


def meanpool2_input_grad(gy: Var) -> Var:
    """gx[n, 2h, 2w, c] = gy / 4 replicated 2x2 (backward of the 2x2 mean pool)."""
    out = Var(K.expand2(_f32(gy.data), 0.25, F32), grad_dtype=F32)
    if _rg(gy):
        out.requires_grad = True

        def bwd():
            if out.grad is not None:
                gy.accum(K.sum2x2(_f32(out.grad), 0.25, gy.gdtype))
        _tape().record(bwd)
    return out


def act_mean_hw_input_grad(x: Var, gout: Var, act) -> Var:
    """gx = act'(x) * gout[n, None, None, :] / (h w) (backward of mean over (h, w) of act(x)), fp32 x."""
    n, h, w, c = x.shape
    out = Var(K.act_mean_hw_bwd(_f32(x.data), _f32(gout.data), act, F32), grad_dtype=F32)
    if _rg(gout):
        out.requires_grad = True

        def bwd():
            if out.grad is None:
                return
            masked = K.norm_act_bwd(_f32(x.data), _f32(out.grad), 0, n, h, w, c, None, None, 1, None, None, None, act,
                                    False, None, None, None, F32)
            gout.accum(K.act_mean_hw_fwd(masked, None))
        _tape().record(bwd)
    return out


def linear_input_grad(gout: torch.Tensor, W: Variable) -> Var:
    """g[m, in] = gout[m, out] @ W^T for a constant upstream gradient (the ones of tf.gradients), differentiable w.r.t. W."""
    m, kout = gout.shape
    kin = W.data.shape[0]
    g = torch.empty((m, kin), dtype=F32, device=gout.device)
    K.sgemm_small(gout, W.data, g, m, kin, kout, False, True, None, None, 0.0)
    out = Var(g, grad_dtype=F32)
    if W.needs_grad and _tape() is not None:
        out.requires_grad = True

        def bwd():
            if out.grad is not None:      # dW[in, out] += c^T @ gout
                K.sgemm_small(_f32(out.grad), gout, W.grad, kin, kout, m, True, False, None, None, 1.0)
        _tape().record(bwd)
    return out


def gradient_penalty_loss(g: Var, scale: float = 10.0, loss_out: torch.Tensor | None = None) -> Var:
    """scale * mean_n (sqrt(sum_{hwc} g^2 + 1e-10) - 1)^2 (ACGAN/train.py:102-104)."""
    accumulate = loss_out is not None
    if loss_out is None:
        loss_out = torch.zeros(1, dtype=F32, device=g.data.device)
    n = g.shape[0]
    dg = K.gp_loss(_f32(g.data).reshape(n, -1), scale, loss_out, accumulate)
    out = Var(loss_out)
    if _rg(g):
        out.requires_grad = True

        def bwd():
            g.accum(dg.reshape(g.data.shape))
        _tape().record(bwd)
    return out
