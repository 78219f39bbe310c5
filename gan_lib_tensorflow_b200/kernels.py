"""Thin tensor-level wrappers over the C ABI (include/ganb200.h).

torch is used for device memory (caching allocator), dtypes and streams only; every arithmetic step is a
libganb200 kernel.  All activations are contiguous NHWC.  No wrapper has a CPU or PyTorch fallback.
"""
from __future__ import annotations

import contextlib
import ctypes
from ctypes import c_float, c_int, c_int64, c_void_p

import torch

from . import cabi
from .cabi import BF16, F32, act_code, check, ptr

_L = None


_TEST_DOUBLE = False


def install_test_double(lib) -> None:
    """Test hook (tests/hostlogic.py): replaces the C-ABI library by an object that records / ignores calls, so that
    variable naming, RNG order and tape wiring can be exercised on a machine without a GPU.  The double itself lives in
    tests/; nothing in the product selects it -- without this call every entry point goes to libganb200.so and
    cabi.lib() raises when the library is missing.  install_test_double(None) restores the real library."""
    global _L, _TEST_DOUBLE
    _L = lib
    _TEST_DOUBLE = lib is not None


def host_logic_only() -> bool:
    """True while a test double is installed (no device, no streams)."""
    return _TEST_DOUBLE


def L():
    global _L
    if _L is None:
        _L = cabi.lib()
        for name in ("ganb_launch_count", "ganb_conv2d_wgrad_workspace", "ganb_upconv_wgrad_workspace",
                     "ganb_norm_act_bwd_sums_offset", "ganb_l1_loss_workspace", "ganb_bn_bwd_vjp_workspace", "ganb_conv2d_small_wgrad_workspace", "ganb_bn_stats_workspace",
                     "ganb_minibatch_std_workspace", "ganb_weight_transform_workspace", "ganb_layer_norm_workspace",
                     "ganb_layer_norm_rows", "ganb_depthwise_conv2d_chunks",
                     "ganb_norm_act_bwd_workspace", "ganb_colsum_workspace"):
            getattr(_L, name).restype = c_int64
    return _L


def _stream():
    if host_logic_only():
        return c_void_p(0)
    return c_void_p(torch.cuda.current_stream().cuda_stream)


@contextlib.contextmanager
def sm_limit(sms: int):
    """Tensor-core kernels issued inside occupy at most `sms` SMs (0 / None: no limit); see ganb_set_sm_limit."""
    if not sms or host_logic_only():
        yield
        return
    prev = L().ganb_set_sm_limit(int(sms))
    try:
        yield
    finally:
        L().ganb_set_sm_limit(prev)


def launch_count() -> int:
    """Kernels launched by libganb200 so far in this process."""
    return int(L().ganb_launch_count())


def dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _tdtype(code: int):
    return torch.bfloat16 if code == BF16 else torch.float32


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def same_pads(in_size: int, k: int, s: int):
    """TF SAME padding: (before, after, out)."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    return total // 2, total - total // 2, out


# ------------------------------------------------------------------------------------------------ casts
def cast(x: torch.Tensor, dtype, scale: float = 1.0) -> torch.Tensor:
    y = torch.empty(x.shape, dtype=dtype, device=x.device)
    check(L().ganb_cast(ptr(x), dt(x), ptr(y), dt(y), c_int64(x.numel()), c_float(scale), _stream()), "ganb_cast")
    return y


def cast_into(x: torch.Tensor, out: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """out = scale * x with the dtype of `out` (persistent destination: the bf16 wire copy of a gradient buffer)."""
    assert x.numel() == out.numel()
    check(L().ganb_cast(ptr(x), dt(x), ptr(out), dt(out), c_int64(x.numel()), c_float(scale), _stream()), "ganb_cast")
    return out


def axpby(x: torch.Tensor, y: torch.Tensor, a: float = 1.0, b: float = 1.0) -> None:
    """y = a*x + b*y (fp32)."""
    assert x.dtype == torch.float32 and y.dtype == torch.float32 and x.numel() == y.numel()
    check(L().ganb_axpby(ptr(x), ptr(y), c_int64(x.numel()), c_float(a), c_float(b), _stream()), "ganb_axpby")


# ------------------------------------------------------------------------------------------------ conv (TC)
def conv_igemm(x, wp, n, h, w, cin, ho, wo, cout, kh, kw, pad_t, pad_l, flip, alpha, bias, residual, act, out_dtype,
               residual_up2=False, stride=1):
    y = torch.empty((n, ho, wo, cout), dtype=out_dtype, device=x.device)
    check(L().ganb_conv2d_igemm(ptr(x), ptr(wp), ptr(y), n, h, w, cin, ho, wo, cout, kh, kw, stride, pad_t, pad_l,
                                int(flip), ptr(alpha), ptr(bias), ptr(residual), int(bool(residual_up2)), act_code(act),
                                BF16 if out_dtype == torch.bfloat16 else F32, _stream()), "ganb_conv2d_igemm")
    return y


def conv_igemm_gated(x, wp, n, h, w, cin, ho, wo, cout, kh, kw, pad_t, pad_l, flip, alpha, gate, gate_act, out_dtype):
    """conv_igemm whose result is multiplied by act'(pre), read off gate = act(pre) (bf16, shape of the output): the data
    gradient through an activation that the producer of the layer's input fused into its epilogue."""
    assert gate.dtype == torch.bfloat16 and tuple(gate.shape) == (n, ho, wo, cout), (gate.dtype, gate.shape)
    y = torch.empty((n, ho, wo, cout), dtype=out_dtype, device=x.device)
    check(L().ganb_conv2d_igemm_gated(ptr(x), ptr(wp), ptr(y), n, h, w, cin, ho, wo, cout, kh, kw, 1, pad_t, pad_l,
                                      int(flip), ptr(alpha), ptr(gate), act_code(gate_act),
                                      BF16 if out_dtype == torch.bfloat16 else F32, _stream()), "ganb_conv2d_igemm_gated")
    return y


class FusedStats:
    """Per-tile column sums (y, y^2) left by a convolution epilogue for the batch norm behind the layer."""

    __slots__ = ("partial", "rows", "groups", "c")

    def __init__(self, partial, rows, groups, c):
        self.partial, self.rows, self.groups, self.c = partial, rows, groups, c

    def finalize(self, count_per_group: int, eps: float):
        """-> (mean, rstd) [groups, c], the values ganb_bn_stats would return."""
        dev = self.partial.device
        mean = torch.empty((self.groups, self.c), dtype=torch.float32, device=dev)
        rstd = torch.empty((self.groups, self.c), dtype=torch.float32, device=dev)
        check(L().ganb_bn_stats_finalize(ptr(self.partial), self.c, self.groups, self.rows, c_int64(count_per_group),
                                         c_float(eps), ptr(mean), ptr(rstd), _stream()), "ganb_bn_stats_finalize")
        return mean, rstd


def conv_stats_rows(n, ho, wo, cout, kh, kw, stride, groups) -> int:
    return int(L().ganb_conv2d_stats_rows(n, ho, wo, cout, kh, kw, stride, groups))


def conv_igemm_stats(x, wp, n, h, w, cin, ho, wo, cout, kh, kw, pad_t, pad_l, flip, alpha, bias, residual, act,
                     out_dtype, groups, residual_up2=False, stride=1):
    """conv_igemm that also returns the FusedStats of its (stored) output for `groups` statistic towers."""
    rows = conv_stats_rows(n, ho, wo, cout, kh, kw, stride, groups)
    y = torch.empty((n, ho, wo, cout), dtype=out_dtype, device=x.device)
    partial = torch.empty((groups, rows, 2, cout), dtype=torch.float32, device=x.device)
    check(L().ganb_conv2d_igemm_stats(ptr(x), ptr(wp), ptr(y), n, h, w, cin, ho, wo, cout, kh, kw, stride, pad_t, pad_l,
                                      int(flip), ptr(alpha), ptr(bias), ptr(residual), int(bool(residual_up2)),
                                      act_code(act), BF16 if out_dtype == torch.bfloat16 else F32, ptr(partial), groups,
                                      _stream()), "ganb_conv2d_igemm_stats")
    return y, FusedStats(partial, rows, groups, cout)


def conv_wgrad(x, dy, dw, n, h, w, cin, ho, wo, cout, kh, kw, pad_t, pad_l, scale, beta, stride=1):
    nbytes = L().ganb_conv2d_wgrad_workspace(n, h, w, cin, ho, wo, cout, kh, kw)
    ws = _ws(nbytes, x.device)
    check(L().ganb_conv2d_wgrad(ptr(x), ptr(dy), ptr(dw), ptr(ws), n, h, w, cin, ho, wo, cout, kh, kw, stride, pad_t, pad_l,
                                ptr(scale), c_float(beta), _stream()), "ganb_conv2d_wgrad")


# ------------------------------------------------------------------------------------------------ conv (TC, TF32 operands)
def round_tf32(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> nearest TF32 value (fp32 storage)."""
    assert x.dtype == torch.float32
    y = torch.empty_like(x)
    check(L().ganb_round_tf32(ptr(x), ptr(y), c_int64(x.numel()), _stream()), "ganb_round_tf32")
    return y


def transpose_tf32(w_hwio: torch.Tensor, taps: int, cin: int, cout: int) -> torch.Tensor:
    """HWIO fp32 filter [taps, cin, cout] -> [taps, cout, cin], rounded to TF32 (the fprop operand copy)."""
    wt = torch.empty((taps, cout, cin), dtype=torch.float32, device=w_hwio.device)
    check(L().ganb_transpose_tf32(ptr(w_hwio), ptr(wt), taps, cin, cout, _stream()), "ganb_transpose_tf32")
    return wt


def conv_igemm_tf32(x, wp, n, h, w, cin, ho, wo, cout, kh, kw, pad_t, pad_l, flip, alpha, bias, residual, act, out_dtype,
                    residual_up2=False, stride=1):
    assert x.dtype == torch.float32 and wp.dtype == torch.float32
    y = torch.empty((n, ho, wo, cout), dtype=out_dtype, device=x.device)
    check(L().ganb_conv2d_igemm_tf32(ptr(x), ptr(wp), ptr(y), n, h, w, cin, ho, wo, cout, kh, kw, stride, pad_t, pad_l,
                                     int(flip), ptr(alpha), ptr(bias), ptr(residual), int(bool(residual_up2)),
                                     act_code(act), BF16 if out_dtype == torch.bfloat16 else F32, _stream()),
          "ganb_conv2d_igemm_tf32")
    return y


def conv_wgrad_tf32(x, dy, dw, n, h, w, cin, ho, wo, cout, kh, kw, pad_t, pad_l, scale, beta, stride=1):
    assert x.dtype == torch.float32 and dy.dtype == torch.float32
    fn = L().ganb_conv2d_wgrad_tf32_workspace
    fn.restype = c_int64
    ws = _ws(fn(n, ho, wo, cin, cout, kh, kw), x.device)
    check(L().ganb_conv2d_wgrad_tf32(ptr(x), ptr(dy), ptr(dw), ptr(ws), n, h, w, cin, ho, wo, cout, kh, kw, stride, pad_t,
                                     pad_l, ptr(scale), c_float(beta), _stream()), "ganb_conv2d_wgrad_tf32")


# ------------------------------------------------------------------------------------------------ sub-pixel UpsampleConv
def upconv_supported(n, h, w, cin, cout) -> bool:
    return bool(L().ganb_upconv_supported(n, h, w, cin, cout))


def upconv_pack(w, we_t, we_n, cin, cout):
    check(L().ganb_upconv_pack(ptr(w), ptr(we_t), ptr(we_n), cin, cout, _stream()), "ganb_upconv_pack")


def upconv_fprop(x, we_t, n, h, w, cin, cout, alpha, bias, act, out_dtype):
    """-> quad-layout tensor, allocated with the logical NHWC shape [n, 2h, 2w, cout]."""
    y = torch.empty((n, 2 * h, 2 * w, cout), dtype=out_dtype, device=x.device)
    check(L().ganb_upconv_fprop(ptr(x), ptr(we_t), ptr(y), n, h, w, cin, cout, ptr(alpha), ptr(bias), act_code(act),
                                BF16 if out_dtype == torch.bfloat16 else F32, _stream()), "ganb_upconv_fprop")
    return y


def upconv_stats_rows(n, h, w, cin, cout, groups) -> int:
    return int(L().ganb_upconv_stats_rows(n, h, w, cin, cout, groups))


def upconv_fprop_stats(x, we_t, n, h, w, cin, cout, alpha, bias, act, out_dtype, groups):
    rows = upconv_stats_rows(n, h, w, cin, cout, groups)
    y = torch.empty((n, 2 * h, 2 * w, cout), dtype=out_dtype, device=x.device)
    partial = torch.empty((groups, rows, 2, cout), dtype=torch.float32, device=x.device)
    check(L().ganb_upconv_fprop_stats(ptr(x), ptr(we_t), ptr(y), n, h, w, cin, cout, ptr(alpha), ptr(bias),
                                      act_code(act), BF16 if out_dtype == torch.bfloat16 else F32, ptr(partial), groups,
                                      _stream()), "ganb_upconv_fprop_stats")
    return y, FusedStats(partial, rows, groups, cout)


def upconv_dgrad(dy_quad, we_n, n, h, w, cin, cout, alpha, out_dtype):
    dx = torch.empty((n, h, w, cin), dtype=out_dtype, device=dy_quad.device)
    check(L().ganb_upconv_dgrad(ptr(dy_quad), ptr(we_n), ptr(dx), n, h, w, cin, cout, ptr(alpha),
                                BF16 if out_dtype == torch.bfloat16 else F32, _stream()), "ganb_upconv_dgrad")
    return dx


def upconv_wgrad(x, dy_quad, dw, n, h, w, cin, cout, scale, beta):
    ws = _ws(L().ganb_upconv_wgrad_workspace(n, h, w, cin, cout), x.device)
    check(L().ganb_upconv_wgrad(ptr(x), ptr(dy_quad), ptr(dw), ptr(ws), n, h, w, cin, cout, ptr(scale), c_float(beta),
                                _stream()), "ganb_upconv_wgrad")


# ------------------------------------------------------------------------------------------------ conv (small)
def conv_smallcin(x, w, n, h, w_in, cs, ho, wo, cl, kh, kw, pad_t, pad_l, flip, w_clcs, alpha, bias, act, out_dtype):
    y = torch.empty((n, ho, wo, cl), dtype=out_dtype, device=x.device)
    check(L().ganb_conv2d_smallcin(ptr(x), ptr(w), ptr(y), n, h, w_in, cs, ho, wo, cl, kh, kw, pad_t, pad_l,
                                   int(flip), int(w_clcs), ptr(alpha), ptr(bias), act_code(act),
                                   BF16 if out_dtype == torch.bfloat16 else F32, _stream()), "ganb_conv2d_smallcin")
    return y


def conv_small_wgrad(xs, yl, dw, n, hs, ws_, cs, hl, wl, cl, kh, kw, pad_t, pad_l, sign, out_clcs, scale, beta):
    nbytes = L().ganb_conv2d_small_wgrad_workspace(n, hl, wl, cs, cl, kh, kw)
    ws = _ws(nbytes, xs.device)
    check(L().ganb_conv2d_small_wgrad(ptr(xs), ptr(yl), dt(yl), ptr(dw), ptr(ws), n, hs, ws_, cs, hl, wl, cl, kh, kw,
                                      pad_t, pad_l, sign, int(out_clcs), ptr(scale), c_float(beta), _stream()),
          "ganb_conv2d_small_wgrad")


def im2col_small(xs, n, hs, ws_, cs, ho, wo, kh, kw, pad_t, pad_l, sign, kpad=32, stride=1):
    out = torch.empty((n, ho, wo, kpad), dtype=torch.bfloat16, device=xs.device)
    check(L().ganb_im2col_small(ptr(xs), ptr(out), n, hs, ws_, cs, ho, wo, kh, kw, stride, pad_t, pad_l, sign, kpad,
                                _stream()),
          "ganb_im2col_small")
    return out


def pack_small(w, out, taps, ci, co, small_is_ci, kpad=32):
    check(L().ganb_pack_small(ptr(w), ptr(out), taps, ci, co, int(bool(small_is_ci)), kpad, _stream()),
          "ganb_pack_small")


def small_wgrad_scatter(r, dw, taps, cs, cl, out_clcs, scale, beta):
    check(L().ganb_small_wgrad_scatter(ptr(r), ptr(dw), taps, cs, cl, int(bool(out_clcs)), ptr(scale), c_float(beta),
                                       _stream()), "ganb_small_wgrad_scatter")


def sgemm_small(a, b, c, m, n, k, trans_a, trans_b, alpha=None, bias=None, beta=0.0):
    check(L().ganb_sgemm_small(ptr(a), ptr(b), ptr(c), m, n, k, int(trans_a), int(trans_b), ptr(alpha), ptr(bias),
                               c_float(beta), _stream()), "ganb_sgemm_small")


# ------------------------------------------------------------------------------------------------ norm / act
def bn_stats(x, n, hw, c, groups, eps):
    mean = torch.empty((groups, c), dtype=torch.float32, device=x.device)
    rstd = torch.empty((groups, c), dtype=torch.float32, device=x.device)
    ws = _ws(L().ganb_bn_stats_workspace(n, hw, c, groups), x.device)
    check(L().ganb_bn_stats(ptr(x), dt(x), n, hw, c, groups, c_float(eps), ptr(mean), ptr(rstd), ptr(ws), _stream()),
          "ganb_bn_stats")
    return mean, rstd


def norm_act_fwd(x, n, h, w, c, mean, rstd, groups, gamma, beta, labels, act, upsample, out_dtype, out=None,
                 out_cstride=0, out_raw=None, raw_cstride=0):
    if out is None:
        s = 2 if int(upsample) == 1 else 1     # upsample == 2: quad-layout INPUT, same resolution
        out = torch.empty((n, s * h, s * w, c), dtype=out_dtype, device=x.device)
    check(L().ganb_norm_act_fwd(ptr(x), dt(x), n, h, w, c, ptr(mean), ptr(rstd), groups, ptr(gamma), ptr(beta), ptr(labels),
                                act_code(act), int(upsample), ptr(out), dt(out), out_cstride, ptr(out_raw),
                                raw_cstride, _stream()), "ganb_norm_act_fwd")
    return out


def _is_peer(sync) -> bool:
    """sync is a peer.PeerComm-like object (key-addressed exchanges) rather than the (allreduce_sum, world) pair."""
    return sync is not None and hasattr(sync, "bn_moments")


def norm_act_bwd(x, dz, dz_cstride, n, h, w, c, mean, rstd, groups, gamma, beta, labels, act, upsample, dgamma, dbeta,
                 add, dx_dtype, sync=None, sync_key=None):
    """sync: cross-GPU batch statistics -- the per-group sums of the local share are all-reduced between the reduction
    and the apply kernels.  Either (allreduce_sum(tensor) -> None, world) (library collective, eager mode) or a
    peer.PeerComm (one peer-memory kernel, capturable; sync_key names the call site)."""
    dx = torch.empty((n, h, w, c), dtype=dx_dtype, device=x.device)
    ws = None
    if mean is not None:
        ws = _ws(L().ganb_norm_act_bwd_workspace(n, h * w, c, groups), x.device)
    n_rows = int(gamma.shape[0]) if (gamma is not None and gamma.dim() == 2) else 1
    args = (ptr(x), dt(x), ptr(dz), dt(dz), dz_cstride, n, h, w, c, ptr(mean), ptr(rstd), groups,
            ptr(gamma), ptr(beta), ptr(labels), n_rows, act_code(act), int(upsample), ptr(dgamma),
            ptr(dbeta), ptr(add), dt(add) if add is not None else F32, ptr(dx), dt(dx), ptr(ws))
    if sync is None or mean is None:
        check(L().ganb_norm_act_bwd(*args, _stream()), "ganb_norm_act_bwd")
        return dx
    check(L().ganb_norm_act_bwd_phase(*args, 1, c_float(1.0), _stream()), "ganb_norm_act_bwd_phase")
    off = int(L().ganb_norm_act_bwd_sums_offset(n, h * w, c, groups))
    sums = ws[off:off + 2 * groups * c * 4].view(torch.float32)
    if _is_peer(sync):
        sync.allreduce(sync_key + "/bwd", sums)
        world = sync.world
    else:
        allreduce, world = sync
        allreduce(sums)
    check(L().ganb_norm_act_bwd_phase(*args, 2, c_float(1.0 / world), _stream()), "ganb_norm_act_bwd_phase")
    return dx


def bn_stats_sync(mean, rstd, eps, sync, sync_key=None):
    """Replaces the local (mean, rstd) [groups, c] by the statistics over all ranks (equal shares per rank)."""
    if _is_peer(sync):
        sync.bn_moments(sync_key + "/fwd", mean, rstd, eps)     # pack + exchange + unpack in one peer-memory kernel
        return
    allreduce, world = sync
    count = mean.numel()
    buf = torch.empty(2 * count, dtype=torch.float32, device=mean.device)
    check(L().ganb_bn_moments_pack(ptr(mean), ptr(rstd), count, c_float(eps), ptr(buf), _stream()), "ganb_bn_moments_pack")
    allreduce(buf)
    check(L().ganb_bn_moments_unpack(ptr(buf), count, c_float(1.0 / world), c_float(eps), ptr(mean), ptr(rstd), _stream()),
          "ganb_bn_moments_unpack")


def meanpool2(x, add, out_dtype):
    n, h, w, c = x.shape
    out = torch.empty((n, h // 2, w // 2, c), dtype=out_dtype, device=x.device)
    check(L().ganb_meanpool2_fwd(ptr(x), dt(x), ptr(add), ptr(out), dt(out), n, h, w, c, _stream()),
          "ganb_meanpool2_fwd")
    return out


def expand2(x, scale, out_dtype):
    n, h, w, c = x.shape
    out = torch.empty((n, 2 * h, 2 * w, c), dtype=out_dtype, device=x.device)
    check(L().ganb_expand2(ptr(x), dt(x), ptr(out), dt(out), n, h, w, c, c_float(scale), _stream()), "ganb_expand2")
    return out


def dilate2d(x, stride, out_h, out_w, out_dtype=None):
    """out[n, i*stride, j*stride, c] = x[n, i, j, c], zero elsewhere."""
    n, h, w, c = x.shape
    out = torch.empty((n, out_h, out_w, c), dtype=out_dtype or x.dtype, device=x.device)
    check(L().ganb_dilate2d(ptr(x), dt(x), ptr(out), dt(out), n, h, w, c, stride, out_h, out_w, _stream()),
          "ganb_dilate2d")
    return out


def sum2x2(x, scale, out_dtype):
    n, h, w, c = x.shape
    out = torch.empty((n, h // 2, w // 2, c), dtype=out_dtype, device=x.device)
    check(L().ganb_sum2x2(ptr(x), dt(x), ptr(out), dt(out), n, h, w, c, c_float(scale), _stream()), "ganb_sum2x2")
    return out


def colsum(x2d, rows, c, out, beta):
    ws = _ws(L().ganb_colsum_workspace(c_int64(rows), c), x2d.device)
    check(L().ganb_colsum(ptr(x2d), dt(x2d), c_int64(rows), c, c_float(beta), ptr(out), ptr(ws), _stream()),
          "ganb_colsum")


def bcast_channels_fwd(e, n, hw, c2, coff, cstride, act, out_raw, out_act):
    check(L().ganb_bcast_channels_fwd(ptr(e), n, hw, c2, coff, cstride, act_code(act), ptr(out_raw), ptr(out_act),
                                      _stream()), "ganb_bcast_channels_fwd")


def bcast_channels_bwd(e, n, hw, c2, coff, cstride, act, d_raw, d_act):
    de = torch.empty((n, c2), dtype=torch.float32, device=e.device)
    gd = dt(d_raw if d_raw is not None else d_act)
    check(L().ganb_bcast_channels_bwd(ptr(e), n, hw, c2, coff, cstride, act_code(act), ptr(d_raw), ptr(d_act), gd,
                                      ptr(de), _stream()), "ganb_bcast_channels_bwd")
    return de


def concat_bwd_x(x, pixels, c1, cstride, act, d_raw, d_act, dx_dtype=torch.float32):
    dx = torch.empty(x.shape, dtype=dx_dtype, device=x.device)
    gd = dt(d_raw if d_raw is not None else d_act)
    check(L().ganb_concat_bwd_x(ptr(x), c_int64(pixels), c1, cstride, act_code(act), ptr(d_raw), ptr(d_act), gd,
                                ptr(dx), dt(dx), _stream()), "ganb_concat_bwd_x")
    return dx


def act_mean_hw_fwd(x, act):
    n, h, w, c = x.shape
    out = torch.empty((n, c), dtype=torch.float32, device=x.device)
    check(L().ganb_act_mean_hw_fwd(ptr(x), n, h * w, c, act_code(act), ptr(out), _stream()), "ganb_act_mean_hw_fwd")
    return out


def act_mean_hw_bwd(x, dout, act, dx_dtype=torch.float32):
    n, h, w, c = x.shape
    dx = torch.empty(x.shape, dtype=dx_dtype, device=x.device)
    check(L().ganb_act_mean_hw_bwd(ptr(x), ptr(dout), n, h * w, c, act_code(act), ptr(dx), dt(dx), _stream()),
          "ganb_act_mean_hw_bwd")
    return dx


def gan_loss(logits, n_real, mode, scale, loss_out, accumulate):
    n = logits.numel()
    dlogits = torch.empty_like(logits)
    check(L().ganb_gan_loss(ptr(logits), n, n_real, mode, c_float(scale), int(accumulate), ptr(loss_out),
                            ptr(dlogits), _stream()), "ganb_gan_loss")
    return dlogits


def l1_loss(targets, outputs, scale, loss_out, accumulate):
    assert targets.dtype == torch.float32 and outputs.dtype == torch.float32 and targets.numel() == outputs.numel()
    d = torch.empty_like(outputs)
    ws = _ws(L().ganb_l1_loss_workspace(c_int64(outputs.numel())), outputs.device)
    check(L().ganb_l1_loss(ptr(targets), ptr(outputs), c_int64(outputs.numel()), c_float(scale), int(accumulate),
                           ptr(loss_out), ptr(d), ptr(ws), _stream()), "ganb_l1_loss")
    return d


# ------------------------------------------------------------------------------------------------ gradient penalty
def interpolate(real, fake, alpha):
    n = real.shape[0]
    out = torch.empty_like(real)
    check(L().ganb_interpolate(ptr(real), ptr(fake), ptr(alpha), n, c_int64(real.numel() // n), ptr(out), _stream()),
          "ganb_interpolate")
    return out


def gp_loss(g, scale, loss_out, accumulate):
    n = g.shape[0]
    dg = torch.empty_like(g)
    ws = torch.empty(n, dtype=torch.float32, device=g.device)
    check(L().ganb_gp_loss(ptr(g), n, c_int64(g.numel() // n), c_float(scale), int(accumulate), ptr(loss_out), ptr(dg),
                           ptr(ws), _stream()), "ganb_gp_loss")
    return dg


def bn_bwd_vjp(x, gy, cot, mean, rstd, gamma, beta, act, dgamma):
    """-> (d/dx, d/dgy) of sum(cot * BNbwd(x, gy)); d/dgamma accumulated into `dgamma` (may be None)."""
    c = x.shape[-1]
    pixels = x.numel() // c
    dx, dgy = torch.empty_like(x), torch.empty_like(gy)
    ws = _ws(L().ganb_bn_bwd_vjp_workspace(c_int64(pixels), c), x.device)
    check(L().ganb_bn_bwd_vjp(ptr(x), ptr(gy), ptr(cot), ptr(mean), ptr(rstd), ptr(gamma), ptr(beta), c_int64(pixels), c,
                              act_code(act), ptr(dx), ptr(dgy), ptr(dgamma), ptr(ws), _stream()), "ganb_bn_bwd_vjp")
    return dx, dgy


def softmax_xent(logits, labels, scale, loss_out, accumulate):
    n, c = logits.shape
    dlogits = torch.empty_like(logits)
    check(L().ganb_softmax_xent(ptr(logits), ptr(labels), n, c, c_float(scale), int(accumulate), ptr(loss_out),
                                ptr(dlogits), _stream()), "ganb_softmax_xent")
    return dlogits


def adam(params, grads, m, v, lr_t, beta1, beta2, eps, grad_scale=1.0):
    check(L().ganb_adam(ptr(params), ptr(grads), ptr(m), ptr(v), c_int64(params.numel()), ptr(lr_t), c_float(beta1),
                        c_float(beta2), c_float(eps), c_float(grad_scale), _stream()), "ganb_adam")


def preprocess_real(data_int, noise, b, hw):
    out = torch.empty((b, hw * 3), dtype=torch.float32, device=data_int.device)
    check(L().ganb_preprocess_real(ptr(data_int), ptr(noise), b, hw, ptr(out), _stream()), "ganb_preprocess_real")
    return out


def embedding_fwd(table, labels, n, dim):
    out = torch.empty((n, dim), dtype=torch.float32, device=table.device)
    check(L().ganb_embedding_fwd(ptr(table), ptr(labels), n, dim, ptr(out), _stream()), "ganb_embedding_fwd")
    return out


def embedding_bwd(dout, labels, n, dim, vocab, dtable):
    check(L().ganb_embedding_bwd(ptr(dout), ptr(labels), n, dim, vocab, ptr(dtable), _stream()), "ganb_embedding_bwd")


# ------------------------------------------------------------------------------------------------ PGGAN / Pix2Pix extras
def pixel_norm_fwd(x, eps, act, out_dtype):
    c = x.shape[-1]
    y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    check(L().ganb_pixel_norm_fwd(ptr(x), dt(x), ptr(y), dt(y), c_int64(x.numel() // c), c, c_float(eps), act_code(act),
                                  _stream()), "ganb_pixel_norm_fwd")
    return y


def pixel_norm_bwd(x, dy, eps, act, dx_dtype):
    c = x.shape[-1]
    dx = torch.empty(x.shape, dtype=dx_dtype, device=x.device)
    check(L().ganb_pixel_norm_bwd(ptr(x), dt(x), ptr(dy), dt(dy), ptr(dx), dt(dx), c_int64(x.numel() // c), c,
                                  c_float(eps), act_code(act), _stream()), "ganb_pixel_norm_bwd")
    return dx


def minibatch_std_fwd(x, cs=None):
    b, h, w, c = x.shape
    cs = cs or (c + 1)
    out = torch.empty((b, h, w, cs), dtype=torch.float32, device=x.device)
    ws = _ws(L().ganb_minibatch_std_workspace(b, h, w, c), x.device)
    check(L().ganb_minibatch_std_fwd(ptr(x), b, h, w, c, cs, ptr(out), ptr(ws), _stream()), "ganb_minibatch_std_fwd")
    return out, ws


def minibatch_std_bwd(x, dout, ws):
    b, h, w, c = x.shape
    cs = dout.shape[-1]
    dx = torch.empty_like(x)
    check(L().ganb_minibatch_std_bwd(ptr(x), ptr(dout), b, h, w, c, cs, ptr(dx), ptr(ws), _stream()),
          "ganb_minibatch_std_bwd")
    return dx


def minibatch_std_sync_fwd(x, sync, key, cs=None):
    """minibatch_std over the GLOBAL batch of a data-parallel run: `sync` is a peer.PeerComm / peer.NcclSync."""
    b, h, w, c = x.shape
    cs = cs or (c + 1)
    m = h * w * c
    fn = L().ganb_minibatch_std_sync_workspace
    fn.restype = c_int64
    ws = torch.empty(int(fn(h, w, c)) // 4, dtype=torch.float32, device=x.device)
    out = torch.empty((b, h, w, cs), dtype=torch.float32, device=x.device)
    args = (ptr(x), b, h, w, c, cs, ptr(out), ptr(ws))
    check(L().ganb_minibatch_std_sync_fwd(*args, 1, sync.world, _stream()), "ganb_minibatch_std_sync_fwd")
    sync.allreduce(key + "/fwd", ws[:2 * m])
    check(L().ganb_minibatch_std_sync_fwd(*args, 2, sync.world, _stream()), "ganb_minibatch_std_sync_fwd")
    return out, ws


def minibatch_std_sync_bwd(x, dout, ws, sync, key):
    b, h, w, c = x.shape
    cs = dout.shape[-1]
    dx = torch.empty_like(x)
    fn = L().ganb_minibatch_std_sync_g_offset
    fn.restype = c_int64
    off = int(fn(h, w, c))
    args = (ptr(x), ptr(dout), b, h, w, c, cs, ptr(dx), ptr(ws))
    check(L().ganb_minibatch_std_sync_bwd(*args, 1, sync.world, _stream()), "ganb_minibatch_std_sync_bwd")
    sync.allreduce(key + "/bwd", ws[off:off + 1])
    check(L().ganb_minibatch_std_sync_bwd(*args, 2, sync.world, _stream()), "ganb_minibatch_std_sync_bwd")
    return dx


def copy_channels(src, src_off, dst, dst_off, c, mask=None, scale=1.0):
    """dst[..., dst_off:dst_off+c] = scale * mask * src[..., src_off:src_off+c] (NHWC, any leading dims)."""
    pixels = src.numel() // src.shape[-1]
    check(L().ganb_copy_channels(ptr(src), dt(src), src.shape[-1], src_off, ptr(dst), dt(dst), dst.shape[-1], dst_off,
                                 c_int64(pixels), c, ptr(mask), c_float(scale), _stream()), "ganb_copy_channels")


# ------------------------------------------------------------------------------------------------ grouped SN / pack
SN_ROWS = 16   # GANB_SN_ROWS of include/ganb200.h (checked by tests/test_host_logic.py)


class SnLayerStruct(ctypes.Structure):
    _fields_ = [("w", c_void_p), ("u", c_void_p), ("u_out", c_void_p), ("u_used", c_void_p), ("v", c_void_p),
                ("b", c_void_p), ("scal", c_void_p), ("g", c_void_p), ("dw", c_void_p), ("t", c_void_p),
                ("work", c_void_p), ("k", ctypes.c_int32), ("c", ctypes.c_int32), ("blk_begin", ctypes.c_int32),
                ("pad_", ctypes.c_int32)]


class PackLayerStruct(ctypes.Structure):
    _fields_ = [("w", c_void_p), ("wn", c_void_p), ("wt", c_void_p), ("taps", ctypes.c_int32), ("ci", ctypes.c_int32),
                ("co", ctypes.c_int32), ("tile_begin", ctypes.c_int32), ("ci_pad", ctypes.c_int32),
                ("pad_", ctypes.c_int32)]


def struct_array_to_device(items, device) -> torch.Tensor:
    """Copies a list of ctypes structures into a device byte tensor (the descriptor table of a grouped launch)."""
    arr = (type(items[0]) * len(items))(*items)
    raw = bytes(arr)
    return torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)


def sn_power_iter(table_dev, count, total_blocks, max_c, assign):
    check(L().ganb_sn_power_iter(ptr(table_dev), count, total_blocks, max_c, int(bool(assign)), _stream()),
          "ganb_sn_power_iter")


def sn_bwd(table_dev, count, total_blocks, max_c):
    check(L().ganb_sn_bwd(ptr(table_dev), count, total_blocks, max_c, _stream()), "ganb_sn_bwd")


def pack_weights(table_dev, count, total_tiles):
    check(L().ganb_pack_weights(ptr(table_dev), count, total_tiles, _stream()), "ganb_pack_weights")


# ------------------------------------------------------------------------------------------------ secondary variants
def weight_transform_fwd(w, g, mask, w_eff, norms, a, c, b):
    """w_eff = w * (g / ||w||) * mask over geometry [a][c][b] (include/ganb200.h); g / mask may be None."""
    ws = _ws(L().ganb_weight_transform_workspace(a, c), w.device) if g is not None else None
    check(L().ganb_weight_transform_fwd(ptr(w), ptr(g), ptr(mask), ptr(w_eff), ptr(norms), ptr(ws), a, c, b, _stream()),
          "ganb_weight_transform_fwd")


def weight_transform_bwd(w, dw_eff, g, mask, norms, dw, dg, a, c, b):
    """Adds the gradients of weight_transform_fwd into dw (and dg)."""
    ws = _ws(L().ganb_weight_transform_workspace(a, c), w.device) if g is not None else None
    check(L().ganb_weight_transform_bwd(ptr(w), ptr(dw_eff), ptr(g), ptr(mask), ptr(norms), ptr(dw), ptr(dg), ptr(ws),
                                        a, c, b, _stream()), "ganb_weight_transform_bwd")


def layer_norm_fwd(x, gamma, beta, eps, act, out_dtype):
    """Per-sample moments over (h, w, c), y = act(batch_normalization(x, mean, var, beta, gamma, eps)).
    Returns (y, mean_rstd [n, 2])."""
    n, c = x.shape[0], x.shape[-1]
    per = x.numel() // n
    y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    mr = torch.empty((n, 2), dtype=torch.float32, device=x.device)
    ws = _ws(L().ganb_layer_norm_workspace(n, c_int64(per), c), x.device)
    check(L().ganb_layer_norm_fwd(ptr(x), dt(x), ptr(gamma), ptr(beta), ptr(y), dt(y), ptr(mr), ptr(ws), n, c_int64(per),
                                  c, c_float(eps), act_code(act), _stream()), "ganb_layer_norm_fwd")
    return y, mr


def layer_norm_bwd(x, dy, mean_rstd, gamma, beta, act, dx_dtype, dgamma, dbeta):
    """Returns dx (None when dx_dtype is None); adds into dgamma / dbeta (either may be None)."""
    n, c = x.shape[0], x.shape[-1]
    per = x.numel() // n
    rows = int(L().ganb_layer_norm_rows(n, c_int64(per)))
    dx = torch.empty(x.shape, dtype=dx_dtype, device=x.device) if dx_dtype is not None else None
    chan = torch.empty((2, rows, c), dtype=torch.float32, device=x.device)
    ws = _ws(L().ganb_layer_norm_workspace(n, c_int64(per), c), x.device)
    check(L().ganb_layer_norm_bwd(ptr(x), dt(x), ptr(dy), dt(dy), ptr(mean_rstd), ptr(gamma), ptr(beta), ptr(dx),
                                  dt(dx) if dx is not None else F32, ptr(chan), ptr(ws), n, c_int64(per), c,
                                  act_code(act), _stream()), "ganb_layer_norm_bwd")
    if dgamma is not None:
        colsum(chan[0], rows, c, dgamma, 1.0)
    if dbeta is not None:
        colsum(chan[1], rows, c, dbeta, 1.0)
    return dx


def scale_dev(x, alpha_dev):
    """alpha * x (fp32) with alpha a 1-element device tensor."""
    assert x.dtype == torch.float32
    y = torch.empty_like(x)
    check(L().ganb_scale_dev(ptr(x), ptr(alpha_dev), ptr(y), c_int64(x.numel()), _stream()), "ganb_scale_dev")
    return y


def lerp_fwd(a, b, alpha_dev):
    """(1 - alpha) * a + alpha * b, fp32, alpha a device scalar."""
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.shape == b.shape
    y = torch.empty_like(a)
    check(L().ganb_lerp_fwd(ptr(a), ptr(b), ptr(y), c_int64(a.numel()), ptr(alpha_dev), _stream()), "ganb_lerp_fwd")
    return y


def lerp_bwd(dy, alpha_dev, da_dtype, db_dtype):
    """((1 - alpha) * dy, alpha * dy); a dtype of None skips that output."""
    assert dy.dtype == torch.float32
    da = torch.empty(dy.shape, dtype=da_dtype, device=dy.device) if da_dtype is not None else None
    db = torch.empty(dy.shape, dtype=db_dtype, device=dy.device) if db_dtype is not None else None
    check(L().ganb_lerp_bwd(ptr(dy), ptr(da), dt(da) if da is not None else F32, ptr(db),
                            dt(db) if db is not None else F32, c_int64(dy.numel()), ptr(alpha_dev), _stream()),
          "ganb_lerp_bwd")
    return da, db


def subsample2d(x, stride, out_dtype=None):
    """y[n, i, j, c] = x[n, i*stride, j*stride, c] (tf.image.resize_nearest_neighbor to 1/stride of the size)."""
    n, h, w, c = x.shape
    y = torch.empty((n, -(-h // stride), -(-w // stride), c), dtype=out_dtype or x.dtype, device=x.device)
    check(L().ganb_subsample2d(ptr(x), dt(x), ptr(y), dt(y), n, h, w, c, stride, 0, _stream()), "ganb_subsample2d")
    return y


def subsample2d_bwd(dy, stride, h, w, out_dtype=None):
    """Gradient of subsample2d: [n, h, w, c] with dy on the sampled grid and zeros elsewhere."""
    n, _, _, c = dy.shape
    dx = torch.empty((n, h, w, c), dtype=out_dtype or dy.dtype, device=dy.device)
    check(L().ganb_subsample2d(ptr(dy), dt(dy), ptr(dx), dt(dx), n, h, w, c, stride, 1, _stream()), "ganb_subsample2d")
    return dx


def sample_grid(samples, nw):
    """[n, h, w, c] samples in (-1, 1) -> uint8 grid [ceil(n / nw) * h, nw * w, c] (common/misc.py:215-244)."""
    n, h, w, c = samples.shape
    grid = torch.empty((-(-n // nw) * h, nw * w, c), dtype=torch.uint8, device=samples.device)
    check(L().ganb_sample_grid(ptr(samples), dt(samples), n, h, w, c, nw, ptr(grid), _stream()), "ganb_sample_grid")
    return grid


def depthwise_fwd(x, filt, bias, ho, wo, stride, pad_t, pad_l, out_dtype):
    """tf.nn.depthwise_conv2d (+ bias): x [n,h,w,c], filt [kh,kw,c,cm] fp32 -> [n,ho,wo,c*cm]."""
    n, h, w, c = x.shape
    kh, kw, _, cm = filt.shape
    y = torch.empty((n, ho, wo, c * cm), dtype=out_dtype, device=x.device)
    check(L().ganb_depthwise_conv2d_fwd(ptr(x), dt(x), ptr(filt), ptr(bias), ptr(y), dt(y), n, h, w, c, cm, ho, wo, kh,
                                        kw, stride, pad_t, pad_l, _stream()), "ganb_depthwise_conv2d_fwd")
    return y


def depthwise_bwd_input(dy, filt, h, w, stride, pad_t, pad_l, out_dtype):
    n, ho, wo, _ = dy.shape
    kh, kw, c, cm = filt.shape
    dx = torch.empty((n, h, w, c), dtype=out_dtype, device=dy.device)
    check(L().ganb_depthwise_conv2d_bwd_input(ptr(dy), dt(dy), ptr(filt), ptr(dx), dt(dx), n, h, w, c, cm, ho, wo, kh, kw,
                                              stride, pad_t, pad_l, _stream()), "ganb_depthwise_conv2d_bwd_input")
    return dx


def depthwise_bwd_filter(x, dy, dfilt, kh, kw, cm, stride, pad_t, pad_l):
    """Adds the filter gradient into dfilt [kh,kw,c,cm]."""
    n, h, w, c = x.shape
    _, ho, wo, _ = dy.shape
    chunks = int(L().ganb_depthwise_conv2d_chunks(n, ho, wo))
    cols = kh * kw * c * cm
    partials = torch.empty((chunks, cols), dtype=torch.float32, device=x.device)
    check(L().ganb_depthwise_conv2d_bwd_filter(ptr(x), dt(x), ptr(dy), dt(dy), ptr(partials), n, h, w, c, cm, ho, wo, kh,
                                               kw, stride, pad_t, pad_l, _stream()), "ganb_depthwise_conv2d_bwd_filter")
    colsum(partials, chunks, cols, dfilt, 1.0)
