"""GPU tests of the fusions into the tensor-core convolution epilogue (through the C ABI): the batch statistics a
convolution leaves for the batch norm behind it must equal ganb_bn_stats over the stored output, for every kernel
variant that produces them (CTA-pair, single-CTA halo, per-tap igemm with several images per tile, sub-pixel UpsampleConv)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BF16 = torch.bfloat16
F32 = torch.float32


def _operands(n, h, w, cin, cout, k, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(n, h, w, cin, device="cuda", generator=g).to(BF16)
    wt = (torch.randn(k * k, cout, cin, device="cuda", generator=g) * (1.5 / (k * np.sqrt(cin)))).to(BF16)
    bias = torch.randn(cout, device="cuda", generator=g) * 0.3
    return x, wt, bias


@pytest.mark.parametrize("n,h,cin,cout,k,groups,out_dtype,res", [
    (128, 32, 256, 256, 3, 2, BF16, None),      # conv_pair_kernel<256>: the dominant layer of the headline
    (64, 32, 128, 128, 3, 2, BF16, "full"),     # conv_pair_kernel<128> + fp32 residual
    (16, 16, 256, 256, 3, 2, BF16, "up2"),      # single-CTA halo kernel, residual read through the nearest upsample
    (128, 8, 1024, 256, 3, 2, BF16, None),      # per-tap igemm, two images per 128-pixel tile
    (64, 4, 256, 128, 3, 1, F32, None),         # igemm, eight images per tile, fp32 output, one tower
    (6, 16, 64, 96, 1, 3, BF16, None),          # 1x1, Cout = 96 (partial channel tile), three towers
    (5, 20, 64, 64, 3, 1, BF16, None),          # ragged image size: tile rows outside the tensor contribute nothing
])
def test_conv_epilogue_statistics_equal_bn_stats(n, h, cin, cout, k, groups, out_dtype, res):
    from gan_lib_tensorflow_b200 import kernels as K

    x, wt, bias = _operands(n, h, h, cin, cout, k, seed=n + h)
    pad = k // 2
    residual = None
    if res == "full":
        residual = torch.randn(n, h, h, cout, device="cuda")
    elif res == "up2":
        residual = torch.randn(n, h // 2, h // 2, cout, device="cuda")
    assert K.conv_stats_rows(n, h, h, cout, k, k, 1, groups) > 0
    y, fused = K.conv_igemm_stats(x, wt, n, h, h, cin, h, h, cout, k, k, pad, pad, False, None, bias, residual, "relu",
                                  out_dtype, groups, residual_up2=(res == "up2"))
    y_ref = K.conv_igemm(x, wt, n, h, h, cin, h, h, cout, k, k, pad, pad, False, None, bias, residual, "relu", out_dtype,
                         residual_up2=(res == "up2"))
    assert torch.equal(y, y_ref)                           # the statistics do not disturb the output
    eps = 1e-5
    mean, rstd = fused.finalize((n // groups) * h * h, eps)
    mean_ref, rstd_ref = K.bn_stats(y, n, h * h, cout, groups, eps)
    torch.cuda.synchronize()
    # against fp64 moments of the stored tensor
    yd = y.double().reshape(groups, -1, cout)
    m64 = yd.mean(dim=1)
    v64 = yd.var(dim=1, unbiased=False)
    r64 = 1.0 / torch.sqrt(v64 + eps)
    std = torch.sqrt(v64 + eps)
    assert float(((mean.double() - m64).abs() / std).max()) < 2e-6
    assert float(((rstd.double() - r64).abs() / r64).max()) < 2e-5
    assert float(((mean - mean_ref).abs() / std.float()).max()) < 4e-6
    assert float(((rstd - rstd_ref).abs() / rstd_ref).max()) < 4e-5


def test_conv_epilogue_statistics_refuse_straddling_tiles():
    from gan_lib_tensorflow_b200 import kernels as K

    assert K.conv_stats_rows(6, 8, 8, 128, 3, 3, 1, 2) == 0     # 3 images per tower, 2 images per 128-pixel tile
    assert K.conv_stats_rows(8, 8, 8, 24, 3, 3, 1, 2) == 0      # Cout not a multiple of 32
    assert K.conv_stats_rows(8, 8, 8, 128, 3, 3, 1, 2) == 2


@pytest.mark.parametrize("n,h,c,groups", [(64, 16, 256, 2), (128, 16, 256, 2), (16, 32, 128, 1)])
def test_upconv_epilogue_statistics_equal_bn_stats(n, h, c, groups):
    from gan_lib_tensorflow_b200 import kernels as K

    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(n, h, h, c, device="cuda", generator=g).to(BF16)
    w = torch.randn(3, 3, c, c, device="cuda", generator=g) * (0.5 / np.sqrt(c))
    we_t = torch.empty(16, c, c, dtype=BF16, device="cuda")
    we_n = torch.empty(16, c, c, dtype=BF16, device="cuda")
    K.upconv_pack(w, we_t, we_n, c, c)
    bias = torch.randn(c, device="cuda", generator=g) * 0.2
    assert K.upconv_stats_rows(n, h, h, c, c, groups) > 0
    y, fused = K.upconv_fprop_stats(x, we_t, n, h, h, c, c, None, bias, None, BF16, groups)
    y_ref = K.upconv_fprop(x, we_t, n, h, h, c, c, None, bias, None, BF16)
    assert torch.equal(y, y_ref)
    eps = 1e-5
    mean, rstd = fused.finalize((n // groups) * 4 * h * h, eps)
    torch.cuda.synchronize()
    yd = y.double().reshape(groups, -1, c)          # quad layout is a pixel permutation inside each image
    m64, v64 = yd.mean(dim=1), yd.var(dim=1, unbiased=False)
    std = torch.sqrt(v64 + eps)
    assert float(((mean.double() - m64).abs() / std).max()) < 2e-6
    assert float(((rstd.double() - 1.0 / std).abs() * std).max()) < 2e-5


def test_generator_with_fused_statistics_matches_the_separate_pass():
    """SNGAN-CIFAR generator forward + a generator step: the fused statistics change the summation order of the moments
    only (fp32 per-tile sums folded in fp64 instead of per-thread fp32 sums), so outputs agree to rounding."""
    from gan_lib_tensorflow_b200 import framework
    from gan_lib_tensorflow_b200 import functional as F
    from gan_lib_tensorflow_b200.SNGAN import gan_cifar_resnet as P

    outs = []
    for fused in (False, True):
        F.FUSED_BN_STATS = fused
        try:
            store = framework.reset_default_graph("cuda", u_seed=2)
            tr = P.Trainer(batch_size=64, seed=0)
            rs = np.random.RandomState(3)
            tr.z_g.copy_(torch.from_numpy(rs.standard_normal((128, 128)).astype("float32")))
            tr.fake_labels.copy_(torch.from_numpy(rs.randint(0, 10, size=128).astype("int32")))
            before = framework.K.launch_count()
            tr.gen_opt.set_lr(0.0)
            tr._g_body()
            torch.cuda.synchronize()
            outs.append((tr.g_loss.item(), store.flat["Generator"].grads.clone(), framework.K.launch_count() - before))
        finally:
            F.FUSED_BN_STATS = True
            framework.set_store(None)
    (l0, g0, n0), (l1, g1, n1) = outs
    assert n1 == n0 - 6                                 # six bn_stats_partial launches fewer per generator pass
    assert abs(l0 - l1) < 1e-3 * max(1.0, abs(l0))
    assert float((g0 - g1).norm() / g0.norm()) < 2e-2   # bf16 rounding flips downstream of 1e-6 moment differences


# ------------------------------------------------------------------------------------------------ activation fused into a conv pair
@pytest.mark.parametrize("n,h,cin,cout,k,act,out_dtype", [
    (128, 32, 128, 128, 3, "relu", BF16),      # conv_pair_kernel<128>: D.Block.1.Conv2's data gradient
    (64, 16, 128, 256, 3, "relu", BF16),       # 256-wide output (label-map block of the critic)
    (128, 8, 128, 128, 3, "relu", BF16),       # per-tap igemm, 64-wide tiles
    (16, 8, 64, 72, 3, "lrelu", BF16),         # partial channel tile, leaky relu
    (8, 16, 32, 128, 1, "lrelu", F32),         # 1x1 filter (shallow-K route), fp32 output
])
def test_gated_data_gradient_equals_gradient_times_activation_derivative(n, h, cin, cout, k, act, out_dtype):
    """ganb_conv2d_igemm_gated(dy, W, gate) == ganb_conv2d_igemm(dy, W) * act'(pre), act' read off gate = act(pre)."""
    from gan_lib_tensorflow_b200 import kernels as K

    x, wt, _ = _operands(n, h, h, cin, cout, k, seed=3 * n + h)
    g = torch.Generator(device="cuda").manual_seed(11)
    pre = torch.randn(n, h, h, cout, device="cuda", generator=g)
    pre[0, 0, 0, :8] = 0.0                                    # exact zeros: relu'(0) = 0, lrelu'(0) = 1
    pre[0, 0, 1, :8] = -0.0
    gate = (torch.relu(pre) if act == "relu" else torch.where(pre >= 0, pre, 0.2 * pre)).to(BF16)
    pad = k // 2
    plain = K.conv_igemm(x, wt, n, h, h, cin, h, h, cout, k, k, pad, pad, True, None, None, None, None, F32)
    gated = K.conv_igemm_gated(x, wt, n, h, h, cin, h, h, cout, k, k, pad, pad, True, None, gate, act, out_dtype)
    torch.cuda.synchronize()
    gf = gate.float()
    if act == "relu":
        want = torch.where(gf > 0, plain, torch.zeros_like(plain))
    else:
        want = torch.where(gf >= 0, plain, 0.2 * plain)
    assert torch.equal(gated, want.to(out_dtype))


@pytest.mark.parametrize("resample,cin,cout,h", [(None, 128, 128, 8), ("down", 256, 128, 16), ("first", 3, 128, 32)])
def test_fused_block_activation_is_bit_identical_to_separate_passes(resample, cin, cout, h):
    """A critic residual block with the Conv1 -> relu -> Conv2 activation fused into the two epilogues gives the same
    bits (output, input gradient, every parameter gradient) as with the separate activation passes."""
    from gan_lib_tensorflow_b200 import framework, functional as F
    from gan_lib_tensorflow_b200.common import resnet_block as rb

    def run(fused):
        old = F.FUSED_CONV_ACT
        F.FUSED_CONV_ACT = fused
        try:
            framework.reset_default_graph("cuda")
            store = framework.get_store()
            np.random.seed(5)
            xs = torch.from_numpy(np.random.RandomState(1).randn(32, h, h, cin).astype("float32")).cuda()
            launches = {}
            with store.variable_scope("Discriminator"):
                with store.gradient_tape() as tape:
                    x = F.Var(xs, requires_grad=True)
                    from gan_lib_tensorflow_b200 import kernels as K
                    before = K.launch_count()
                    if resample == "first":
                        y = rb.OptimizedResBlockDisc1(x, DIM_D=cout, spectral_normed=True, name_prefix="D.Block.1")
                    else:
                        y = rb.ResidualBlock(x, cin, cout, 3, "D.Block", spectral_normed=True, resample=resample)
                    for v in store.vars.values():
                        if v.trainable and v.grad is None:
                            v.grad = torch.zeros_like(v.data)
                    cot = torch.from_numpy(np.random.RandomState(2).randn(*y.shape).astype("float32")).cuda()
                    tape.backward(y, grad=cot.to(y.gdtype))
                    launches["n"] = K.launch_count() - before
            torch.cuda.synchronize()
            grads = {k: v.grad.clone() for k, v in store.vars.items() if v.trainable}
            return y.data.clone(), x.grad.clone(), grads, launches["n"]
        finally:
            F.FUSED_CONV_ACT = old

    y1, dx1, g1, n1 = run(True)
    y0, dx0, g0, n0 = run(False)
    assert n1 == n0 - 2                       # one forward and one backward pass fewer
    assert torch.equal(y1, y0) and torch.equal(dx1, dx0)
    assert g1.keys() == g0.keys() and len(g1) >= 4
    for k in g1:
        assert torch.equal(g1[k], g0[k]), k
