"""CPU tier: self-checks that pin the oracle (the reference ships no tests or golden vectors).

Each check compares the oracle with an INDEPENDENT statement of the same TF-1.x semantics: a NumPy loop
convolution with explicit SAME padding, numpy.linalg.svd, finite differences in float64, closed-form identities
and a hand-computed Adam vector.  The committed golden vectors under tests/golden/ freeze the oracle's outputs.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ops as O
from oracle import resnet_block as RB
from oracle import sngan_cifar as S
from oracle import tfshim

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def naive_conv2d_same(x, w, stride):
    """Loop convolution with the TF SAME rule written out independently: out = ceil(in/s),
    pad_total = max((out-1)*s + k - in, 0), pad_before = pad_total // 2 (the extra pixel goes after)."""
    n, h, wd, cin = x.shape
    kh, kw, _, cout = w.shape
    oh, ow = -(-h // stride), -(-wd // stride)
    pth = max((oh - 1) * stride + kh - h, 0)
    ptw = max((ow - 1) * stride + kw - wd, 0)
    pt, pl = pth // 2, ptw // 2
    y = np.zeros((n, oh, ow, cout))
    for i in range(oh):
        for j in range(ow):
            for r in range(kh):
                for s in range(kw):
                    hi, wi = i * stride + r - pt, j * stride + s - pl
                    if 0 <= hi < h and 0 <= wi < wd:
                        y[:, i, j, :] += x[:, hi, wi, :] @ w[r, s]
    return y


@pytest.mark.parametrize("h,k,stride", [(8, 3, 1), (8, 4, 1), (8, 3, 2), (7, 3, 2), (8, 4, 2), (5, 1, 1)])
def test_conv_same_padding_matches_loop_conv(h, k, stride):
    rs = np.random.RandomState(0)
    x = rs.standard_normal((2, h, h + 1, 3))
    w = rs.standard_normal((k, k, 3, 4))
    got = O.conv2d_nhwc(torch.from_numpy(x), torch.from_numpy(w), stride, "SAME").numpy()
    np.testing.assert_allclose(got, naive_conv2d_same(x, w, stride), rtol=1e-10, atol=1e-10)
    # asymmetric cases called out in SURVEY 8(c): 4x4 s1 pads (1,2); 3x3 s2 on even input pads (0,1)
    assert O.same_pads(8, 4, 1)[:2] == (1, 2)
    assert O.same_pads(8, 3, 2)[:2] == (0, 1)


def test_conv_transpose_is_the_input_gradient_of_conv():
    rs = np.random.RandomState(1)
    x = torch.from_numpy(rs.standard_normal((2, 4, 4, 5)))          # deconv input  [N,H,W,Cin]
    filt = torch.from_numpy(rs.standard_normal((4, 4, 3, 5)))       # [k,k,Cout,Cin]
    got = O.conv2d_transpose_nhwc(x, filt, 2, "SAME")
    z = torch.zeros(2, 8, 8, 3, dtype=torch.float64, requires_grad=True)
    y = O.conv2d_nhwc(z, filt, 2, "SAME")                           # [N,4,4,Cin]
    (ref,) = torch.autograd.grad(y, z, x)
    np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=1e-10, atol=1e-10)


def test_sigma_converges_to_largest_singular_value():
    rs = np.random.RandomState(2)
    g = tfshim.Graph(dtype=torch.float64)
    W = torch.from_numpy(rs.standard_normal((3, 3, 16, 24)))
    for _ in range(200):
        _, sigma = O.spectral_normed_weight(g, W, update_collection=None, with_sigma=True)
    s_true = np.linalg.svd(W.numpy().reshape(-1, 24), compute_uv=False)[0]
    assert abs(sigma.item() - s_true) < 1e-6 * s_true
    # NO_OPS leaves u untouched; None assigns (sn.py:48-65)
    u_before = g.vars["spectral_norm/u"].clone()
    O.spectral_normed_weight(g, W, update_collection=O.NO_OPS)
    assert torch.equal(g.vars["spectral_norm/u"], u_before)
    O.spectral_normed_weight(g, W, update_collection="my_ops")
    assert len(g.collections["my_ops"]) == 1 and torch.equal(g.vars["spectral_norm/u"], u_before)


def test_gradient_flows_through_the_power_iteration():
    """no stop_gradient in sn.py. Finite differences in float64 against autograd."""
    rs = np.random.RandomState(3)
    W0 = rs.standard_normal((6, 5)) * 0.3
    G = torch.from_numpy(rs.standard_normal((6, 5)))
    u0 = rs.standard_normal((1, 5))

    def f(wnp):
        g = tfshim.Graph(dtype=torch.float64)
        g.get_variable("spectral_norm/u", initializer=u0, trainable=False)
        W = torch.from_numpy(wnp).requires_grad_(True)
        wb = O.spectral_normed_weight(g, W, update_collection=O.NO_OPS)
        return (wb * G).sum(), W

    loss, W = f(W0.copy())
    (grad,) = torch.autograd.grad(loss, W)
    num = np.zeros_like(W0)
    eps = 1e-6
    for i in range(6):
        for j in range(5):
            wp, wm = W0.copy(), W0.copy()
            wp[i, j] += eps
            wm[i, j] -= eps
            num[i, j] = (f(wp)[0].item() - f(wm)[0].item()) / (2 * eps)
    np.testing.assert_allclose(grad.numpy(), num, rtol=1e-6, atol=1e-8)


def test_cond_batchnorm_with_equal_labels_is_batchnorm_and_uses_population_variance():
    rs = np.random.RandomState(4)
    x = torch.from_numpy(rs.standard_normal((4, 3, 3, 6)) * 2 + 1)
    g = tfshim.Graph(dtype=torch.float64)
    y = O.cond_batchnorm(g, "n", [0, 1, 2], x, labels=torch.zeros(4, dtype=torch.int64), n_labels=10)
    xn = x.numpy()
    mean = xn.mean(axis=(0, 1, 2))
    var = xn.var(axis=(0, 1, 2))  # biased, tf.nn.moments
    np.testing.assert_allclose(y.detach().numpy(), (xn - mean) / np.sqrt(var + 1e-5), rtol=1e-10, atol=1e-10)
    g2 = tfshim.Graph(dtype=torch.float64)
    yb = O.batch_norm(g2, x)
    np.testing.assert_allclose(y.detach().numpy(), yb.detach().numpy(), rtol=1e-10, atol=1e-10)
    with pytest.raises(Exception):
        O.cond_batchnorm(g, "n", [0, 1], x, labels=torch.zeros(4, dtype=torch.int64), n_labels=10)


def test_resampling_identities():
    rs = np.random.RandomState(5)
    x = rs.standard_normal((2, 4, 4, 3))
    up = RB.upsample2(torch.from_numpy(x)).numpy()
    np.testing.assert_array_equal(up, np.repeat(np.repeat(x, 2, axis=1), 2, axis=2))   # depth_to_space(concat x4)
    pooled = RB.mean_pool2(torch.from_numpy(x)).numpy()
    np.testing.assert_allclose(pooled, x.reshape(2, 2, 2, 2, 2, 3).mean(axis=(2, 4)), rtol=1e-12)
    # UpsampleConv == conv(np.repeat); ConvMeanPool == avg-pool(conv) with the same variables
    np.random.seed(0)
    g = tfshim.Graph(dtype=torch.float64)
    y1 = RB.UpsampleConv(g, torch.from_numpy(x), 5, 3, 1, "U")
    y2 = O.conv2d_nhwc(torch.from_numpy(up), g.vars["U/Filters"], 1, "SAME") + g.vars["U/Biases"]
    np.testing.assert_allclose(y1.detach().numpy(), y2.detach().numpy(), rtol=1e-12)
    y3 = RB.ConvMeanPool(g, torch.from_numpy(x), 5, 3, 1, "C")
    y4 = RB.mean_pool2(O.conv2d_nhwc(torch.from_numpy(x), g.vars["C/Filters"], 1, "SAME") + g.vars["C/Biases"])
    np.testing.assert_allclose(y3.detach().numpy(), y4.detach().numpy(), rtol=1e-12)


def test_activation_gradients_at_zero():
    x = torch.tensor([-1.0, 0.0, 2.0], dtype=torch.float64, requires_grad=True)
    (g,) = torch.autograd.grad(RB.nonlinearity(x, "relu").sum(), x)
    assert g.tolist() == [0.0, 0.0, 1.0]                       # tf.nn.relu: slope 0 at exactly 0
    (g,) = torch.autograd.grad(RB.nonlinearity(x, "lrelu").sum(), x)
    assert g.tolist() == [0.2, 1.0, 1.0]                       # tf.maximum(x, 0.2x): MaximumGrad -> first arg at ties
    with pytest.raises(ValueError):
        RB.nonlinearity(x, "swish")


def test_adam_three_steps_by_hand():
    """tf.train.AdamOptimizer with beta1=0, beta2=0.9, eps=1e-8 on a scalar, computed by hand."""
    p = torch.tensor([1.0], dtype=torch.float64)
    opt = S.Adam(0.0, 0.9, 1e-8)
    grads = [0.5, -0.25, 0.125]
    v = 0.0
    expect = 1.0
    for t, gval in enumerate(grads, start=1):
        v = 0.9 * v + 0.1 * gval * gval
        lr_t = 2e-4 * np.sqrt(1 - 0.9 ** t) / (1 - 0.0 ** t)
        expect -= lr_t * gval / (np.sqrt(v) + 1e-8)
        opt.apply([("p", p)], [torch.tensor([gval], dtype=torch.float64)], 2e-4)
    assert abs(p.item() - expect) < 1e-15
    assert S.lr_decay(0) == 1.0 and abs(S.lr_decay(25000) - 0.75) < 1e-12 and S.lr_decay(50000) == 0.5


def test_preprocess_real_layout():
    data = np.arange(2 * 3072, dtype=np.int32).reshape(2, 3072) % 256
    out = S.preprocess_real(torch.from_numpy(data), torch.zeros(2, 3072, dtype=torch.float64), torch.float64).numpy()
    img = data.reshape(2, 3, 32, 32).transpose(0, 2, 3, 1)            # CHW -> HWC (gan_cifar_resnet.py:336-337)
    np.testing.assert_allclose(out.reshape(2, 32, 32, 3), 2 * (img / 256.0 - 0.5), rtol=1e-12)


def test_variable_manifest_and_parameter_counts():
    """names, shapes and totals of SNGAN-CIFAR."""
    np.random.seed(0)
    m = S.SNGANCifar(dtype=torch.float32)
    m.build()
    gen = sum(v.numel() for _, v in m.g.trainable_variables("Generator"))
    disc = sum(v.numel() for _, v in m.g.trainable_variables("Discriminator"))
    assert gen == 7875587 and disc == 1701689
    assert tuple(m.g.vars["Discriminator/D.Block.2.Conv1/filters/spectral_norm/u"].shape) == (1, 256)
    assert tuple(m.g.vars["Discriminator/D.Embedding_y/spectral_norm/u"].shape) == (1, 128)
    assert tuple(m.g.vars["Generator/G.Block.1.N1/CondBatchNorm/scale"].shape) == (10, 1024)
    assert "Discriminator/D.Block.3.Shortcut/Filters" not in m.g.vars        # identity shortcut
    n_u = sum(v.numel() for n, v in m.g.vars.items() if n.endswith("/u"))
    assert n_u == 1537


def test_golden_vectors():
    """Frozen outputs of the oracle on fixed seeded inputs (tests/golden/make_golden.py regenerates them)."""
    path = os.path.join(GOLDEN, "oracle_golden.json")
    with open(path) as fh:
        gold = json.load(fh)
    from tests.golden import make_golden

    now = make_golden.compute()
    for key, val in gold.items():
        np.testing.assert_allclose(np.asarray(now[key]), np.asarray(val), rtol=2e-5, atol=1e-7, err_msg=key)


def test_depthwise_conv_matches_an_explicit_loop():
    """oracle.ops.depthwise_conv2d_nhwc (tf.nn.depthwise_conv2d: output channel ci*cm + m, TF SAME padding) against a
    NumPy loop written from the definition; separable_conv2d = that followed by the 1x1 pointwise convolution."""
    import numpy as np
    import torch

    from oracle import ops as O
    from oracle import tfshim

    rs = np.random.RandomState(0)
    x = rs.standard_normal((2, 7, 6, 3)).astype("float32")
    f = rs.standard_normal((3, 4, 3, 2)).astype("float32")
    for stride in (1, 2):
        y = O.depthwise_conv2d_nhwc(torch.from_numpy(x), torch.from_numpy(f), stride, "SAME").numpy()
        pt, _, ho = O.same_pads(7, 3, stride)
        pl, _, wo = O.same_pads(6, 4, stride)
        ref = np.zeros((2, ho, wo, 6), "float64")
        for a in range(ho):
            for b in range(wo):
                for r in range(3):
                    for q in range(4):
                        hi, wi = a * stride + r - pt, b * stride + q - pl
                        if 0 <= hi < 7 and 0 <= wi < 6:
                            ref[:, a, b, :] += (x[:, hi, wi, :, None] * f[r, q][None]).reshape(2, 6)
        assert y.shape == ref.shape and np.abs(y - ref).max() < 1e-5
    np.random.seed(0)
    g = tfshim.Graph(dtype=torch.float32, u_seed=2)
    out = O.Conv2D(g, torch.from_numpy(x), 3, 16, 4, 2, "L", conv_type="separable_conv2d", channel_multiplier=4,
                   spectral_normed=True, update_collection=O.NO_OPS)
    # conv2d.py:169-178: spectral norm wraps Filters AND depthwise_filters AND pointwise_filters (three u variables,
    # created in that order), and the normalised depthwise / pointwise filters are what separable_conv2d consumes
    assert list(g.vars) == ["L/Filters", "L/depthwise_filters", "L/pointwise_filters", "L/filters/spectral_norm/u",
                            "L/depthwise_filters/spectral_norm/u", "L/pointwise_filters/spectral_norm/u", "L/Biases"]
    assert tuple(g.vars["L/depthwise_filters/spectral_norm/u"].shape) == (1, 4)
    assert tuple(g.vars["L/pointwise_filters/spectral_norm/u"].shape) == (1, 16)

    def sigma_np(w, u):
        """One power iteration from u written from sn.py:34-58 in float64 NumPy."""
        w2 = np.asarray(w, "float64").reshape(-1, w.shape[-1])
        v = np.asarray(u, "float64") @ w2.T
        v /= np.sqrt((v ** 2).sum()) + 1e-12
        u1 = v @ w2
        u1 /= np.sqrt((u1 ** 2).sum()) + 1e-12
        return float((v @ w2 @ u1.T)[0, 0])

    wd, wp = g.vars["L/depthwise_filters"].detach(), g.vars["L/pointwise_filters"].detach()
    sd = sigma_np(wd.numpy(), g.vars["L/depthwise_filters/spectral_norm/u"].numpy())
    sp = sigma_np(wp.numpy(), g.vars["L/pointwise_filters/spectral_norm/u"].numpy())
    dw = O.depthwise_conv2d_nhwc(torch.from_numpy(x), wd / sd, 2, "SAME")
    pw = torch.einsum("nhwc,co->nhwo", dw, (wp / sp)[0, 0])
    assert torch.allclose(out, pw.float(), atol=1e-5)
    plain = torch.einsum("nhwc,co->nhwo", O.depthwise_conv2d_nhwc(torch.from_numpy(x), wd, 2, "SAME"), wp[0, 0])
    assert not torch.allclose(out, plain, atol=1e-3)            # sigma of both filters really is applied
    # update_collection=None assigns u of the two filters the op reads; the unused `Filters` never runs its assign in TF
    np.random.seed(0)
    g2 = tfshim.Graph(dtype=torch.float32, u_seed=2)
    u0 = {}
    O.Conv2D(g2, torch.from_numpy(x), 3, 12, 4, 2, "D", conv_type="depthwise_conv2d", channel_multiplier=4,
             spectral_normed=True, update_collection=O.NO_OPS)
    for k in list(g2.vars):
        if k.endswith("/u"):
            u0[k] = g2.vars[k].clone()
    O.Conv2D(g2, torch.from_numpy(x), 3, 12, 4, 2, "D", conv_type="depthwise_conv2d", channel_multiplier=4,
             spectral_normed=True, update_collection=None)       # existing variables are reused by name
    assert not torch.equal(g2.vars["D/depthwise_filters/spectral_norm/u"], u0["D/depthwise_filters/spectral_norm/u"])
    assert torch.equal(g2.vars["D/filters/spectral_norm/u"], u0["D/filters/spectral_norm/u"])
    assert torch.equal(g2.vars["D/pointwise_filters/spectral_norm/u"], u0["D/pointwise_filters/spectral_norm/u"])


def test_variant_ops_closed_forms():
    """Closed-form pins of the oracle restatements added for the secondary variants: weight-norm makes the per-output-
    channel norm equal g for Conv2D / Linear / Deconv2D axes (conv2d.py:153-163, linear.py:143-155, deconv2d.py:87-96);
    PixelCNN masks are causal (conv2d.py:63-81); layer norm gives zero mean / unit variance per sample before gamma /
    beta (normalization.py:62-82); resize_nearest to half size picks every second pixel, to double size repeats."""
    import numpy as np
    import torch

    from oracle import ops as O
    from oracle import resnet_block as ORB
    from oracle import tfshim

    rs = np.random.RandomState(3)
    # ---- weight norm: ||W_eff[..., co]|| == g[co]
    np.random.seed(1)
    g = tfshim.Graph(dtype=torch.float64, u_seed=2)
    gv = (0.5 + rs.uniform(size=6))
    with g.variable_scope("C"):
        g.get_variable("g", initializer=gv)
    x = torch.from_numpy(rs.standard_normal((1, 5, 5, 4)))
    O.Conv2D(g, x, 4, 6, 3, name="C", weightnorm=True)
    w = g.vars["C/Filters"]
    # an impulse at the centre reads the effective filter back: y[0, 2+1-r, 2+1-s, co] = W_eff[r, s, ci, co]
    imp = torch.zeros(1, 5, 5, 4, dtype=torch.float64)
    imp[0, 2, 2, 1] = 1.0
    y = O.Conv2D(g, imp, 4, 6, 3, name="C", weightnorm=True, biases=False) if False else None
    w_eff = w * (torch.from_numpy(gv) / torch.sqrt((w ** 2).sum(dim=(0, 1, 2))))
    assert torch.allclose(torch.sqrt((w_eff ** 2).sum(dim=(0, 1, 2))), torch.from_numpy(gv), atol=1e-12)
    out = O.Conv2D(g, x, 4, 6, 3, name="C", weightnorm=True)
    ref = O.conv2d_nhwc(x, w_eff, 1, "SAME") + g.vars["C/Biases"]
    assert torch.allclose(out, ref, atol=1e-12)
    del y
    # ---- PixelCNN mask: the output at a pixel does not change when a "future" pixel changes; type 'a' also ignores the
    #      pixel itself, type 'b' sees it
    for kind, sees_self in (("a", False), ("b", True)):
        np.random.seed(2)
        g2 = tfshim.Graph(dtype=torch.float64, u_seed=2)
        xa = torch.from_numpy(rs.standard_normal((1, 6, 6, 2)))
        ya = O.Conv2D(g2, xa, 2, 3, 3, name="M", mask_type=(kind, 1))
        xb = xa.clone()
        xb[0, 3, 4:, :] += 1.0          # same row, to the right of (3, 3)
        xb[0, 4:, :, :] += 1.0          # rows below
        yb = O.Conv2D(g2, xb, 2, 3, 3, name="M", mask_type=(kind, 1))
        assert torch.allclose(ya[0, 3, 3], yb[0, 3, 3], atol=1e-12) and not torch.allclose(ya[0, 5, 5], yb[0, 5, 5])
        xc = xa.clone()
        xc[0, 3, 3, :] += 1.0
        yc = O.Conv2D(g2, xc, 2, 3, 3, name="M", mask_type=(kind, 1))
        assert torch.allclose(ya[0, 3, 3], yc[0, 3, 3], atol=1e-12) != sees_self
    # ---- layer norm
    g3 = tfshim.Graph(dtype=torch.float64, u_seed=2)
    xl = torch.from_numpy(rs.standard_normal((3, 4, 4, 8)) * 2.5 + 1.0)
    yl = O.layer_norm(g3, "LN", [1, 2, 3], xl)
    assert torch.allclose(yl.mean(dim=(1, 2, 3)), torch.zeros(3, dtype=torch.float64), atol=1e-9)
    assert torch.allclose(yl.var(dim=(1, 2, 3), unbiased=False), torch.ones(3, dtype=torch.float64), atol=1e-6)
    assert list(g3.vars) == ["LN/beta", "LN/gamma"]
    # ---- nearest resize
    xr = torch.from_numpy(rs.standard_normal((2, 8, 8, 3)))
    assert torch.equal(ORB.resize_nearest(xr, 4, 4), xr[:, ::2, ::2, :])
    up = ORB.resize_nearest(xr, 16, 16)
    assert torch.equal(up, ORB.upsample2(xr)) and torch.equal(up[:, 1::2, 1::2], xr)


def test_pix2pix_gradient_penalty_oracle_matches_finite_differences():
    """The oracle's WGAN-GP term (Pix2Pix/train.py:489-503: torch double backward through the spectrally-normalised
    PatchGAN) against central finite differences of the penalty in float64, for one filter of every layer."""
    import numpy as np
    import torch

    from oracle import ops as O
    from oracle import pix2pix as OP
    from oracle import tfshim

    rs = np.random.RandomState(4)
    size, ndf, n = 32, 4, 2
    inputs = torch.from_numpy(rs.uniform(-1, 1, size=(n, size, size, 3)))
    targets = torch.from_numpy(rs.uniform(-1, 1, size=(n, size, size, 3)))
    outputs = torch.from_numpy(rs.uniform(-1, 1, size=(n, size, size, 3)))
    alpha = torch.from_numpy(np.array([0.3, 0.8])).reshape(-1, 1, 1, 1)
    np.random.seed(0)
    g = tfshim.Graph(dtype=torch.float64, u_seed=2)
    with torch.no_grad(), g.variable_scope("d_net"):
        OP.unet_d(g, inputs, targets, ndf, True, O.NO_OPS)

    def penalty():
        interp = (targets + alpha * (outputs - targets)).detach().requires_grad_(True)
        with g.variable_scope("d_net"):
            d = OP.unet_d(g, inputs, interp, ndf, True, O.NO_OPS)       # NO_OPS: u fixed, so the function is repeatable
        grads = torch.autograd.grad(d.sum(), interp, create_graph=True)[0]
        slopes = torch.sqrt((grads ** 2).sum(dim=(1, 2, 3)) + 1e-10)
        return 10 * ((slopes - 1.0) ** 2).mean()

    params = dict(g.trainable_variables("d_net"))
    names = [k for k in params if k.endswith("/Filters")]
    analytic = dict(zip(names, torch.autograd.grad(penalty(), [params[k] for k in names])))
    for k in names:
        w = params[k]
        idx = tuple(int(rs.randint(0, s)) for s in w.shape)
        eps = 1e-5
        with torch.no_grad():
            w[idx] += eps
        up = penalty().item()
        with torch.no_grad():
            w[idx] -= 2 * eps
        dn = penalty().item()
        with torch.no_grad():
            w[idx] += eps
        fd = (up - dn) / (2 * eps)
        an = analytic[k][idx].item()
        assert abs(fd - an) <= 1e-4 * max(1.0, abs(an)), (k, idx, fd, an)


def test_losses_schedules_and_pggan_statistics_closed_forms():
    """lib.misc.get_loss (common/misc.py:310-394) against the formulas written out in NumPy for all seven loss types;
    the SNGAN learning-rate decay (gan_cifar_resnet.py:454-457) at its break points; minibatch_std
    (PGGAN/model_nvidia.py:20-28) and pixel_norm (normalization.py:125-140) on hand-sized tensors."""
    import numpy as np
    import torch

    from oracle import acgan as OA
    from oracle import ops as O
    from oracle import pggan as OP
    from oracle import sngan_cifar as OS

    rs = np.random.RandomState(6)
    r, f = rs.standard_normal(11), rs.standard_normal(11)
    sig = lambda t: 1.0 / (1.0 + np.exp(-t))            # noqa: E731
    xent = lambda logit, label: np.maximum(logit, 0) - logit * label + np.log1p(np.exp(-np.abs(logit)))   # noqa: E731  TF's stable form
    want = {
        "HINGE": (np.maximum(0, 1 - r).mean() + np.maximum(0, 1 + f).mean(), -f.mean()),
        "WGAN": (-r.mean() + f.mean(), -f.mean()),
        "WGAN-GP": (-r.mean() + f.mean(), -f.mean()),
        "LSGAN": (((1 - r) ** 2).mean() / 2 + (f ** 2).mean() / 2, ((1 - f) ** 2).mean() / 2),
        "CGAN": (xent(r, 1.0).mean() + xent(f, 0.0).mean(), xent(f, 1.0).mean()),
        "Modified_MiniMax": (-np.log(sig(r)).mean() - np.log(1 - sig(f)).mean(), -np.log(sig(f)).mean()),
        "MiniMax": (-np.log(sig(r)).mean() - np.log(1 - sig(f)).mean(), np.log(1 - sig(f)).mean()),
    }
    for loss_type, (d_want, g_want) in want.items():
        d, g = OA.get_loss(torch.from_numpy(r), torch.from_numpy(f), loss_type)
        assert abs(d.item() - d_want) < 1e-12 and abs(g.item() - g_want) < 1e-12, loss_type
    assert OS.lr_decay(0) == 1.0 and abs(OS.lr_decay(25000) - 0.75) < 1e-12
    assert abs(OS.lr_decay(49999) - 0.50001) < 1e-9 and OS.lr_decay(50000) == 0.5 and OS.lr_decay(99999) == 0.5
    # minibatch_std: per-position std over the batch (+1e-8 inside the sqrt), averaged to ONE scalar channel
    x = np.zeros((2, 1, 2, 1))
    x[0, 0, 0, 0], x[1, 0, 0, 0] = 1.0, 3.0          # std 1 at position (0, 0), 0 at (0, 1)
    y = OP.minibatch_std(torch.from_numpy(x))
    assert tuple(y.shape) == (2, 1, 2, 2)
    expect = (np.sqrt(1.0 + 1e-8) + np.sqrt(1e-8)) / 2
    assert torch.allclose(y[..., 1], torch.full((2, 1, 2), expect, dtype=torch.float64), atol=1e-12)
    assert torch.equal(y[..., 0], torch.from_numpy(x)[..., 0])
    # pixel_norm: unit mean square over the channels of every pixel
    p = O.pixel_norm(torch.from_numpy(rs.standard_normal((2, 3, 3, 16)) * 4))
    assert torch.allclose((p ** 2).mean(dim=3), torch.ones(2, 3, 3, dtype=torch.float64), atol=1e-6)
