"""GPU parity tests for the widened rows of SURVEY 8(a): strided / asymmetric-SAME / explicitly padded convolutions
(Pix2Pix encoders and PatchGAN, a-3 / a-17), Deconv2D (a-4), pixel norm (+ leaky ReLU) and minibatch standard deviation
(PGGAN, a-8 / a-16), U-Net channel concatenation and dropout masks (a-17).  Same protocol as test_gpu_ops.py: the
CUDA path through the C ABI against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest
import torch

from tests.test_gpu_ops import TOL_F32, _bf16_repr, check, env, rel, run_pair  # noqa: F401

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------------ strided Conv2D
@pytest.mark.parametrize("n,h,w,cin,cout,k,stride,sn", [
    (4, 32, 32, 64, 128, 4, 2, False),    # Pix2Pix encoder: 4x4 s2 SAME (pads 1, 1)
    (3, 16, 16, 128, 64, 3, 2, True),     # 3x3 s2 SAME on an even size: pads (0, 1) -- TF's asymmetric rule
    (2, 16, 16, 64, 64, 4, 1, False),     # 4x4 s1 SAME: pads (1, 2) (Pix2Pix decoders)
    (5, 9, 9, 72, 40, 3, 2, False),       # odd size, ragged channels
    (4, 32, 32, 3, 64, 4, 2, False),      # RGB encoder_1: 48 im2col columns (64-wide route), stride 2
    (2, 32, 32, 6, 64, 4, 2, True),       # PatchGAN input (image + target = 6 channels): 96 columns
    (2, 16, 16, 128, 1, 4, 1, True),      # PatchGAN head: Cout = 1
    (2, 32, 32, 16, 3, 4, 1, False),      # U-Net decoder_1: 4x4 s1 SAME to RGB (48 im2col columns on the gradient side)
    (1, 64, 64, 8, 8, 4, 2, False),       # 8-channel layers (ngf = 8)
    (1, 64, 64, 16, 8, 4, 1, False),
])
def test_strided_conv2d_matches_oracle(env, n, h, w, cin, cout, k, stride, sn):
    store, tfshim = env
    from gan_lib_tensorflow_b200.common.ops import conv2d as P
    from oracle import ops as O

    x = np.random.RandomState(41).standard_normal((n, h, w, cin)).astype("float32")
    prod, refs = run_pair(
        store, tfshim,
        lambda xv: P.Conv2D(xv, cin, cout, k, stride, "L", spectral_normed=sn, update_collection="NO_OPS"),
        lambda g, xt: O.Conv2D(g, xt, cin, cout, k, stride, "L", spectral_normed=sn, update_collection=O.NO_OPS),
        x)
    check(prod, refs)


@pytest.mark.parametrize("stride", [1, 2])
def test_explicitly_padded_valid_conv(env, stride):
    """tf.pad(x, 1) followed by a 4x4 VALID convolution (Pix2Pix/networks.py:287-354)."""
    store, tfshim = env
    from gan_lib_tensorflow_b200.common.ops import conv2d as P
    from oracle import ops as O

    x = np.random.RandomState(42).standard_normal((3, 16, 16, 64)).astype("float32")

    def orc(g, xt):
        xp = torch.nn.functional.pad(xt, (0, 0, 1, 1, 1, 1))
        return O.Conv2D(g, xp, 64, 128, 4, stride, "L", padding="VALID")

    prod, refs = run_pair(store, tfshim,
                          lambda xv: P.Conv2D(xv, 64, 128, 4, stride, "L", padding=(1, 1, 1, 1)), orc, x)
    check(prod, refs)


# ------------------------------------------------------------------------------------------------ Deconv2D
@pytest.mark.parametrize("n,h,cin,cout,k", [(4, 8, 128, 64, 4), (2, 16, 64, 64, 3), (3, 5, 72, 40, 4)])
def test_deconv2d_matches_oracle(env, n, h, cin, cout, k):
    """tf.nn.conv2d_transpose to [n, 2h, 2w, cout] (deconv2d.py:99-109): forward, dx, dFilters, dBiases.  Inputs,
    filters and the cotangent are bf16-representable so the fp32 oracle is the exact reference of the bf16 kernels."""
    store, tfshim = env
    from gan_lib_tensorflow_b200.common.ops import deconv2d as P
    from oracle import ops as O

    x = _bf16_repr(np.random.RandomState(43).standard_normal((n, h, h, cin)).astype("float32"))
    cot = _bf16_repr(np.random.RandomState(44).standard_normal((n, 2 * h, 2 * h, cout)).astype("float32"))
    wv = _bf16_repr((np.random.RandomState(45).standard_normal((k, k, cout, cin)) * 0.05).astype("float32"))

    def prod_fn(xv):
        with store.variable_scope("L"):
            store.get_variable("Filters", initializer=wv)
        return P.Deconv2D(xv, cin, cout, k, name="L")

    def orc_fn(g, xt):
        with g.variable_scope("L"):
            g.get_variable("Filters", initializer=wv)
        return O.Deconv2D(g, xt, cin, cout, k, name="L")

    prod, refs = run_pair(store, tfshim, prod_fn, orc_fn, x, cot_np=cot, bf16=False)
    check(prod, refs, tol_fp32=2e-3)   # only the bf16 rounding of dx remains (everything else is exact products)


# ------------------------------------------------------------------------------------------------ PGGAN pieces
@pytest.mark.parametrize("shape,act", [((4, 8, 8, 512), "lrelu"), ((3, 5, 7, 64), None), ((2, 4, 4, 12), "lrelu")])
def test_pixel_norm_lrelu(env, shape, act):
    store, _ = env
    from gan_lib_tensorflow_b200 import functional as F
    from gan_lib_tensorflow_b200.common.ops import normalization as P
    from oracle import ops as O

    rs = np.random.RandomState(46)
    x = (rs.standard_normal(shape) * 1.5).astype("float32")
    cot = rs.standard_normal(shape).astype("float32")
    xv = F.Var(torch.from_numpy(x).cuda(), requires_grad=True)
    with store.gradient_tape() as tape:
        out = P.pixel_norm(xv, act=act)
        tape.backward(out, grad=torch.from_numpy(cot).cuda())
    xt = torch.from_numpy(x).double().requires_grad_(True)
    yo = O.pixel_norm(xt)
    if act == "lrelu":
        yo = torch.maximum(yo, 0.2 * yo)       # PGGAN/model_nvidia.py:15-17
    (dx,) = torch.autograd.grad(yo, xt, torch.from_numpy(cot).double())
    assert rel(out.data.cpu().numpy(), yo.detach().numpy()) < TOL_F32
    assert rel(xv.grad.cpu().numpy(), dx.numpy()) < TOL_F32


@pytest.mark.parametrize("b,h,c", [(16, 4, 512), (6, 8, 32), (4, 2, 3)])
def test_minibatch_std(env, b, h, c):
    """PGGAN/model_nvidia.py:20-28."""
    store, _ = env
    from gan_lib_tensorflow_b200 import functional as F

    rs = np.random.RandomState(47)
    x = rs.standard_normal((b, h, h, c)).astype("float32")
    cot = rs.standard_normal((b, h, h, c + 1)).astype("float32")
    xv = F.Var(torch.from_numpy(x).cuda(), requires_grad=True)
    with store.gradient_tape() as tape:
        out = F.minibatch_std(xv)
        tape.backward(out, grad=torch.from_numpy(cot).cuda())
    xt = torch.from_numpy(x).double().requires_grad_(True)
    m = xt.mean(dim=0, keepdim=True)
    v = ((xt - m) * (xt - m)).mean(dim=0, keepdim=True)
    std = torch.sqrt(v + 1e-8).mean().reshape(1, 1, 1, 1).expand(b, h, h, 1)
    yo = torch.cat([xt, std], dim=3)
    (dx,) = torch.autograd.grad(yo, xt, torch.from_numpy(cot).double())
    assert out.shape == (b, h, h, c + 1)
    assert rel(out.data.cpu().numpy(), yo.detach().numpy()) < TOL_F32
    assert rel(xv.grad.cpu().numpy(), dx.numpy()) < 5e-5


# ------------------------------------------------------------------------------------------------ Pix2Pix pieces
def test_concat_channels_and_dropout(env):
    store, _ = env
    from gan_lib_tensorflow_b200 import functional as F

    rs = np.random.RandomState(48)
    a = rs.standard_normal((3, 8, 8, 64)).astype("float32")
    b = _bf16_repr(rs.standard_normal((3, 8, 8, 32)).astype("float32"))
    mask = (rs.uniform(size=(3, 8, 8, 96)) < 0.5).astype("float32")
    cot = rs.standard_normal((3, 8, 8, 96)).astype("float32")
    av = F.Var(torch.from_numpy(a).cuda(), requires_grad=True)
    bv = F.Var(torch.from_numpy(b).cuda().to(torch.bfloat16), requires_grad=True)
    with store.gradient_tape() as tape:
        cat = F.concat_channels(av, bv, out_dtype=torch.float32)
        out = F.dropout(cat, torch.from_numpy(mask).cuda(), 0.5)
        tape.backward(out, grad=torch.from_numpy(cot).cuda())
    ref = np.concatenate([a, b], axis=3) * mask / 0.5
    dref = cot * mask / 0.5
    assert rel(out.data.cpu().numpy(), ref) < 1e-7
    assert rel(av.grad.cpu().numpy(), dref[..., :64]) < 1e-7
    assert rel(bv.grad.float().cpu().numpy(), dref[..., 64:]) < 4e-3     # bf16 gradient storage


# ------------------------------------------------------------------------------------------------ Pix2Pix networks
def check_band(prod, refs, factor=2.0, floor=2e-3, tag=""):
    """Deep composites (16 convolutions, instance norms in between): two bf16 implementations whose fp32 accumulators
    differ by 1e-7 flip a few bf16 roundings per layer, the flips compound (x3-10 per layer, tests/probe_unet_layers.py)
    and both end at the bf16 noise floor.  The band is therefore measured, not assumed: the product must be as close
    to the fp32 oracle as the bf16-operand ORACLE is (x factor), tensor by tensor."""
    from tests.test_gpu_ops import _report
    f32, b16 = refs["fp32"], refs["bf16"]
    pairs = [("out", prod["out"], b16["out"], f32["out"])]
    if f32["dx"] is not None:
        pairs.append(("dx", prod["dx"], b16["dx"], f32["dx"]))
    gmax = max([np.linalg.norm(g) for g in f32["params"].values()] + [1e-30])
    worst = []
    for name, gr in f32["params"].items():
        if np.linalg.norm(gr) < 5e-2 * gmax:
            continue   # analytically-zero gradients (biases in front of a norm) hold rounding residue only
        pairs.append((name, prod["params"][name], b16["params"][name], gr))
    for name, p, b, f in pairs:
        e_prod, e_orc = rel(p, f), rel(b, f)
        worst.append((e_prod / (factor * e_orc + floor), name, e_prod, e_orc))
    worst.sort(reverse=True)
    _report(f"band {tag}: " + " ".join(f"{n}={ep:.2e}/{eo:.2e}" for _, n, ep, eo in worst[:8]))
    ratio, name, e_prod, e_orc = worst[0]
    assert ratio <= 1.0, (name, e_prod, e_orc)
def test_pix2pix_unet_generator(env):
    """Pix2Pix/networks.py:174-284, ngf = 8, with dropout masks, at 512x512 so that the bottleneck is 2x2: at the
    256x256 of config 4 the instance norm of encoder_8 sees ONE pixel, its output is `beta` plus the rounding residue
    of x*inv + (beta - mean*inv) (tf.nn.batch_normalization's formula), and the relu mask behind it is decided by that
    residue -- no two implementations (TF included) agree there, so the parity test keeps every norm well-posed."""
    store, tfshim = env
    from gan_lib_tensorflow_b200.Pix2Pix import networks as P
    from oracle import pix2pix as OP

    n, ngf, size = 1, 8, 512
    rs = np.random.RandomState(51)
    x = rs.uniform(-1, 1, size=(n, size, size, 3)).astype("float32")
    masks = [(rs.uniform(size=(n, s, s, ngf * 8)) < 0.5).astype("float32") for s in (4, 8, 16)]
    cot = rs.standard_normal((n, size, size, 3)).astype("float32")
    prod, refs = run_pair(
        store, tfshim,
        lambda xv: P.unet_g(xv, 3, ngf, keep_masks=[torch.from_numpy(m).cuda() for m in masks]),
        lambda g, xt: OP.unet_g(g, xt, 3, ngf, keep_masks=[torch.from_numpy(m) for m in masks]),
        x, cot_np=cot)
    assert rel(prod["out"], refs["bf16"]["out"]) < 1e-2
    check_band(prod, refs, tag="unet_g")


def test_pix2pix_patchgan_discriminator(env):
    """Pix2Pix/networks.py:287-354 at 64x64: [n,64,64,3]x2 -> [n,6,6,1], spectral-normed."""
    store, tfshim = env
    from gan_lib_tensorflow_b200 import functional as F
    from gan_lib_tensorflow_b200.Pix2Pix import networks as P
    from oracle import ops as O
    from oracle import pix2pix as OP

    n, ndf = 3, 16
    rs = np.random.RandomState(52)
    x = rs.uniform(-1, 1, size=(n, 64, 64, 3)).astype("float32")
    tgt = rs.uniform(-1, 1, size=(n, 64, 64, 3)).astype("float32")
    prod, refs = run_pair(
        store, tfshim,
        lambda xv: P.unet_d(xv, F.Var(torch.from_numpy(tgt).cuda()), ndf, True, "NO_OPS"),
        lambda g, xt: OP.unet_d(g, xt, torch.from_numpy(tgt), ndf, True, O.NO_OPS), x)
    assert prod["out"].shape == (n, 6, 6, 1)
    check(prod, refs, tol_impl=6e-3, tol_fp32=1e-1, tag="unet_d")   # 5 layers of lrelu-mask sensitivity vs fp32


def test_pix2pix_512_unet_generator_and_discriminator(env):
    """Pix2Pix/networks.py:359-536: unet_generator (nine encoder / decoder levels) and unet_discriminator (n_layers = 4).
    The generator runs at 1024x1024 with ngf = 8 so that encoder_9's instance norm still sees 2x2 pixels (see
    test_pix2pix_unet_generator); forward only for G (its backward is the same ops as unet_g's), forward + backward for D."""
    store, tfshim = env
    from gan_lib_tensorflow_b200 import functional as F
    from gan_lib_tensorflow_b200.Pix2Pix import networks as P
    from oracle import ops as O
    from oracle import ops as O_ops
    from oracle import pix2pix as OP

    n, ngf, size = 1, 8, 1024
    rs = np.random.RandomState(53)
    x = rs.uniform(-1, 1, size=(n, size, size, 3)).astype("float32")
    masks = [(rs.uniform(size=(n, s, s, ngf * 8)) < 0.5).astype("float32") for s in (4, 8, 16)]
    np.random.seed(0)
    with store.variable_scope("G"):
        out = P.unet_generator(torch.from_numpy(x).cuda(), 3, ngf,
                               keep_masks=[torch.from_numpy(m).cuda() for m in masks])
    torch.cuda.synchronize()
    assert tuple(out.shape) == (n, size, size, 3)
    assert "G/encoder_9/Conv2D/Filters" in store.vars and "G/decoder_9/InstanceNorm/gamma" in store.vars
    errs = {}
    for mode in (True, False):
        O_ops.BF16_OPERANDS = mode
        try:
            np.random.seed(0)
            g = tfshim.Graph(dtype=torch.float32, u_seed=2)
            with torch.no_grad(), g.variable_scope("G"):
                ref = OP.unet_generator(g, torch.from_numpy(x), 3, ngf, keep_masks=[torch.from_numpy(m) for m in masks])
            errs[mode] = rel(out.data.float().cpu().numpy(), ref.numpy())
            if mode:
                ref16 = ref.numpy()
            else:
                e_orc = rel(ref16, ref.numpy())
        finally:
            O_ops.BF16_OPERANDS = False
    print(f"unet_generator 1024: vs bf16-oracle {errs[True]:.2e}, vs fp32 {errs[False]:.2e} (bf16-oracle vs fp32 {e_orc:.2e})")
    assert sorted(k for k in store.vars) == sorted(k for k in g.vars if "/moving_" not in k)
    assert errs[True] < 1e-2 and errs[False] <= 1.5 * e_orc + 1e-2

    ndf = 16
    xd = rs.uniform(-1, 1, size=(2, 128, 128, 3)).astype("float32")
    tgt = rs.uniform(-1, 1, size=(2, 128, 128, 3)).astype("float32")
    prod, refs = run_pair(
        store, tfshim,
        lambda xv: P.unet_discriminator(xv, F.Var(torch.from_numpy(tgt).cuda()), ndf, True, "NO_OPS"),
        lambda g_, xt: OP.unet_discriminator(g_, xt, torch.from_numpy(tgt), ndf, True, O.NO_OPS), xd)
    assert prod["out"].shape == (2, 6, 6, 1)
    # six layers of leaky-ReLU mask flips (unet_d's five sit at 6e-3 against the bf16-operand oracle): measured band
    assert rel(prod["out"], refs["bf16"]["out"]) < 4e-3
    check_band(prod, refs, tag="unet_discriminator")


# ------------------------------------------------------------------------------------------------ SNGAN ImageNet-128
def test_imagenet_generator_forward_full_width(env):
    """gan_imagNet_resnet.py:241-271 at the real widths (DIM_G = 128: 1024 -> 64 channels, 4x4 -> 128x128), 1000-class
    conditional batch norm; forward against both oracles."""
    store, tfshim = env
    from gan_lib_tensorflow_b200 import functional as F
    from gan_lib_tensorflow_b200.SNGAN import gan_imagNet_resnet as P
    from oracle import ops as O_ops
    from oracle import sngan_imagenet as OI

    n = 4
    rs = np.random.RandomState(61)
    z = rs.standard_normal((n, 128)).astype("float32")
    labels = rs.randint(0, 1000, size=n).astype("int32")
    np.random.seed(0)
    out = P.Generator(n, torch.from_numpy(labels).cuda(), noise=torch.from_numpy(z).cuda())
    got = out.data.float().cpu().numpy()
    assert got.shape == (n, 49152)
    errs = {}
    for mode in (True, False):
        O_ops.BF16_OPERANDS = mode
        np.random.seed(0)
        g = tfshim.Graph(dtype=torch.float32, u_seed=2)
        with torch.no_grad():
            ref = OI.Generator(g, n, torch.from_numpy(labels).long(), torch.from_numpy(z)).numpy()
        errs[mode] = rel(got, ref)
    O_ops.BF16_OPERANDS = False
    assert errs[True] < 2e-2 and errs[False] < 3e-2, errs      # 16 bf16 layers deep, batch statistics over 4 samples
    assert set(v.key for v in store.trainable_variables("Generator")) == set(n_ for n_, _ in g.trainable_variables())


def test_imagenet_discriminator_full_width(env):
    """gan_imagNet_resnet.py:274-334 at DIM_D = 128 (64 -> 1024 channels, label map concatenated at 16x16)."""
    store, tfshim = env
    from gan_lib_tensorflow_b200.SNGAN import gan_imagNet_resnet as P
    from oracle import ops as O
    from oracle import sngan_imagenet as OI

    n = 2
    rs = np.random.RandomState(62)
    x = rs.uniform(-1, 1, size=(n, 49152)).astype("float32")
    labels = rs.randint(0, 1000, size=n).astype("int32")
    prod, refs = run_pair(
        store, tfshim,
        lambda xv: P.Discriminator(xv, torch.from_numpy(labels).cuda(), update_collection="NO_OPS")[0],
        lambda g, xt: OI.Discriminator(g, xt, torch.from_numpy(labels).long(), update_collection=O.NO_OPS)[0],
        x, cot_np=rs.standard_normal((n,)).astype("float32"))
    assert rel(prod["out"], refs["bf16"]["out"]) < 5e-3
    check_band(prod, refs, tag="imagenet_d")


# ------------------------------------------------------------------------------------------------ Pix2Pix training
@pytest.mark.parametrize("loss_type", ["HINGE", "LSGAN"])
def test_pix2pix_training_steps(env, loss_type):
    """Pix2Pix/train.py:447-539: critic gradients of get_loss(loss_type) with D(real) and D(fake) both assigning u
    (update_collection=None: two different sigma in one gradient computation) and generator gradients of
    gan_weight * GAN + l1_weight * L1 through the U-Net with dropout, vs the oracle; then one Adam iteration."""
    store, tfshim = env
    from gan_lib_tensorflow_b200.Pix2Pix import train as PT
    from oracle import ops as O_ops
    from oracle import pix2pix as OP

    n, ngf, size = 1, 8, 512
    rs = np.random.RandomState(95)
    x = rs.uniform(-1, 1, size=(n, size, size, 3)).astype("float32")
    tgt = rs.uniform(-1, 1, size=(n, size, size, 3)).astype("float32")
    masks = [(rs.uniform(size=(n, s, s, ngf * 8)) < 0.5).astype("float32") for s in (4, 8, 16)]
    tr = PT.Trainer(ngf=ngf, ndf=ngf, size=size, loss_type=loss_type, seed=0, max_steps=100)
    xd, td = torch.from_numpy(x).cuda(), torch.from_numpy(tgt).cuda()
    md = [torch.from_numpy(m).cuda() for m in masks]
    dl = tr.players.gradients("d", lambda: tr.d_loss(xd, td, md))
    d_grads = {v.key: v.grad.cpu().numpy().copy() for v in store.trainable_variables("d_net")}
    u_after_d = {k: v.data.cpu().numpy().copy() for k, v in store.vars.items() if k.endswith("/u")}
    gl = tr.players.gradients("g", lambda: tr.g_loss(xd, td, md))
    g_grads = {v.key: v.grad.cpu().numpy().copy() for v in store.trainable_variables("g_net")}
    d_loss, g_loss = float(dl.data.item()), float(gl.data.item())
    refs = {}
    for mode in (True, False):
        O_ops.BF16_OPERANDS = mode
        try:
            np.random.seed(0)
            g = tfshim.Graph(dtype=torch.float32, u_seed=2)
            ol = OP.Pix2PixLosses(g, ngf, ngf, size, loss_type)
            mt = [torch.from_numpy(m) for m in masks]
            dc, dp, dg = ol.d_grads(torch.from_numpy(x), torch.from_numpy(tgt), mt)
            u_ref = {k_: v.detach().numpy().copy() for k_, v in g.vars.items() if k_.endswith("/u")}
            gc, gp, gg = ol.g_grads(torch.from_numpy(x), torch.from_numpy(tgt), mt)
            refs[mode] = dict(d=dc.item(), g=gc.item(), u=u_ref,
                              dg={nm: t.numpy() for (nm, _), t in zip(dp, dg) if t is not None},
                              gg={nm: t.numpy() for (nm, _), t in zip(gp, gg) if t is not None})
        finally:
            O_ops.BF16_OPERANDS = False
    assert set(d_grads) == set(refs[False]["dg"]) and set(g_grads) == set(refs[False]["gg"])
    assert abs(d_loss - refs[True]["d"]) < 5e-3 * max(1.0, abs(refs[True]["d"]))
    assert abs(g_loss - refs[True]["g"]) < 5e-3 * max(1.0, abs(refs[True]["g"]))
    for name, u in refs[False]["u"].items():          # two assignments during the critic pass (real, then fake)
        assert rel(u_after_d[name], u) < 1e-3, name
    for got, key in ((d_grads, "dg"), (g_grads, "gg")):
        gmax = max(np.linalg.norm(t) for t in refs[False][key].values())
        for name, f32 in refs[False][key].items():
            if np.linalg.norm(f32) < 5e-2 * gmax:
                continue
            e_prod, e_orc = rel(got[name], f32), rel(refs[True][key][name], f32)
            assert e_prod <= 2.0 * e_orc + 5e-3, (key, name, e_prod, e_orc)
    lr0 = tr.learning_rate()
    d, g_ = tr.train_iteration(xd, td, n_dis=2, mask_fn=lambda: md)
    torch.cuda.synchronize()
    assert np.isfinite(d.data.item()) and np.isfinite(g_.data.item())
    assert tr.global_step == 1 and tr.learning_rate() < lr0 == 0.0002


def test_pix2pix_gradient_penalty(env):
    """Pix2Pix --loss_type WGAN-GP (train.py:489-503): the penalty of the spectrally-normalised PatchGAN on the
    interpolates, a THIRD update_collection=None evaluation of D inside the critic step, differentiated through D's
    backward pass (Pix2Pix/gp.py) -- value, total critic loss, u after three assignments and every critic gradient
    against the oracle's torch double backward."""
    store, tfshim = env
    from gan_lib_tensorflow_b200.Pix2Pix import train as PT
    from oracle import ops as O_ops
    from oracle import pix2pix as OP
    from tests.test_gpu_ops import _report

    n, ngf, size = 2, 8, 512
    rs = np.random.RandomState(96)
    x = rs.uniform(-1, 1, size=(n, size, size, 3)).astype("float32")
    tgt = rs.uniform(-1, 1, size=(n, size, size, 3)).astype("float32")
    alpha = np.array([0.37, 0.81], dtype="float32")
    tr = PT.Trainer(ngf=ngf, ndf=ngf, size=size, loss_type="WGAN-GP", seed=0, max_steps=100)
    xd, td = torch.from_numpy(x).cuda(), torch.from_numpy(tgt).cuda()
    dl = tr.players.gradients("d", lambda: tr.d_loss(xd, td, None, gp_alpha=torch.from_numpy(alpha).cuda()))
    torch.cuda.synchronize()
    d_grads = {v.key: v.grad.cpu().numpy().copy() for v in store.trainable_variables("d_net")}
    u_after = {k: v.data.cpu().numpy().copy() for k, v in store.vars.items() if k.endswith("/u")}
    d_loss, pen = float(dl.data.item()), float(tr.last_penalty.item())
    refs = {}
    for mode in (True, False):
        O_ops.BF16_OPERANDS = mode
        try:
            np.random.seed(0)
            g = tfshim.Graph(dtype=torch.float32, u_seed=2)
            ol = OP.Pix2PixLosses(g, ngf, ngf, size, "WGAN-GP")
            dc, dp, dg = ol.d_grads(torch.from_numpy(x), torch.from_numpy(tgt), None, gp_alpha=torch.from_numpy(alpha))
            refs[mode] = dict(d=dc.item(), pen=float(ol.last_penalty.item()),
                              u={k_: v.detach().numpy().copy() for k_, v in g.vars.items() if k_.endswith("/u")},
                              dg={nm: t.numpy() for (nm, _), t in zip(dp, dg) if t is not None})
        finally:
            O_ops.BF16_OPERANDS = False
    _report(f"pix2pix WGAN-GP: penalty product {pen:.5f} bf16-oracle {refs[True]['pen']:.5f} fp32-oracle {refs[False]['pen']:.5f}; "
            f"d_loss {d_loss:.5f} / {refs[True]['d']:.5f} / {refs[False]['d']:.5f}")
    assert set(d_grads) == set(refs[False]["dg"])
    assert refs[False]["pen"] > 1e-3                                            # the term is active in this test
    assert abs(pen - refs[True]["pen"]) <= 1e-2 * refs[True]["pen"] + 1e-4
    assert abs(pen - refs[False]["pen"]) <= 3e-2 * refs[False]["pen"] + 1e-4
    assert abs(d_loss - refs[True]["d"]) < 5e-3 * max(1.0, abs(refs[True]["d"]))
    for name, u in refs[False]["u"].items():          # three assignments: D(real), D(fake), D(interpolates)
        assert rel(u_after[name], u) < 1e-3, name
    gmax = max(np.linalg.norm(t) for t in refs[False]["dg"].values())
    lines = []
    for name, f32 in refs[False]["dg"].items():
        if np.linalg.norm(f32) < 5e-2 * gmax:
            continue
        e_impl = rel(d_grads[name], refs[True]["dg"][name])
        e_prod, e_orc = rel(d_grads[name], f32), rel(refs[True]["dg"][name], f32)
        lines.append(f"{name}={e_impl:.1e}/{e_prod:.1e}/{e_orc:.1e}")
        assert e_prod <= 2.0 * e_orc + 1e-2, (name, e_impl, e_prod, e_orc)
    _report("pix2pix WGAN-GP gradients (vs bf16-oracle / vs fp32 / bf16-oracle vs fp32): " + " ".join(lines))
    # the step runs through the optimiser, eagerly and as a captured graph
    d1 = tr.d_step(xd, td)
    tr.g_step(xd, td)
    tr.capture(xd, td)
    d2 = tr.d_step(xd, td)
    torch.cuda.synchronize()
    assert np.isfinite(float(d1.data.reshape(-1)[0])) and np.isfinite(float(d2.reshape(-1)[0]))
