"""GPU parity tests for the secondary layer variants: weight-norm of Conv2D / conv2d_.Conv2D /
Linear / Deconv2D (common/ops/conv2d.py:153-163, linear.py:143-155, deconv2d.py:87-96), PixelCNN masks
(conv2d.py:63-81, 165-167), layer norm of the critic (normalization.py:62-82, NORMALIZATION_D), the PGGAN fade-in
with alpha in device memory and the CUDA-graph capture of the PGGAN training ops that it enables.  Same protocol as
test_gpu_ops.py: CUDA path through the C ABI against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest
import torch

from tests.test_gpu_graphs import Snapshot, _compare, _state
from tests.test_gpu_ops import TOL_F32, _bf16_repr, check, env, rel, run_pair  # noqa: F401

pytestmark = pytest.mark.gpu


def _g_values(cout, seed):
    """Target norms away from their initial value (= the norms of the initial filters), so that g / ||W|| != 1."""
    return (0.5 + np.random.RandomState(seed).uniform(size=cout)).astype("float32")


# ------------------------------------------------------------------------------------------------ weight-norm
@pytest.mark.parametrize("n,h,cin,cout,k,sn,module", [
    (4, 16, 64, 64, 3, False, "conv2d"),
    (3, 8, 72, 40, 3, True, "conv2d"),       # ragged channels; spectral norm of the weight-normed filters
    (5, 16, 3, 128, 3, False, "conv2d"),     # RGB side (im2col route takes its operand from the effective filters)
    (2, 16, 128, 256, 1, False, "conv2d_"),
])
def test_weightnorm_conv2d(env, n, h, cin, cout, k, sn, module):
    store, tfshim = env
    import importlib

    P = importlib.import_module("gan_lib_tensorflow_b200.common.ops." + module)
    from oracle import ops as O

    x = np.random.RandomState(1).standard_normal((n, h, h, cin)).astype("float32")
    gv = _g_values(cout, 2)

    def prod_fn(xv):
        with store.variable_scope("L"):
            store.get_variable("g", initializer=gv)
        return P.Conv2D(xv, cin, cout, k, name="L", weightnorm=True, spectral_normed=sn, update_collection="NO_OPS")

    def orc_fn(g, xt):
        with g.variable_scope("L"):
            g.get_variable("g", initializer=gv)
        return O.Conv2D(g, xt, cin, cout, k, name="L", weightnorm=True, spectral_normed=sn,
                        update_collection="NO_OPS")

    prod, refs = run_pair(store, tfshim, prod_fn, orc_fn, x)
    assert "L/g" in refs["fp32"]["params"] and "L/Filters" in refs["fp32"]["params"]
    check(prod, refs, tag=f"weightnorm {module} {cin}->{cout} k{k} sn={sn}")


def test_weightnorm_initial_g_is_the_initial_norm(env):
    """conv2d.py:154-158: g starts at the norms of the initial filter values, so the first forward pass is the plain
    convolution."""
    store, _ = env
    from gan_lib_tensorflow_b200 import functional as F
    from gan_lib_tensorflow_b200.common.ops import conv2d as P

    x = torch.from_numpy(np.random.RandomState(3).standard_normal((2, 8, 8, 64)).astype("float32")).cuda()
    np.random.seed(0)
    a = P.Conv2D(F.Var(x), 64, 64, 3, name="A", weightnorm=True)
    np.random.seed(0)
    b = P.Conv2D(F.Var(x), 64, 64, 3, name="B", weightnorm=False)
    torch.cuda.synchronize()
    w = store.vars["A/Filters"].data
    assert rel(store.vars["A/g"].data.cpu().numpy(), w.pow(2).sum(dim=(0, 1, 2)).sqrt().cpu().numpy()) < 1e-6
    assert rel(a.data.float().cpu().numpy(), b.data.float().cpu().numpy()) < 5e-3   # W*(g/|W|) rounds differently


def test_weightnorm_linear(env):
    store, tfshim = env
    from gan_lib_tensorflow_b200.common.ops import linear as PL
    from oracle import ops as O

    # Linear: norms over axis 0 (linear.py:146)
    x = np.random.RandomState(4).standard_normal((32, 300)).astype("float32")
    gv = _g_values(128, 5)

    def prod_fn(xv):
        with store.variable_scope("Lin"):
            store.get_variable("g", initializer=gv)
        return PL.Linear(xv, 300, 128, "Lin", weightnorm=True)

    def orc_fn(g, xt):
        with g.variable_scope("Lin"):
            g.get_variable("g", initializer=gv)
        return O.Linear(g, xt, 300, 128, "Lin", weightnorm=True)

    prod, refs = run_pair(store, tfshim, prod_fn, orc_fn, x)
    check(prod, refs, tag="weightnorm linear")


@pytest.mark.parametrize("n,h,cin,cout,k", [(4, 8, 128, 64, 4), (3, 5, 72, 40, 4)])
def test_weightnorm_deconv2d(env, n, h, cin, cout, k):
    """deconv2d.py:87-96: one norm per OUTPUT channel = axis 2 of [k, k, Cout, Cin] (geometry with an inner axis)."""
    store, tfshim = env
    from gan_lib_tensorflow_b200.common.ops import deconv2d as P
    from oracle import ops as O

    x = _bf16_repr(np.random.RandomState(43).standard_normal((n, h, h, cin)).astype("float32"))
    cot = _bf16_repr(np.random.RandomState(44).standard_normal((n, 2 * h, 2 * h, cout)).astype("float32"))
    gv = _g_values(cout, 6)

    def prod_fn(xv):
        with store.variable_scope("L"):
            store.get_variable("g", initializer=gv)
        return P.Deconv2D(xv, cin, cout, k, name="L", weight_norm=True)

    def orc_fn(g, xt):
        with g.variable_scope("L"):
            g.get_variable("g", initializer=gv)
        return O.Deconv2D(g, xt, cin, cout, k, name="L", weight_norm=True)

    prod, refs = run_pair(store, tfshim, prod_fn, orc_fn, x, cot_np=cot, bf16=False)
    check(prod, refs, tag=f"weightnorm deconv {cin}->{cout}")   # effective filters are not bf16-representable: 1e-2


# ------------------------------------------------------------------------------------------------ PixelCNN masks
@pytest.mark.parametrize("mask_type,cin,cout,k,weightnorm", [
    (("a", 1), 64, 64, 3, False),
    (("b", 1), 64, 128, 5, False),
    (("a", 3), 72, 48, 3, False),      # three colour channels interleaved over the channel axis
    (("b", 3), 72, 72, 3, True),       # mask applied after weight-norm (conv2d.py:153-167)
])
def test_masked_conv2d(env, mask_type, cin, cout, k, weightnorm):
    store, tfshim = env
    from gan_lib_tensorflow_b200.common.ops import conv2d as P
    from oracle import ops as O

    x = np.random.RandomState(7).standard_normal((3, 8, 8, cin)).astype("float32")

    def prod_fn(xv):
        return P.Conv2D(xv, cin, cout, k, name="L", mask_type=mask_type, weightnorm=weightnorm)

    def orc_fn(g, xt):
        return O.Conv2D(g, xt, cin, cout, k, name="L", mask_type=mask_type, weightnorm=weightnorm)

    prod, refs = run_pair(store, tfshim, prod_fn, orc_fn, x)
    check(prod, refs, tag=f"mask {mask_type} {cin}->{cout} k{k} wn={weightnorm}")
    # masked taps receive no gradient (with weight-norm they do: the norm runs over ALL of W, conv2d.py:154-160)
    m = P.pixelcnn_mask(mask_type, k, cin, cout)
    dW = prod["params"]["L/Filters"]
    assert weightnorm or np.all(dW[m == 0] == 0.0)
    # same initial values as the oracle (the halved fans of conv2d.py:99-101)
    np.random.seed(0)
    g = tfshim.Graph(dtype=torch.float32, u_seed=2)
    O.Conv2D(g, torch.from_numpy(x), cin, cout, k, name="L", mask_type=mask_type, weightnorm=weightnorm)
    w_o = dict(g.trainable_variables())["L/Filters"].detach().numpy()
    assert np.array_equal(store.vars["L/Filters"].data.cpu().numpy(), w_o)


# ------------------------------------------------------------------------------------------------ layer norm
@pytest.mark.parametrize("shape,act,xdtype", [
    ((8, 16, 16, 128), "relu", torch.float32),
    ((5, 8, 8, 128), None, torch.float32),
    ((3, 32, 32, 64), "lrelu", torch.bfloat16),
    ((2, 4, 4, 256), "relu", torch.float32),
    ((64, 8, 8, 128), "relu", torch.bfloat16),
])
def test_layer_norm_act(env, shape, act, xdtype):
    """tf.contrib.layers.layer_norm(begin_norm_axis=1, begin_params_axis=-1) + the activation behind it, forward and
    backward (dx, dgamma, dbeta) against the float64 oracle."""
    store, tfshim = env
    from gan_lib_tensorflow_b200 import functional as F
    from gan_lib_tensorflow_b200.common.ops import normalization as P
    from oracle import ops as O

    rs = np.random.RandomState(8)
    c = shape[-1]
    x = (rs.standard_normal(shape) * 1.7 + 0.3).astype("float32")
    if xdtype == torch.bfloat16:
        x = _bf16_repr(x)
    cot = rs.standard_normal(shape).astype("float32")
    gam = (1.0 + 0.3 * rs.standard_normal(c)).astype("float32")
    bet = (0.2 * rs.standard_normal(c)).astype("float32")

    with store.variable_scope("LN"):
        store.get_variable("beta", initializer=bet)
        store.get_variable("gamma", initializer=gam)
    xv = F.Var(torch.from_numpy(x).cuda().to(xdtype), requires_grad=True, grad_dtype=torch.float32)
    with store.gradient_tape() as tape:
        out = P.layer_norm("LN", [1, 2, 3], xv, act=act, out_dtype=torch.float32)
        for v in store.vars.values():
            if v.trainable and v.grad is None:
                v.grad = torch.zeros_like(v.data)
        tape.backward(out, grad=torch.from_numpy(cot).cuda())
    torch.cuda.synchronize()

    g = tfshim.Graph(dtype=torch.float64, u_seed=2)
    with g.variable_scope("LN"):
        g.get_variable("beta", initializer=bet.astype("float64"))
        g.get_variable("gamma", initializer=gam.astype("float64"))
    xt = torch.from_numpy(x).double().requires_grad_(True)
    yo = O.layer_norm(g, "LN", [1, 2, 3], xt)
    if act == "relu":
        yo = torch.relu(yo)
    elif act == "lrelu":
        yo = torch.maximum(yo, 0.2 * yo)
    params = dict(g.trainable_variables())
    dx, dgam, dbet = torch.autograd.grad(yo, [xt, params["LN/gamma"], params["LN/beta"]],
                                         torch.from_numpy(cot).double())
    errs = dict(out=rel(out.data.cpu().numpy(), yo.detach().numpy()), dx=rel(xv.grad.float().cpu().numpy(), dx.numpy()),
                dgamma=rel(store.vars["LN/gamma"].grad.cpu().numpy(), dgam.numpy()),
                dbeta=rel(store.vars["LN/beta"].grad.cpu().numpy(), dbet.numpy()))
    print("layer_norm", shape, act, errs)
    for k, v in errs.items():
        assert v < 5e-5, (k, v)


def test_layer_normed_critic_block(env):
    """A 'down' residual block of the critic with NORMALIZATION_D (gan_cifar_resnet.py:99-100, 259-270)."""
    store, tfshim = env
    from gan_lib_tensorflow_b200.SNGAN import gan_cifar_resnet as PS
    from oracle import sngan_cifar as OS
    from tests.test_gpu_ops import TOL_BLOCK_FP32, TOL_BLOCK_IMPL

    x = _bf16_repr(np.random.RandomState(9).standard_normal((4, 16, 16, 128)).astype("float32"))
    old_p, old_o = PS.NORMALIZATION_D, OS.NORMALIZATION_D
    PS.NORMALIZATION_D = OS.NORMALIZATION_D = True
    try:
        def prod_fn(xv):
            return PS._block(xv, 128, 128, 3, "D.Block.2", resample="down", spectral_normed=True,
                             update_collection="NO_OPS")

        def orc_fn(g, xt):
            return OS._block(g, xt, 128, 128, 3, "D.Block.2", resample="down", spectral_normed=True,
                             update_collection="NO_OPS")

        prod, refs = run_pair(store, tfshim, prod_fn, orc_fn, x)
    finally:
        PS.NORMALIZATION_D, OS.NORMALIZATION_D = old_p, old_o
    assert any(k.endswith("N1/gamma") for k in refs["fp32"]["params"]), list(refs["fp32"]["params"])
    check(prod, refs, tol_impl=TOL_BLOCK_IMPL, tol_fp32=TOL_BLOCK_FP32, tag="layer-normed D block")


# ------------------------------------------------------------------------------------------------ fade-in
def test_lerp_device_alpha(env):
    store, _ = env
    from gan_lib_tensorflow_b200 import functional as F

    rs = np.random.RandomState(10)
    a = rs.standard_normal((3, 8, 8, 40)).astype("float32")
    b = rs.standard_normal((3, 8, 8, 40)).astype("float32")
    cot = rs.standard_normal(a.shape).astype("float32")
    alpha = torch.zeros(1, device="cuda")
    for val in (0.0, 0.3, 1.0):
        alpha.fill_(val)
        av = F.Var(torch.from_numpy(a).cuda(), requires_grad=True)
        bv = F.Var(torch.from_numpy(b).cuda().bfloat16(), requires_grad=True)
        with store.gradient_tape() as tape:
            out = F.lerp(av, bv, alpha)
            tape.backward(out, grad=torch.from_numpy(cot).cuda())
        torch.cuda.synchronize()
        b16 = _bf16_repr(b)
        assert rel(out.data.cpu().numpy(), (1 - val) * a + val * b16) < 1e-6
        assert rel(av.grad.float().cpu().numpy(), (1 - val) * cot) < (1e-6 if av.grad.dtype == torch.float32 else 4e-3) \
            or val == 1.0
        assert rel(bv.grad.float().cpu().numpy(), val * cot) < 4e-3 or val == 0.0
        if val == 1.0:
            assert float(av.grad.float().abs().max()) == 0.0
        if val == 0.0:
            assert float(bv.grad.float().abs().max()) == 0.0
    # a Python float is the same op
    out2 = F.lerp(F.Var(torch.from_numpy(a).cuda()), F.Var(torch.from_numpy(b).cuda()), 0.25)
    assert rel(out2.data.cpu().numpy(), 0.75 * a + 0.25 * b) < 1e-6


def test_pggan_captured_steps_follow_alpha(env):
    """PGGAN training ops as CUDA graphs (PGGAN/train.py:83, 184: alpha is fed per step): a replay with a NEW alpha
    must reproduce the eager step at that alpha from the same state."""
    store, _ = env
    from gan_lib_tensorflow_b200.PGGAN import train as PT

    b = 4
    tr = PT.Trainer(block_count=2, trans=True, inputs_norm=True, batch_size=b, seed=0)
    rs = np.random.RandomState(11)
    real = torch.from_numpy(rs.uniform(-1, 1, size=(b, tr.size, tr.size, 3)).astype("float32")).cuda()
    z = torch.from_numpy(rs.standard_normal((b, tr.z_dim)).astype("float32")).cuda()
    tr.d_step(real, z, 0.1)
    tr.g_step(z, 0.1)
    snap = Snapshot(store, tr.players, tr)

    ld = float(tr.d_step(real, z, 0.7).data.reshape(-1)[0])
    d_eager = _state(store, "d_net")
    lg = float(tr.g_step(z, 0.7).data.reshape(-1)[0])
    g_eager = _state(store, "g_net")

    snap.restore()
    tr.capture()                      # captured while alpha holds 0.0
    snap.restore()
    assert tr.players.captured("d") and tr.players.captured("g")
    ld_g = float(tr.d_step(real, z, 0.7).reshape(-1)[0])
    d_graph = _state(store, "d_net")
    lg_g = float(tr.g_step(z, 0.7).reshape(-1)[0])
    g_graph = _state(store, "g_net")
    print(f"pggan losses eager {ld:.6f} {lg:.6f} graph {ld_g:.6f} {lg_g:.6f}")
    assert abs(ld - ld_g) < 1e-4 * max(1.0, abs(ld)) and abs(lg - lg_g) < 1e-4 * max(1.0, abs(lg))
    _compare("pggan d_step", d_eager, d_graph)
    _compare("pggan g_step", g_eager, g_graph)
    # and a different alpha gives a different loss from the same state
    snap.restore()
    ld_other = float(tr.d_step(real, z, 0.2).reshape(-1)[0])
    assert abs(ld_other - ld_g) > 1e-6


# ------------------------------------------------------------------------------------------------ ResNet PGGAN pieces
@pytest.mark.parametrize("shape,dtype", [((3, 8, 8, 3), torch.float32), ((2, 16, 12, 3), torch.float32),
                                         ((2, 6, 4, 16), torch.bfloat16)])
def test_subsample2_nearest_half_resize(env, shape, dtype):
    """tf.image.resize_nearest_neighbor to half the size (common/resnet_block.py:286-287) and its gradient."""
    store, _ = env
    from gan_lib_tensorflow_b200 import functional as F
    from oracle import resnet_block as ORB

    rs = np.random.RandomState(12)
    x = _bf16_repr(rs.standard_normal(shape).astype("float32"))
    n, h, w, c = shape
    oh, ow = -(-h // 2), -(-w // 2)
    cot = _bf16_repr(rs.standard_normal((n, oh, ow, c)).astype("float32"))
    xv = F.Var(torch.from_numpy(x).cuda().to(dtype), requires_grad=True)
    with store.gradient_tape() as tape:
        out = F.subsample2(xv)
        tape.backward(out, grad=torch.from_numpy(cot).cuda().to(out.gdtype))
    torch.cuda.synchronize()
    xt = torch.from_numpy(x).requires_grad_(True)
    yo = ORB.resize_nearest(xt, oh, ow)
    (dx,) = torch.autograd.grad(yo, xt, torch.from_numpy(cot))
    assert np.array_equal(out.data.float().cpu().numpy(), yo.detach().numpy())
    assert np.array_equal(out.data.float().cpu().numpy(), x[:, ::2, ::2, :])
    assert np.array_equal(xv.grad.float().cpu().numpy(), dx.numpy())


@pytest.mark.parametrize("cin,cout,k", [(3, 128, 3), (3, 64, 1), (128, 128, 3)])
def test_inputs_norm_with_spectral_norm_conv(env, cin, cout, k):
    """The ResNet PGGAN critic combines inputs_norm and spectral_normed on every layer (common/resnet_block.py:276-337):
    y = (sqrt(2 / fan_in) / sigma) * conv(x, W) + b."""
    store, tfshim = env
    from gan_lib_tensorflow_b200.common.ops import conv2d as P
    from oracle import ops as O

    x = np.random.RandomState(13).standard_normal((4, 8, 8, cin)).astype("float32")
    kw = dict(inputs_norm=True, spectral_normed=True, update_collection="NO_OPS")
    prod, refs = run_pair(store, tfshim, lambda xv: P.Conv2D(xv, cin, cout, k, name="L", **kw),
                          lambda g, xt: O.Conv2D(g, xt, cin, cout, k, name="L", **kw), x)
    check(prod, refs, tag=f"inputs_norm+sn {cin}->{cout} k{k}")


def test_inputs_norm_with_spectral_norm_linear(env):
    """D.Output of the ResNet PGGAN critic: Linear 512 -> 1 with inputs_norm and spectral norm (:333-337)."""
    store, tfshim = env
    from gan_lib_tensorflow_b200.common.ops import linear as P
    from oracle import ops as O

    x = np.random.RandomState(14).standard_normal((6, 512)).astype("float32")
    kw = dict(inputs_norm=True, spectral_normed=True, update_collection="NO_OPS")
    prod, refs = run_pair(store, tfshim, lambda xv: P.Linear(xv, 512, 1, "L", **kw),
                          lambda g, xt: O.Linear(g, xt, 512, 1, "L", **kw), x)
    check(prod, refs, tag="inputs_norm+sn linear 512->1")


@pytest.mark.parametrize("bc,trans", [(1, True), (2, False)])
def test_resnet_pggan_forward_backward(env, bc, trans):
    """PGGAN/model_resnet.py: generator forward and critic forward + backward (all parameter gradients and the gradient
    wrt the image, which runs through the nearest half-resize of the skip path) against both oracles."""
    store, tfshim = env
    from gan_lib_tensorflow_b200 import functional as F
    from gan_lib_tensorflow_b200.PGGAN import model_resnet as P
    from oracle import ops as O_ops
    from oracle import pggan as OP

    n, alpha = 4, 0.3
    size = 4 * 2 ** bc
    rs = np.random.RandomState(15)
    z = rs.standard_normal((n, 512)).astype("float32")
    img = rs.uniform(-1, 1, size=(n, size, size, 3)).astype("float32")
    cot = rs.standard_normal(n).astype("float32")
    np.random.seed(0)
    pm = P.PGGAN(block_count=bc, trans=trans, inputs_norm=True)
    fake = pm.get_generator(torch.from_numpy(z).cuda(), alpha)
    xv = F.Var(torch.from_numpy(img).cuda(), requires_grad=True)
    with store.gradient_tape() as tape:
        logits = pm.get_discriminator(xv, alpha, update_collection="NO_OPS")
        for v in store.trainable_variables("d_net"):
            if v.grad is None:
                v.grad = torch.zeros_like(v.data)
        tape.backward(logits, grad=torch.from_numpy(cot).cuda())
    torch.cuda.synchronize()
    assert tuple(fake.shape) == (n, size, size, 3)
    got = dict(fake=fake.data.float().cpu().numpy(), logits=logits.data.float().cpu().numpy(),
               dx=xv.grad.float().cpu().numpy())
    got.update({v.key: v.grad.cpu().numpy() for v in store.trainable_variables("d_net")})
    refs = {}
    for mode in (True, False):
        O_ops.BF16_OPERANDS = mode
        try:
            np.random.seed(0)
            g = tfshim.Graph(dtype=torch.float32, u_seed=2)
            om = OP.PGGANResNet(bc, trans, True)
            with torch.no_grad():
                of = om.get_generator(g, torch.from_numpy(z), alpha)
            xt = torch.from_numpy(img).clone().requires_grad_(True)
            lo = om.get_discriminator(g, xt, alpha, update_collection=O_ops.NO_OPS)
            params = g.trainable_variables("d_net")
            grads = torch.autograd.grad(lo, [xt] + [p_ for _, p_ in params], torch.from_numpy(cot))
            r = dict(fake=of.numpy(), logits=lo.detach().numpy(), dx=grads[0].numpy())
            r.update({nm: gr.numpy() for (nm, _), gr in zip(params, grads[1:])})
            refs[mode] = r
        finally:
            O_ops.BF16_OPERANDS = False
    assert set(got) == set(refs[False])
    worst = {}
    for name in got:
        e_impl = rel(got[name], refs[True][name])
        e_prod = rel(got[name], refs[False][name])
        e_orc = rel(refs[True][name], refs[False][name])
        worst[name] = (e_impl, e_prod, e_orc)
        # the implementation follows the bf16-operand oracle, and sits no further from fp32 than that oracle does
        assert e_impl < 2e-2 and e_prod <= 1.5 * e_orc + 1e-2, (name, e_impl, e_prod, e_orc)
    top = sorted(worst.items(), key=lambda kv: -kv[1][0])[:3]
    print(f"resnet pggan bc={bc} trans={trans}: worst vs bf16-oracle", [(k, f"{v[0]:.1e}") for k, v in top],
          "fake", f"{worst['fake'][0]:.1e}", "logits", f"{worst['logits'][0]:.1e}")
    assert worst["logits"][0] < 4e-3 and worst["fake"][0] < 2e-2


# ------------------------------------------------------------------------------------------------ depthwise / separable
@pytest.mark.parametrize("n,h,cin,cm,k,stride,padding", [
    (3, 16, 64, 2, 3, 1, "SAME"),
    (2, 16, 72, 1, 4, 2, "SAME"),       # Pix2Pix encoder geometry (4x4 s2), ragged channels
    (2, 9, 8, 3, 3, 2, "VALID"),
    (2, 16, 3, 4, 4, 2, (1, 1, 1, 1)),  # RGB input, explicit pads (PatchGAN layer_1 geometry)
])
def test_depthwise_conv2d(env, n, h, cin, cm, k, stride, padding):
    """conv_type='depthwise_conv2d' (common/ops/conv2d.py:188-197): forward, dx, d(depthwise_filters), dBiases.  Input
    and cotangent are bf16-representable, so the fp32 oracle is the exact reference of the fp32-arithmetic kernels."""
    store, tfshim = env
    from gan_lib_tensorflow_b200.common.ops import conv2d as P
    from oracle import ops as O

    x = _bf16_repr(np.random.RandomState(20).standard_normal((n, h, h, cin)).astype("float32"))
    kw = dict(conv_type="depthwise_conv2d", channel_multiplier=cm, padding=padding)
    prod, refs = run_pair(store, tfshim, lambda xv: P.Conv2D(xv, cin, cin * cm, k, stride, "L", **kw),
                          lambda g, xt: O.Conv2D(g, xt, cin, cin * cm, k, stride, "L", **kw), x, bf16=False)
    assert "L/depthwise_filters" in refs["fp32"]["params"] and "L/Filters" not in refs["fp32"]["params"]
    check(prod, refs, tol_fp32=2e-3, tag=f"depthwise {cin}x{cm} k{k} s{stride} {padding}")   # dx is stored in bf16
    with pytest.raises(ValueError):     # tf.nn.bias_add fails unless output_dim == input_dim * channel_multiplier
        P.Conv2D(torch.zeros(1, 8, 8, cin).cuda(), cin, cin * cm + 8, k, stride, "Bad", **kw)


@pytest.mark.parametrize("n,h,cin,cout,cm,k,stride,sn", [
    (3, 16, 64, 128, 1, 4, 2, False),
    (2, 16, 64, 64, 2, 4, 1, True),      # spectral_normed: depthwise AND pointwise filters are normalised (conv2d.py:169-178)
    (3, 16, 64, 128, 1, 4, 2, True),     # PatchGAN layer shape of Pix2Pix with --conv_type separable_conv2d
    (2, 32, 3, 64, 4, 4, 2, False),      # RGB input: 12 depthwise channels into the pointwise conv
    (2, 8, 128, 8, 1, 3, 1, False),
])
def test_separable_conv2d(env, n, h, cin, cout, cm, k, stride, sn):
    """conv_type='separable_conv2d' (conv2d.py:198-208): depthwise kernel + 1x1 pointwise conv on the tensor cores."""
    store, tfshim = env
    from gan_lib_tensorflow_b200.common.ops import conv2d as P
    from oracle import ops as O

    x = np.random.RandomState(21).standard_normal((n, h, h, cin)).astype("float32")
    kw = dict(conv_type="separable_conv2d", channel_multiplier=cm, spectral_normed=sn, update_collection="NO_OPS")
    prod, refs = run_pair(store, tfshim, lambda xv: P.Conv2D(xv, cin, cout, k, stride, "L", **kw),
                          lambda g, xt: O.Conv2D(g, xt, cin, cout, k, stride, "L", **kw), x)
    assert {"L/depthwise_filters", "L/pointwise_filters", "L/Biases"} <= set(refs["fp32"]["params"])
    assert float(np.abs(prod["params"]["L/Filters"]).max()) == 0.0        # `Filters` exists but is never read
    if sn:
        for scope in ("filters", "depthwise_filters", "pointwise_filters"):
            assert "L/%s/spectral_norm/u" % scope in store.vars
        # sigma really is applied: the un-normalised layer gives a different result
        from gan_lib_tensorflow_b200 import functional as F
        np.random.seed(0)
        plain = P.Conv2D(F.Var(torch.from_numpy(x).cuda()), cin, cout, k, stride, "L", conv_type="separable_conv2d",
                         channel_multiplier=cm, spectral_normed=False).data.float().cpu().numpy()
        assert rel(plain, refs["fp32"]["out"]) > 0.03
    check(prod, refs, tol_impl=3e-3, tag=f"separable {cin}->{cout} cm{cm} k{k} s{stride} sn={sn}")


def test_spectral_normed_depthwise_conv2d(env):
    """conv_type='depthwise_conv2d' with spectral_normed (conv2d.py:173-175, 188-197): W_d / sigma_d is what the op
    reads; u of the depthwise filter follows update_collection, the u of `Filters` / `pointwise_filters` (whose
    normalised values no op reads) keep their initial values."""
    store, tfshim = env
    from gan_lib_tensorflow_b200.common.ops import conv2d as P
    from oracle import ops as O

    n, h, cin, cm, k, stride = 2, 16, 64, 2, 3, 1
    x = _bf16_repr(np.random.RandomState(23).standard_normal((n, h, h, cin)).astype("float32"))
    kw = dict(conv_type="depthwise_conv2d", channel_multiplier=cm, spectral_normed=True, update_collection=None)
    prod, refs = run_pair(store, tfshim, lambda xv: P.Conv2D(xv, cin, cin * cm, k, stride, "L", **kw),
                          lambda g, xt: O.Conv2D(g, xt, cin, cin * cm, k, stride, "L", **kw), x, bf16=False)
    check(prod, refs, tol_fp32=2e-3, tag="depthwise spectral_normed")
    u_rng = np.random.RandomState(2)          # the store's u stream (u_seed=2): filters, depthwise, pointwise
    from gan_lib_tensorflow_b200.framework import truncated_normal
    u_f, u_d, u_p = (truncated_normal([1, c], u_rng) for c in (cin * cm, cm, cin * cm))
    got = {s: store.vars["L/%s/spectral_norm/u" % s].data.cpu().numpy() for s in
           ("filters", "depthwise_filters", "pointwise_filters")}
    assert np.array_equal(got["filters"], u_f) and np.array_equal(got["pointwise_filters"], u_p)
    assert not np.allclose(got["depthwise_filters"], u_d)          # u <- u' (update_collection=None)


def test_pix2pix_patchgan_with_separable_convs(env):
    """Pix2Pix --conv_type separable_conv2d --channel_multiplier 1 through unet_d (Pix2Pix/train.py:31-33, 464-465)."""
    store, tfshim = env
    from gan_lib_tensorflow_b200 import functional as F
    from gan_lib_tensorflow_b200.Pix2Pix import networks as P
    from oracle import ops as O
    from oracle import pix2pix as OP

    n, ndf = 2, 16
    rs = np.random.RandomState(22)
    x = rs.uniform(-1, 1, size=(n, 64, 64, 3)).astype("float32")
    tgt = rs.uniform(-1, 1, size=(n, 64, 64, 3)).astype("float32")
    prod, refs = run_pair(
        store, tfshim,
        lambda xv: P.unet_d(xv, F.Var(torch.from_numpy(tgt).cuda()), ndf, True, "NO_OPS",
                            conv_type="separable_conv2d", channel_multiplier=1),
        lambda g, xt: OP.unet_d(g, xt, torch.from_numpy(tgt), ndf, True, O.NO_OPS, conv_type="separable_conv2d",
                                channel_multiplier=1), x)
    assert prod["out"].shape == (n, 6, 6, 1)
    check(prod, refs, tol_impl=1.2e-2, tol_fp32=1.2e-1, tag="unet_d separable")
