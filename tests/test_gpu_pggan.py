"""GPU parity tests for PGGAN (SURVEY 8(a) a-16, config 5 of BASELINE.json): PGGAN/model_nvidia.py generator and
discriminator incl. inputs_norm, the fade-in skip connections, minibatch-stddev and the (C+1)-channel convolution
behind it -- CUDA path through the C ABI against the CPU oracle (oracle/pggan.py) on the same seeded inputs."""
import numpy as np
import pytest
import torch

from tests.test_gpu_ops import check, env, rel, run_pair  # noqa: F401
from tests.test_gpu_wide import check_band

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,cin,cout,sn", [(4, 513, 512, True), (3, 77, 40, False)])
def test_ragged_input_channel_conv(env, n, cin, cout, sn):
    """3x3 convolution whose input-channel count is not a multiple of 8 (model_nvidia.py:226: 512 + 1 std channel)."""
    store, tfshim = env
    from gan_lib_tensorflow_b200.common.ops import conv2d as P
    from oracle import ops as O

    x = np.random.RandomState(3).standard_normal((n, 4, 4, cin)).astype("float32")
    uc = "NO_OPS" if sn else None
    prod, refs = run_pair(
        store, tfshim,
        lambda xv: P.Conv2D(xv, cin, cout, 3, 1, "D.Conv", spectral_normed=sn, update_collection=uc),
        lambda g, xt: O.Conv2D(g, xt, cin, cout, 3, 1, "D.Conv", spectral_normed=sn,
                               update_collection=O.NO_OPS if sn else None), x)
    check(prod, refs, tag=f"ragged cin={cin}")


@pytest.mark.parametrize("cin,cout,k", [(64, 128, 3), (3, 64, 1), (128, 3, 1)])
def test_inputs_norm_conv_and_linear(env, cin, cout, k):
    """inputs_norm (conv2d.py:93-95, linear.py:47-49) carried by the GEMM epilogue's alpha."""
    store, tfshim = env
    from gan_lib_tensorflow_b200.common.ops import conv2d as P
    from oracle import ops as O

    x = np.random.RandomState(4).standard_normal((3, 8, 8, cin)).astype("float32")
    prod, refs = run_pair(
        store, tfshim,
        lambda xv: P.Conv2D(xv, cin, cout, k, 1, "c", inputs_norm=True),
        lambda g, xt: O.Conv2D(g, xt, cin, cout, k, 1, "c", inputs_norm=True), x)
    check(prod, refs, tag=f"inputs_norm {cin}->{cout} k{k}")


def test_inputs_norm_linear(env):
    store, tfshim = env
    from gan_lib_tensorflow_b200 import functional as F  # noqa: F401
    from gan_lib_tensorflow_b200.common.ops import linear as P
    from oracle import ops as O

    x = np.random.RandomState(5).standard_normal((6, 512)).astype("float32")
    prod, refs = run_pair(
        store, tfshim,
        lambda xv: P.Linear(xv, 512, 8192, "G.Input", inputs_norm=True),
        lambda g, xt: O.Linear(g, xt, 512, 8192, "G.Input", inputs_norm=True), x)
    check(prod, refs, tag="inputs_norm linear")


@pytest.mark.parametrize("bc,trans,inputs_norm", [(2, True, True), (1, False, False), (0, False, True)])
def test_pggan_generator(env, bc, trans, inputs_norm):
    """model_nvidia.py:73-129 at full width (512 channels), 4x4 -> 4 * 2^bc, fade-in alpha = 0.3."""
    store, tfshim = env
    from gan_lib_tensorflow_b200.PGGAN import model_nvidia as P
    from oracle import pggan as OP

    n, alpha = 4, 0.3
    z = np.random.RandomState(71).standard_normal((n, 512)).astype("float32")
    pm = P.PGGAN(block_count=bc, trans=trans, inputs_norm=inputs_norm)
    om = OP.PGGAN(bc, trans, inputs_norm)
    prod, refs = run_pair(store, tfshim, lambda zv: pm.get_generator(zv, alpha),
                          lambda g, zt: om.get_generator(g, zt, alpha), z)
    size = 4 * 2 ** bc
    assert prod["out"].shape == (n, size, size, 3)
    assert set(prod["params"]) == set(refs["fp32"]["params"])
    assert rel(prod["out"], refs["bf16"]["out"]) < 5e-3
    check_band(prod, refs, tag=f"pggan_g bc={bc}")


@pytest.mark.parametrize("bc,trans", [(2, True), (1, False), (0, False)])
def test_pggan_discriminator(env, bc, trans):
    """model_nvidia.py:164-237: fromRGB (+ fade-in), spectrally-normalised blocks, minibatch-stddev, 513-channel conv."""
    store, tfshim = env
    from gan_lib_tensorflow_b200.PGGAN import model_nvidia as P
    from oracle import ops as O
    from oracle import pggan as OP

    n, alpha = 4, 0.3
    size = 4 * 2 ** bc
    rs = np.random.RandomState(72)
    x = rs.uniform(-1, 1, size=(n, size, size, 3)).astype("float32")
    pm = P.PGGAN(block_count=bc, trans=trans, inputs_norm=False)
    om = OP.PGGAN(bc, trans, False)
    prod, refs = run_pair(
        store, tfshim,
        lambda xv: pm.get_discriminator(xv, alpha, spectral_normed=True, update_collection="NO_OPS"),
        lambda g, xt: om.get_discriminator(g, xt, alpha, spectral_normed=True, update_collection=O.NO_OPS),
        x, cot_np=rs.standard_normal((n,)).astype("float32"))
    assert prod["out"].shape == (n,)
    assert set(prod["params"]) == set(refs["fp32"]["params"])
    assert rel(prod["out"], refs["bf16"]["out"]) < 5e-3
    check_band(prod, refs, tag=f"pggan_d bc={bc}")


# ------------------------------------------------------------------------------------------------ ACGAN (config 2)
def test_acgan_generator(env):
    """ACGAN/model.py:27-57: conditional-BN generator (library Normalize dispatch), batch 8."""
    store, tfshim = env
    from gan_lib_tensorflow_b200.ACGAN import model as P
    from oracle import acgan as OA

    n = 8
    rs = np.random.RandomState(81)
    z = rs.standard_normal((n, 128)).astype("float32")
    labels = rs.randint(0, 10, size=n).astype("int32")
    prod, refs = run_pair(
        store, tfshim,
        lambda zv: P.ACGAN().get_generator(zv, torch.from_numpy(labels).cuda()),
        lambda g, zt: OA.ACGAN().get_generator(g, zt, torch.from_numpy(labels).long()), z)
    assert prod["out"].shape == (n, 32, 32, 3)
    assert set(prod["params"]) == set(refs["fp32"]["params"])
    assert rel(prod["out"], refs["bf16"]["out"]) < 1e-2
    check_band(prod, refs, tag="acgan_g")


@pytest.mark.parametrize("loss_type", ["HINGE", "WGAN", "LSGAN", "CGAN", "Modified_MiniMax", "MiniMax"])
def test_acgan_discriminator_step_losses(env, loss_type):
    """ACGAN/model.py:59-90 + the critic / auxiliary-classifier losses of ACGAN/train.py:89-115 (without the gradient
    penalty) for every loss_type of common/misc.py:310-394: D on 8 real + 8 fake images, gradient of
    d_loss = d_loss_gan + d_loss_acgan; and the generator-side loss values on the same logits."""
    store, tfshim = env
    from gan_lib_tensorflow_b200 import functional as F
    from gan_lib_tensorflow_b200.ACGAN import model as P
    from oracle import acgan as OA
    from oracle import ops as O_ops

    n = 8
    rs = np.random.RandomState(82)
    real = rs.uniform(-1, 1, size=(n, 32, 32, 3)).astype("float32")
    fake = rs.uniform(-1, 1, size=(n, 32, 32, 3)).astype("float32")
    rl = rs.randint(0, 10, size=n).astype("int32")
    fl = rs.randint(0, 10, size=n).astype("int32")
    # ---- product
    np.random.seed(0)
    pm = P.ACGAN()
    rl_d, fl_d = torch.from_numpy(rl).cuda(), torch.from_numpy(fl).cuda()
    with store.gradient_tape() as tape:
        d_real, a_real = pm.get_discriminator(F.Var(torch.from_numpy(real).cuda()), rl_d, update_collection=None)
        fv = F.Var(torch.from_numpy(fake).cuda(), requires_grad=True)
        d_fake, a_fake = pm.get_discriminator(fv, fl_d, update_collection="NO_OPS", reuse=True)
        d_loss, parts = P.discriminator_losses(d_real, a_real, rl_d, d_fake, loss_type=loss_type)
        for v in store.vars.values():
            if v.trainable and v.grad is None:
                v.grad = torch.zeros_like(v.data)
        tape.backward(d_loss)
    g_loss, gparts = P.generator_losses(d_fake, a_fake, fl_d, loss_type=loss_type, acgan_scale_G=0.1)
    torch.cuda.synchronize()
    got = {k: v.grad.cpu().numpy() for k, v in store.vars.items() if v.trainable}
    got_losses = [float(d_loss.data.item()), float(parts["d_loss_acgan"].item()), float(g_loss.data.item())]
    with pytest.raises(NotImplementedError):
        P.discriminator_losses(d_real, a_real, rl_d, d_fake, gradient_penalty=True)
    # ---- oracles
    res = {}
    for mode in (True, False):
        O_ops.BF16_OPERANDS = mode
        np.random.seed(0)
        g = tfshim.Graph(dtype=torch.float32, u_seed=2)
        om = OA.ACGAN()
        o_real, oa_real = om.get_discriminator(g, torch.from_numpy(real), torch.from_numpy(rl).long())
        ft = torch.from_numpy(fake).clone().requires_grad_(True)
        o_fake, oa_fake = om.get_discriminator(g, ft, torch.from_numpy(fl).long(), update_collection=O_ops.NO_OPS,
                                               reuse=True)
        dl_gan, gl_gan = OA.get_loss(o_real, o_fake, loss_type)
        dl_ac = OA.sparse_softmax_xent_mean(oa_real, torch.from_numpy(rl))
        gl = gl_gan + 0.1 * OA.sparse_softmax_xent_mean(oa_fake, torch.from_numpy(fl))
        params = g.trainable_variables()
        grads = torch.autograd.grad(dl_gan + dl_ac, [ft] + [p for _, p in params])
        res[mode] = {"losses": [float(dl_gan + dl_ac), float(dl_ac), float(gl)], "dx": grads[0].numpy(),
                     "params": {name: gr.numpy() for (name, _), gr in zip(params, grads[1:])}}
    O_ops.BF16_OPERANDS = False
    assert set(got) == set(res[False]["params"])
    for mode in (True, False):
        assert max(abs(a - b) / (abs(b) + 1e-3) for a, b in zip(got_losses, res[mode]["losses"])) < 1e-2
    # four batch-normed blocks deep at batch 8: measured band (tests/test_gpu_wide.py::check_band) -- the product must
    # be as close to the fp32 oracle as the bf16-operand oracle is, tensor by tensor
    prod = {"out": np.asarray(got_losses), "dx": fv.grad.float().cpu().numpy(), "params": got}
    refs = {k: {"out": np.asarray(res[m]["losses"]), "dx": res[m]["dx"], "params": res[m]["params"]}
            for k, m in (("bf16", True), ("fp32", False))}
    check_band(prod, refs, tag=f"acgan_d {loss_type}")
    # the head layers sit behind the whole (compounding) forward pass but in front of no backward depth
    for name in ("d_net/D.Output/W", "d_net/D.ACGANOutput/W", "d_net/D.NoneBlock.4.Conv2/Filters"):
        assert rel(got[name], res[True]["params"][name]) < 5e-2, name


@pytest.mark.parametrize("resample,cin,cout,h,act", [
    ("down", 128, 128, 16, "lrelu"), (None, 128, 128, 8, "lrelu"), (None, 64, 128, 8, "lrelu"),
    (None, 128, 128, 8, "relu"),
])
def test_batch_normed_discriminator_block(env, resample, cin, cout, h, act):
    """ResidualBlock with spectral_normed=False under a 'D.' name: the library dispatch gives plain batch norm
    (common/resnet_block.py:36-39) -- ACGAN's D blocks; fp32 residual stream in, leaky ReLU."""
    store, tfshim = env
    from gan_lib_tensorflow_b200.common import resnet_block as P
    from oracle import resnet_block as ORB
    from tests.test_gpu_ops import TOL_BLOCK_FP32, TOL_BLOCK_IMPL

    x = np.random.RandomState(23).standard_normal((8, h, h, cin)).astype("float32")
    prod, refs = run_pair(
        store, tfshim,
        lambda xv: P.ResidualBlock(xv, cin, cout, 3, "D.B", spectral_normed=False, resample=resample, activation_fn=act),
        lambda g, xt: ORB.ResidualBlock(g, xt, cin, cout, 3, "D.B", spectral_normed=False, resample=resample,
                                        activation_fn=act), x)
    check(prod, refs, tol_impl=TOL_BLOCK_IMPL, tol_fp32=TOL_BLOCK_FP32, tag=f"bn block {resample} {act}")


def test_first_block_lrelu_unnormalised(env):
    store, tfshim = env
    from gan_lib_tensorflow_b200.common import resnet_block as P
    from oracle import resnet_block as ORB
    from tests.test_gpu_ops import TOL_BLOCK_FP32, TOL_BLOCK_IMPL

    x = np.random.RandomState(24).uniform(-1, 1, size=(6, 32, 32, 3)).astype("float32")
    prod, refs = run_pair(
        store, tfshim,
        lambda xv: P.OptimizedResBlockDisc1(xv, 128, activation_fn="lrelu"),
        lambda g, xt: ORB.OptimizedResBlockDisc1(g, xt, 128, activation_fn="lrelu"), x)
    check(prod, refs, tol_impl=TOL_BLOCK_IMPL, tol_fp32=TOL_BLOCK_FP32, tag="first block lrelu")


def _band(got, ref_b16, ref_f32, factor=2.0, floor=5e-3):
    gmax = max(np.linalg.norm(t) for t in ref_f32.values())
    for name, f32 in ref_f32.items():
        if np.linalg.norm(f32) < 5e-2 * gmax:
            continue
        e_prod, e_orc = rel(got[name], f32), rel(ref_b16[name], f32)
        assert e_prod <= factor * e_orc + floor, (name, e_prod, e_orc)


@pytest.mark.parametrize("bc,trans,model", [(1, False, "nvidia"), (2, True, "nvidia"), (1, False, "resnet"),
                                            (2, True, "resnet")])
def test_pggan_training_steps(env, bc, trans, model):
    """PGGAN/train.py:103-136: critic and generator gradients of the hinge losses (D(real) with update_collection=None,
    fake branch NO_OPS, alpha fade-in) through PGGAN.train.Trainer vs the oracle; then one Adam step moves both."""
    store, tfshim = env
    from gan_lib_tensorflow_b200.PGGAN import train as PT
    from oracle import ops as O_ops
    from oracle import pggan as OP

    n, alpha = 4, 0.4
    size = 4 * 2 ** bc
    rs = np.random.RandomState(91)
    real = rs.uniform(-1, 1, size=(n, size, size, 3)).astype("float32")
    z = rs.standard_normal((n, 512)).astype("float32")
    tr = PT.Trainer(bc, trans, inputs_norm=True, batch_size=n, seed=0, model=model)
    real_d, z_d = torch.from_numpy(real).cuda(), torch.from_numpy(z).cuda()
    dl = tr.players.gradients("d", lambda: tr.d_loss(real_d, z_d, alpha))
    d_grads = {v.key: v.grad.cpu().numpy().copy() for v in store.trainable_variables("d_net")}
    u_after = {k: v.data.cpu().numpy().copy() for k, v in store.vars.items() if k.endswith("/u")}
    gl = tr.players.gradients("g", lambda: tr.g_loss(z_d, alpha))
    g_grads = {v.key: v.grad.cpu().numpy().copy() for v in store.trainable_variables("g_net")}
    d_loss, g_loss = float(dl.data.item()), float(gl.data.item())
    refs = {}
    for mode in (True, False):
        O_ops.BF16_OPERANDS = mode
        try:
            np.random.seed(0)
            g = tfshim.Graph(dtype=torch.float32, u_seed=2)
            ol = OP.PGGANLosses(g, bc, trans, True, size, model=model)
            dc, dp, dg = ol.d_grads(torch.from_numpy(real), torch.from_numpy(z), alpha)
            u_ref = {k_: v.detach().numpy().copy() for k_, v in g.vars.items() if k_.endswith("/u")}
            gc, gp, gg = ol.g_grads(torch.from_numpy(z), alpha)
            refs[mode] = dict(d=dc.item(), g=gc.item(), u=u_ref,
                              dg={nm: t.numpy() for (nm, _), t in zip(dp, dg) if t is not None},
                              gg={nm: t.numpy() for (nm, _), t in zip(gp, gg) if t is not None})
        finally:
            O_ops.BF16_OPERANDS = False
    assert set(d_grads) == set(refs[False]["dg"]) and set(g_grads) == set(refs[False]["gg"])
    assert abs(d_loss - refs[True]["d"]) < 2e-3 and abs(g_loss - refs[True]["g"]) < 2e-3
    for name, u in refs[False]["u"].items():          # u <- u' by the D(real) pass only
        assert rel(u_after[name], u) < 1e-4, name
    _band(d_grads, refs[True]["dg"], refs[False]["dg"])
    _band(g_grads, refs[True]["gg"], refs[False]["gg"])
    # one reference iteration (G step, then D steps) runs and moves the parameters
    before = store.flat["g_net"].params.clone()
    d, g_ = tr.train_iteration(1, iter([real_d] * 2), n_dis=2)
    torch.cuda.synchronize()
    assert np.isfinite(d.data.item()) and np.isfinite(g_.data.item())
    assert not torch.equal(before, store.flat["g_net"].params)


def test_acgan_trainer_runs_the_reference_iteration_without_the_penalty(env):
    """ACGAN/train.py:191-203 loop structure (G step skipped at step 0, n_dis critic steps, LR decay on the generator's
    global step), here without the penalty term (its parity is test_acgan_gradient_penalty)."""
    store, _ = env
    from gan_lib_tensorflow_b200.ACGAN import train as AT

    tr = AT.Trainer(batch_size=8, gradient_penalty=False, seed=0, max_iter=10)
    rs = np.random.RandomState(3)
    data = torch.from_numpy(rs.randint(0, 256, size=(8, 3072)).astype("int32")).cuda()
    labels = torch.from_numpy(rs.randint(0, 10, size=8).astype("int32")).cuda()
    real = tr.preprocess(data, None)
    assert real.shape == (8, 32, 32, 3) and float(real.min()) >= -1.0 and float(real.max()) < 1.0
    batches = iter([(real, labels)] * 8)
    d0, g0 = tr.train_iteration(0, batches, n_dis=2)
    assert g0 is None and tr.global_step == 0 and abs(tr.learning_rate() - 0.0004) < 1e-12
    before = store.flat["g_net"].params.clone()
    d1, g1 = tr.train_iteration(1, batches, n_dis=2)
    torch.cuda.synchronize()
    assert np.isfinite(d1.data.item()) and np.isfinite(g1.data.item())
    assert tr.global_step == 1 and tr.learning_rate() < 0.0004
    assert not torch.equal(before, store.flat["g_net"].params)
    assert set(tr.last_d) == {"d_loss_gan", "d_loss_acgan"}


def test_acgan_gradient_penalty(env):
    """ACGAN/train.py:97-105: WGAN-GP on the interpolates, differentiated w.r.t. D's parameters THROUGH the backward
    pass (grad-grad of six batch norms, ganb_bn_bwd_vjp).  Value and every parameter gradient of the penalty alone vs
    the fp32 oracle (torch double backward); the product runs its convolutions on bf16 operands in both passes."""
    store, tfshim = env
    from gan_lib_tensorflow_b200.ACGAN import gp as GP
    from gan_lib_tensorflow_b200.ACGAN import train as AT
    from oracle import acgan as OA
    from oracle import ops as O_ops
    from tests.test_gpu_ops import _report

    n = 8
    rs = np.random.RandomState(97)
    real = rs.uniform(-1, 1, size=(n, 32, 32, 3)).astype("float32")
    fake = rs.uniform(-1, 1, size=(n, 32, 32, 3)).astype("float32")
    alpha = rs.uniform(0, 1, size=n).astype("float32")
    labels = rs.randint(0, 10, size=n).astype("int32")
    tr = AT.Trainer(batch_size=n, gradient_penalty=True, seed=0)
    # perturb gamma / beta so that the batch-norm parameters matter
    prs = np.random.RandomState(5)
    pert = {}
    for k, v in store.vars.items():
        if "BatchNorm" in k:
            pert[k] = (v.data.cpu().numpy() + 0.2 * prs.standard_normal(tuple(v.data.shape))).astype("float32")
            v.data.copy_(torch.from_numpy(pert[k]))
    store.zero_grad("d_net")
    with store.gradient_tape() as tape, store.frozen_scopes("g_net"):
        gp = GP.gradient_penalty(torch.from_numpy(real).cuda(), torch.from_numpy(fake).cuda(),
                                 torch.from_numpy(alpha).cuda())
        tape.backward(gp)
    torch.cuda.synchronize()
    got = {v.key: v.grad.cpu().numpy().copy() for v in store.trainable_variables("d_net")}
    gp_val = float(gp.data.item())
    # ---- oracles: fp32 (the reference formula) and bf16-operand (rounded convolutions, differentiated twice)
    refs = {}
    lab = torch.from_numpy(labels).long()
    for mode in (False, True):
        O_ops.BF16_OPERANDS = mode
        try:
            np.random.seed(0)
            g = tfshim.Graph(dtype=torch.float32, u_seed=2)
            om = OA.ACGAN()
            with torch.no_grad():
                om.get_discriminator(g, torch.zeros(2, 32, 32, 3), lab[:2])
                om.get_generator(g, torch.zeros(2, 128), lab[:2])
            for k, arr in pert.items():
                g.assign(k, torch.from_numpy(arr))
            ref = OA.gradient_penalty(g, om, torch.from_numpy(real), torch.from_numpy(fake), torch.from_numpy(alpha), lab)
            params = g.trainable_variables("d_net")
            grads = torch.autograd.grad(ref, [p for _, p in params], allow_unused=True)
            refs[mode] = (float(ref.detach()), {nm: (t.numpy() if t is not None else None)
                                                for (nm, _), t in zip(params, grads)})
        finally:
            O_ops.BF16_OPERANDS = False
    f32_val, f32 = refs[False]
    b16_val, b16 = refs[True]
    gmax = max(np.linalg.norm(t) for t in f32.values() if t is not None)
    errs = {}
    for nm, t in f32.items():
        if t is None or np.linalg.norm(t) < 5e-2 * gmax:
            continue
        errs[nm] = (rel(got[nm], t), rel(b16[nm], t))
    _report(f"acgan gradient penalty: value {gp_val:.5f} vs fp32 oracle {f32_val:.5f} / bf16 oracle {b16_val:.5f}; "
            "prod-vs-fp32/bf16oracle-vs-fp32: " +
            " ".join(f"{k.split('/', 1)[1]}={v[0]:.2e}/{v[1]:.2e}" for k, v in sorted(errs.items(), key=lambda kv: -kv[1][0])[:10]))
    assert abs(gp_val - f32_val) < 1e-2 * abs(f32_val) + 1e-4
    # batch-8 batch norms differentiated twice: the bf16 band is wide (the two ORACLES differ by 15-20 %); the product
    # must be as close to fp32 as the bf16-operand oracle is.  The pieces are held tightly by the unit tests below.
    assert len(errs) >= 8
    for nm, (e_prod, e_orc) in errs.items():
        assert e_prod <= 1.5 * e_orc + 1e-2, (nm, e_prod, e_orc)
    # heads that take no part in the critic logit's input gradient get no penalty gradient
    assert np.abs(got["d_net/D.ACGANOutput/W"]).max() == 0.0
    # and the full critic loss of the trainer includes the term
    d = tr.d_loss(torch.from_numpy(real).cuda(), torch.from_numpy(labels).cuda(), torch.randn(n, 128, device="cuda"),
                  torch.from_numpy(labels).cuda(), torch.from_numpy(alpha).cuda())
    assert "gradient_penalty" in tr.last_d and np.isfinite(d.data.item())


def test_second_order_ops_match_torch_double_backward(env):
    """The differentiable backward ops of functional.py ("second order") one by one, fp32-exact where the op is fp32:
    bn_act_input_grad (ganb_bn_bwd_vjp: grad-grad of a training-mode batch norm + leaky relu) against torch's double
    backward; conv2d_input_grad's two VJPs with bf16-representable operands."""
    store, tfshim = env
    from gan_lib_tensorflow_b200 import functional as F
    from gan_lib_tensorflow_b200 import kernels as K
    from gan_lib_tensorflow_b200.common.ops import conv2d as C
    from tests.test_gpu_ops import _bf16_repr

    rs = np.random.RandomState(11)
    # ---------------- batch norm + lrelu: gx = d<gy, lrelu(BN(x))>/dx ; L2 = <cot, gx>
    n, h, c = 6, 8, 32
    x = (rs.standard_normal((n, h, h, c)) * 1.5 + 0.3).astype("float32")
    gy = rs.standard_normal((n, h, h, c)).astype("float32")
    cot = rs.standard_normal((n, h, h, c)).astype("float32")
    gam = (1 + 0.3 * rs.standard_normal(c)).astype("float32")
    bet = (0.2 * rs.standard_normal(c)).astype("float32")
    with store.variable_scope("t"):
        gv = store.get_variable("gamma", initializer=gam)
        bv = store.get_variable("beta", initializer=bet)
    gv.grad, bv.grad = torch.zeros_like(gv.data), torch.zeros_like(bv.data)
    xv = F.Var(torch.from_numpy(x).cuda(), requires_grad=True)
    gyv = F.Var(torch.from_numpy(gy).cuda(), requires_grad=True)
    mean, rstd = K.bn_stats(xv.data, n, h * h, c, 1, 1e-5)
    with store.gradient_tape() as tape:
        gx = F.bn_act_input_grad(xv, gyv, gv, bv, mean, rstd, 'lrelu')
        tape.backward(gx, grad=torch.from_numpy(cot).cuda())
    xt = torch.from_numpy(x).double().requires_grad_(True)
    gyt = torch.from_numpy(gy).double().requires_grad_(True)
    gt = torch.from_numpy(gam).double().requires_grad_(True)
    bt = torch.from_numpy(bet).double()
    mu = xt.mean(dim=(0, 1, 2)); var = xt.var(dim=(0, 1, 2), unbiased=False)
    z = (xt - mu) * torch.rsqrt(var + 1e-5) * gt + bt
    y = torch.where(z >= 0, z, 0.2 * z)
    gx_ref, = torch.autograd.grad(y, xt, gyt, create_graph=True)
    d_x, d_gy, d_g = torch.autograd.grad(gx_ref, (xt, gyt, gt), torch.from_numpy(cot).double())
    assert rel(gx.data.cpu().numpy(), gx_ref.detach().numpy()) < 2e-5
    assert rel(xv.grad.cpu().numpy(), d_x.numpy()) < 1e-4
    assert rel(gyv.grad.cpu().numpy(), d_gy.numpy()) < 2e-5
    assert rel(gv.grad.cpu().numpy(), d_g.numpy()) < 1e-4
    # ---------------- convolution: gx = d<gy, conv(x, W)>/dx ; VJPs w.r.t. gy (forward conv) and W (filter gradient)
    for cin, cout, k in ((64, 128, 3), (3, 64, 3), (64, 64, 1)):
        name = f"c{cin}_{cout}_{k}"
        np.random.seed(3)
        x0 = F.Var(torch.zeros(4, 16, 16, cin, device="cuda"))
        C.Conv2D(x0, cin, cout, k, 1, name, biases=False)               # creates the filter variable
        W = store.vars[name + "/Filters"]
        W.data.copy_(W.data.to(torch.bfloat16).float())                  # bf16-representable filter
        store.bump(W.root)
        W.grad = torch.zeros_like(W.data)
        gy2 = _bf16_repr(rs.standard_normal((4, 16, 16, cout)).astype("float32"))
        cot2 = _bf16_repr(rs.standard_normal((4, 16, 16, cin)).astype("float32"))
        gyv2 = F.Var(torch.from_numpy(gy2).cuda(), requires_grad=True)
        with store.gradient_tape() as tape:
            gx2 = F.conv2d_input_grad(gyv2, W, (4, 16, 16, cin), k, k)
            tape.backward(gx2, grad=torch.from_numpy(cot2).cuda())
        torch.cuda.synchronize()
        wt = W.data.cpu().double().requires_grad_(True)
        gt2 = torch.from_numpy(gy2).double().requires_grad_(True)
        xz = torch.zeros(4, 16, 16, cin, dtype=torch.double, requires_grad=True)
        from oracle import ops as O
        yy = O.conv2d_nhwc(xz, wt, 1, "SAME")
        gx_r, = torch.autograd.grad(yy, xz, gt2, create_graph=True)
        d_gy2, d_w = torch.autograd.grad(gx_r, (gt2, wt), torch.from_numpy(cot2).double())
        assert rel(gx2.data.cpu().numpy(), gx_r.detach().numpy()) < 1e-5, name
        assert rel(gyv2.grad.float().cpu().numpy(), d_gy2.numpy()) < 1e-5, name
        assert rel(W.grad.cpu().numpy(), d_w.numpy()) < 1e-4, name
