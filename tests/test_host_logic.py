"""CPU tier: the C-ABI library loads and exports everything include/ganb200.h declares (no compute calls), the
ctypes mirrors of its structs match the C layout, and the host-side mirror of the reference interface (variable
scopes, reuse, NumPy-RNG order, spectral-norm update_collection bookkeeping, optimistic restore) behaves like the
reference.  Layer functions run against a recording test double (tests/hostlogic.py): every kernel call is a no-op, so nothing
here computes -- arithmetic is only ever checked on the GPU tier against the oracle."""
import ctypes
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


from tests.hostlogic import RecordingLib  # noqa: E402


@pytest.fixture()
def host():
    from gan_lib_tensorflow_b200 import framework
    from tests import hostlogic

    rec = hostlogic.install(RecordingLib())
    store = framework.reset_default_graph("cpu", u_seed=2)
    yield store, rec
    framework.set_store(None)
    hostlogic.uninstall()


# ------------------------------------------------------------------------------------------------ C ABI
def test_library_loads_and_exports_every_declared_symbol():
    from gan_lib_tensorflow_b200 import cabi

    lib = cabi.lib()
    syms = cabi.header_symbols()
    assert len(syms) >= 35
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.ganb_abi_version() == 1
    assert isinstance(lib.ganb_last_error(), bytes)


def test_struct_mirrors_match_the_c_layout():
    from gan_lib_tensorflow_b200 import kernels as K

    src = '#include "ganb200.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(){printf("%zu %zu %zu %zu %zu %d\\n",' \
          'sizeof(ganb_sn_layer), offsetof(ganb_sn_layer, k), offsetof(ganb_sn_layer, blk_begin),' \
          'sizeof(ganb_pack_layer), offsetof(ganb_pack_layer, tile_begin), GANB_SN_ROWS);}'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        with open(c, "w") as fh:
            fh.write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    got = [ctypes.sizeof(K.SnLayerStruct), K.SnLayerStruct.k.offset, K.SnLayerStruct.blk_begin.offset,
           ctypes.sizeof(K.PackLayerStruct), K.PackLayerStruct.tile_begin.offset, K.SN_ROWS]
    assert [int(v) for v in out] == got


def test_no_cpu_fallback_when_library_or_gpu_is_missing(monkeypatch):
    from gan_lib_tensorflow_b200 import cabi, framework

    from gan_lib_tensorflow_b200 import kernels as K
    assert not K.host_logic_only()          # no test double installed: the product has no other way to run without a GPU
    monkeypatch.setenv("GANB_HOST_LOGIC_ONLY", "1")      # the round-1 switch is gone: the variable changes nothing
    monkeypatch.setattr(cabi, "_lib", None)
    monkeypatch.setattr(cabi, "LIB_PATH", "/nonexistent/libganb200.so")
    with pytest.raises(cabi.GanbError):
        cabi.lib()
    framework.set_store(None)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            framework.get_store()


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gan_lib_tensorflow_b200")):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(dirpath, f)


# ------------------------------------------------------------------------------------------------ variables
def test_variable_names_reuse_and_trainable_filter(host):
    store, _ = host
    from gan_lib_tensorflow_b200.common.ops import conv2d, linear
    from gan_lib_tensorflow_b200.framework import Var

    x = Var(torch.zeros(2, 8, 8, 16))
    with store.variable_scope("Discriminator"):
        conv2d.Conv2D(x, 16, 8, 3, 1, "D.Block.2.Conv1", spectral_normed=True, update_collection="NO_OPS")
        linear.Linear(Var(torch.zeros(2, 16)), 16, 4, "D.Output", spectral_normed=True, update_collection="NO_OPS")
    names = list(store.vars)
    assert names == ["Discriminator/D.Block.2.Conv1/Filters",
                     "Discriminator/D.Block.2.Conv1/filters/spectral_norm/u",      # conv2d.py:170 + sn.py:28,32
                     "Discriminator/D.Block.2.Conv1/Biases",
                     "Discriminator/D.Output/W", "Discriminator/D.Output/spectral_norm/u", "Discriminator/D.Output/b"]
    assert store.vars[names[1]].trainable is False and tuple(store.vars[names[1]].data.shape) == (1, 8)
    assert [v.name for v in store.trainable_variables("Discriminator")] == [n + ":0" for n in names if "/u" not in n]
    # reuse=True on a missing variable raises, on an existing one returns the same object
    with store.variable_scope("Discriminator", reuse=True):
        conv2d.Conv2D(x, 16, 8, 3, 1, "D.Block.2.Conv1", spectral_normed=True, update_collection="NO_OPS")
        with pytest.raises(ValueError):
            conv2d.Conv2D(x, 16, 8, 3, 1, "D.Block.9.Conv1")
    assert len(store.vars) == 6
    with pytest.raises(ValueError):
        conv2d.Conv2D(x, 32, 8, 3, 1, "bad")        # input_dim does not match the tensor


def test_pggan_variable_manifest_matches_the_oracle(host):
    """PGGAN/model_nvidia.py: scope-qualified variable names, shapes and NumPy-RNG initial values of G and D (incl. the
    fade-in toRGB / fromRGB pairs and the 513-channel D.Conv) agree with the oracle restatement."""
    store, _ = host
    from gan_lib_tensorflow_b200.PGGAN import model_nvidia as P
    from oracle import ops as O
    from oracle import pggan as OP
    from oracle import tfshim

    np.random.seed(0)
    pm = P.PGGAN(block_count=2, trans=True, inputs_norm=True)
    fake = pm.get_generator(torch.zeros(2, 512), 0.5)
    assert tuple(fake.shape) == (2, 16, 16, 3)
    logits = pm.get_discriminator(fake, 0.5, update_collection="NO_OPS")
    assert tuple(logits.shape) == (2,)
    np.random.seed(0)
    g = tfshim.Graph(dtype=torch.float32, u_seed=2)
    om = OP.PGGAN(2, True, True)
    with torch.no_grad():
        of = om.get_generator(g, torch.zeros(2, 512), 0.5)
        om.get_discriminator(g, of, 0.5, update_collection=O.NO_OPS)
    assert list(store.vars) == list(g.vars)
    for name, v in store.vars.items():
        assert tuple(v.data.shape) == tuple(g.vars[name].shape), name
        if "spectral_norm/u" not in name:
            np.testing.assert_array_equal(v.data.numpy(), g.vars[name].detach().numpy(), err_msg=name)
    assert "d_net/D.Conv/Filters" in store.vars and tuple(store.vars["d_net/D.Conv/Filters"].data.shape) == (3, 3, 513, 512)
    assert "g_net/G.2_toRGB1/Filters" in store.vars and "d_net/D.2_fromRGB2/filters/spectral_norm/u" in store.vars
    assert [v.key for v in store.trainable_variables("g_net")] == [n for n, _ in g.trainable_variables("g_net")]


def test_resnet_pggan_variable_manifest_matches_the_oracle(host):
    """PGGAN/model_resnet.py over common/resnet_block.py:192-349: names, shapes and initial values of the ResNet
    variant (feature-map fade-in blocks G.2_toRGB1 / G.2_toRGB2, RGB-input residual blocks D.2_fromRGB1 / 2) agree with
    the oracle restatement; the manifest carries the double scope of Normalize + BatchNorm."""
    store, _ = host
    from gan_lib_tensorflow_b200.PGGAN import model_resnet as P
    from oracle import ops as O
    from oracle import pggan as OP
    from oracle import tfshim

    np.random.seed(0)
    pm = P.PGGAN(block_count=2, trans=True, inputs_norm=True)
    fake = pm.get_generator(torch.zeros(2, 512), 0.5)
    assert tuple(fake.shape) == (2, 16, 16, 3)
    logits = pm.get_discriminator(fake, 0.5, update_collection="NO_OPS")
    assert tuple(logits.shape) == (2,)
    np.random.seed(0)
    g = tfshim.Graph(dtype=torch.float32, u_seed=2)
    om = OP.PGGANResNet(2, True, True)
    with torch.no_grad():
        of = om.get_generator(g, torch.zeros(2, 512), 0.5)
        assert tuple(of.shape) == (2, 16, 16, 3)
        om.get_discriminator(g, of, 0.5, update_collection=O.NO_OPS)
    # BatchNorm moving averages are write-only state in the reference and are not kept by the product
    # (the product creates N1's constant-initialised beta / gamma before the shortcut's filters: same NumPy stream)
    assert sorted(store.vars) == sorted(k for k in g.vars if "/moving_" not in k)
    norm = lambda ks: [k for k in ks if "/BatchNorm/" not in k]  # noqa: E731
    assert norm(store.vars) == norm(g.vars)
    for name, v in store.vars.items():
        assert tuple(v.data.shape) == tuple(g.vars[name].shape), name
        if "spectral_norm/u" not in name:
            np.testing.assert_array_equal(v.data.numpy(), g.vars[name].detach().numpy(), err_msg=name)
    for name in ("g_net/G.N0/BatchNorm/beta", "g_net/G.UpBlock.1.Shortcut/Filters", "g_net/G.2_toRGB2.Conv1/Filters",
                 "g_net/G.Output_Normalize/BatchNorm/gamma", "d_net/D.2_fromRGB1.Shortcut/filters/spectral_norm/u",
                 "d_net/D.DownBlock.1.Conv2/Filters", "d_net/D.NoneBlock.Conv1/Biases", "d_net/D.Output/W"):
        assert name in store.vars, name
    assert tuple(store.vars["g_net/G.Conv/Filters"].data.shape) == (3, 3, 1024, 1024)
    assert tuple(store.vars["d_net/D.2_fromRGB2.Conv1/Filters"].data.shape) == (3, 3, 3, 512)


def test_cross_gpu_batch_statistics_call_sequence(host):
    """store.bn_sync = (allreduce, world): a batch-statistics normalisation all-reduces [mean | E[x^2]] between
    ganb_bn_stats and the normalise kernel, and [sum(dy) | sum(dy*xhat)] between the two backward phases; instance
    statistics and un-normalised activations are never reduced."""
    store, rec = host
    from gan_lib_tensorflow_b200 import functional as F
    from gan_lib_tensorflow_b200.framework import Var

    reduced = []
    store.bn_sync = (lambda t: reduced.append((len(rec.calls), tuple(t.shape))), 4)
    x = Var(torch.zeros(4, 8, 8, 16, dtype=torch.bfloat16), requires_grad=True)
    with store.gradient_tape() as tape:
        y, _ = F.norm_act(x, stats="batch", act="relu")
        tape.backward(y, grad=torch.zeros(4, 8, 8, 16, dtype=torch.bfloat16))
    names = rec.names()
    assert names == ["ganb_bn_stats", "ganb_bn_moments_pack", "ganb_bn_moments_unpack", "ganb_norm_act_fwd",
                     "ganb_norm_act_bwd_phase", "ganb_norm_act_bwd_sums_offset", "ganb_norm_act_bwd_phase"]
    assert [r[0] for r in reduced] == [2, 6]                  # after pack, and between the two backward phases
    assert reduced[0][1] == (2 * 16,)      # [mean | E[x^2]] of 16 channels, one statistic group
    phases = [c[1][-3] for c in rec.calls if c[0] == "ganb_norm_act_bwd_phase"]
    scales = [c[1][-2].value for c in rec.calls if c[0] == "ganb_norm_act_bwd_phase"]
    assert phases == [1, 2] and scales == [1.0, 0.25]
    n_before = len(reduced)
    with store.gradient_tape() as tape:
        y, _ = F.norm_act(x, stats="instance", act=None)
        z, _ = F.norm_act(x, stats=None, act="relu")
    assert len(reduced) == n_before


def test_two_evaluations_of_one_network_on_a_tape_keep_separate_sn_state(host):
    """PGGAN / Pix2Pix call D on real and on fake images as separate layer calls inside one gradient computation
    (PGGAN/train.py:103-107): D(real) assigns u, D(fake) then sees the new u, so the two calls have different sigma.
    Each evaluation must keep its own sigma / v / u' / G buffers until the backward pass; outside a tape, and for a
    repeated evaluation with unchanged u, the state is shared (one grouped launch per weight version)."""
    store, rec = host
    from gan_lib_tensorflow_b200.common.ops import conv2d
    from gan_lib_tensorflow_b200.common.ops.sn import spectral_normed_weight
    from gan_lib_tensorflow_b200.framework import Var

    x = Var(torch.zeros(2, 8, 8, 16), requires_grad=True)

    def layer(mode, reuse):
        with store.variable_scope("d_net", reuse=reuse):
            return conv2d.Conv2D(x, 16, 16, 3, 1, "D.Conv", spectral_normed=True, update_collection=mode)

    layer("NO_OPS", False)                                    # graph construction, no tape
    w = store.vars["d_net/D.Conv/Filters"]
    u = store.vars["d_net/D.Conv/filters/spectral_norm/u"]
    base = store.sn_groups["d_net"].entries[w.key]
    n0 = rec.names().count("ganb_sn_power_iter")
    with store.gradient_tape() as tape:
        with store.variable_scope("d_net", reuse=True), store.variable_scope("D.Conv"), store.variable_scope("filters"):
            e_real = spectral_normed_weight(w, update_collection=None).entry           # assigns u
            e_fake = spectral_normed_weight(w, update_collection="NO_OPS").entry       # re-evaluated from the new u
            e_again = spectral_normed_weight(w, update_collection="NO_OPS").entry      # unchanged u: shared
    assert e_real is base and e_fake is not base and e_again is e_fake
    assert e_fake.g.data_ptr() != e_real.g.data_ptr() and e_fake.scal.data_ptr() != e_real.scal.data_ptr()
    assert rec.names().count("ganb_sn_power_iter") == n0 + 2
    assert e_fake.u is u and e_fake.group is store.sn_shadow["d_net"][0]
    with store.gradient_tape():                                 # a new tape starts from the shared state again
        with store.variable_scope("d_net", reuse=True), store.variable_scope("D.Conv"), store.variable_scope("filters"):
            assert spectral_normed_weight(w, update_collection="NO_OPS").entry is base


def test_two_player_trainers_call_sequence(host):
    """training.TwoPlayer through the PGGAN trainer (PGGAN/train.py:103-136, 182-190), kernels recorded on CPU: the
    critic step evaluates D twice (power iteration with assignment for D(real), a second state for D(fake)), leaves G's
    parameters without gradient work, ends in one Adam launch + operand repack of d_net; the generator step runs the
    power iteration once (NO_OPS) and updates g_net only."""
    store, rec = host
    from gan_lib_tensorflow_b200.PGGAN import train as PT

    tr = PT.Trainer(block_count=1, trans=True, inputs_norm=True, batch_size=2, seed=0)
    assert set(store.flat) == {"d_net", "g_net"}
    real, z = torch.zeros(2, 8, 8, 3), torch.zeros(2, 512)
    n0 = len(rec.calls)
    tr.d_step(real, z, 0.25)
    d_calls = rec.names()[n0:]
    assert d_calls.count("ganb_sn_power_iter") == 2 and d_calls.count("ganb_sn_bwd") == 2
    assert d_calls.count("ganb_adam") == 1 and d_calls[-1] in ("ganb_pack_weights", "ganb_pack_small")
    assert "ganb_minibatch_std_fwd" in d_calls and "ganb_minibatch_std_bwd" in d_calls
    d_params = {v.key for v in store.trainable_variables("d_net")}
    assert all(v.grad is not None for v in store.trainable_variables("d_net")) and len(d_params) > 10
    n1 = len(rec.calls)
    tr.g_step(z, 0.25)
    g_calls = rec.names()[n1:]
    assert g_calls.count("ganb_sn_power_iter") == 1 and g_calls.count("ganb_sn_bwd") == 0     # d_net is frozen
    assert g_calls.count("ganb_adam") == 1 and "ganb_pixel_norm_bwd" in g_calls
    assert tr.players.opt["d"].t == 1 and tr.players.opt["g"].t == 1
    assert abs(tr.alpha(50000) - 0.5) < 1e-12


def test_resnet_pggan_trainer_call_sequence(host):
    """PGGAN/train.py with --model resnet (train.py:61-66): the same two training ops over model_resnet.PGGAN.  The
    critic's skip path resizes the image with ganb_subsample2d, whose gradient (scatter = 1) only runs in the generator
    step (the critic step needs no image gradient); the fade-in blends feature maps with ganb_lerp_*."""
    store, rec = host
    from gan_lib_tensorflow_b200.PGGAN import train as PT

    tr = PT.Trainer(block_count=1, trans=True, inputs_norm=True, batch_size=2, seed=0, model="resnet")
    real, z = torch.zeros(2, 8, 8, 3), torch.zeros(2, 512)
    n0 = len(rec.calls)
    tr.d_step(real, z, 0.25)
    d_calls = rec.names()[n0:]
    assert d_calls.count("ganb_sn_power_iter") == 2 and d_calls.count("ganb_sn_bwd") == 2
    assert d_calls.count("ganb_subsample2d") == 2          # D(real) and D(fake), forward only
    assert d_calls.count("ganb_lerp_fwd") == 3 and d_calls.count("ganb_lerp_bwd") == 2   # G (no tape), D twice
    assert d_calls.count("ganb_adam") == 1
    n1 = len(rec.calls)
    tr.g_step(z, 0.25)
    g_calls = rec.names()[n1:]
    assert g_calls.count("ganb_subsample2d") == 2          # forward + the scatter of the image gradient
    assert g_calls.count("ganb_lerp_fwd") == 2 and g_calls.count("ganb_lerp_bwd") == 2
    assert g_calls.count("ganb_sn_power_iter") == 1 and g_calls.count("ganb_sn_bwd") == 0
    with pytest.raises(NotImplementedError):
        PT.Trainer(block_count=1, trans=False, model="vgg")


def test_pix2pix_model_net_type_dispatch(host):
    """Pix2Pix/model.py:15-100: 'UNet' -> the nine-level unet_generator / unet_discriminator, 'UNet_Attention' -> unet_g /
    unet_d; conv_type / channel_multiplier reach every Conv2D; the unrunnable families raise."""
    store, rec = host
    from gan_lib_tensorflow_b200.Pix2Pix.model import Pix2Pix
    from gan_lib_tensorflow_b200.Pix2Pix.train import Trainer

    m = Pix2Pix()
    x = torch.zeros(1, 512, 512, 3)
    out = m.get_generator(x, 3, ngf=8, net_type='UNet')
    assert tuple(out.shape) == (1, 512, 512, 3) and "g_net/encoder_9/Conv2D/Filters" in store.vars
    d = m.get_discriminator(x, out, ndf=8, update_collection="NO_OPS", net_type='UNet')
    assert tuple(d.shape) == (1, 30, 30, 1) and "d_net/layer_6/Conv2D/filters/spectral_norm/u" in store.vars
    for net_type in ('ResNet', 'VGG', 'nope'):
        with pytest.raises(NotImplementedError):
            m.get_generator(x, 3, net_type=net_type, reuse=True)
        with pytest.raises(NotImplementedError):
            m.get_discriminator(x, x, net_type=net_type, reuse=True)
    from gan_lib_tensorflow_b200 import framework
    store2 = framework.reset_default_graph("cpu", u_seed=2)
    tr = Trainer(ngf=8, ndf=8, size=256, conv_type='separable_conv2d', channel_multiplier=1)
    assert "g_net/encoder_9/Conv2D/Filters" not in store2.vars          # default net_type: unet_g / unet_d
    assert "g_net/encoder_8/Conv2D/pointwise_filters" in store2.vars and "d_net/layer_5/Conv2D/depthwise_filters" in store2.vars
    n0 = len(rec.calls)
    tr.d_step(torch.zeros(1, 256, 256, 3), torch.zeros(1, 256, 256, 3))
    calls = rec.names()[n0:]
    assert calls.count("ganb_depthwise_conv2d_fwd") == 16 + 2 * 5 and calls.count("ganb_depthwise_conv2d_bwd_filter") == 2 * 5


def test_acgan_narrow_generator_variant(host):
    """ACGAN/model_.py: 4x4x128 seed, 128-channel blocks, same names."""
    store, _ = host
    from gan_lib_tensorflow_b200.ACGAN.model_ import ACGAN

    out = ACGAN().get_generator(torch.zeros(2, 128), labels=torch.zeros(2, dtype=torch.int32))
    assert tuple(out.shape) == (2, 32, 32, 3)
    assert tuple(store.vars["g_net/G.Input/W"].data.shape) == (128, 2048)
    assert tuple(store.vars["g_net/G.3.Conv2/Filters"].data.shape) == (3, 3, 128, 128)
    assert "g_net/G.1.Shortcut/Filters" in store.vars      # 'up' blocks always have a conv shortcut


def test_pix2pix_gradient_penalty_call_sequence(host):
    """Pix2Pix --loss_type WGAN-GP (train.py:489-503, Pix2Pix/gp.py): the critic step evaluates D three times with
    update_collection=None (three power iterations, three spectral-norm backward passes over three state generations),
    interpolates once, evaluates the penalty kernel once, and every one of the 5 convolutions emits 3 filter gradients
    (real / fake passes + the penalty's wgrad(c_l, gy_l))."""
    store, rec = host
    from gan_lib_tensorflow_b200.Pix2Pix import train as PT

    tr = PT.Trainer(ngf=8, ndf=8, size=256, loss_type='WGAN-GP')
    x, t = torch.zeros(2, 256, 256, 3), torch.zeros(2, 256, 256, 3)
    n0 = len(rec.calls)
    token = store.tape_token
    loss = tr.players.gradients("d", lambda: tr.d_loss(x, t, None, gp_alpha=torch.tensor([0.3, 0.6])))
    names = rec.names()[n0:]
    assert tuple(loss.shape) == (1,)
    assert names.count("ganb_sn_power_iter") == 3 and names.count("ganb_sn_bwd") == 3
    assert names.count("ganb_interpolate") == 1 and names.count("ganb_gp_loss") == 1
    assert names.count("ganb_conv2d_wgrad") == 15
    # the inner tapes leave the outer bookkeeping alone: every tape has its own token, the outer one is current again
    # after each inner tape and nothing is recording once the step is over
    assert store.tape_token == token and store.tape is None and store._token_seq == token + 3
    with pytest.raises(NotImplementedError):
        PT.Trainer(ngf=8, ndf=8, size=256, loss_type='WGAN-GP', conv_type='separable_conv2d', channel_multiplier=1)


def test_legacy_conv2d_signature_and_pixelnorm_alias(host):
    store, _ = host
    from gan_lib_tensorflow_b200.common import resnet_block
    from gan_lib_tensorflow_b200.common.ops import conv2d_, pixelnorm
    from gan_lib_tensorflow_b200.framework import Var

    conv2d_.Conv2D(Var(torch.zeros(1, 4, 4, 8)), 8, 8, 3, 1, "L", spectral_normed=True, update_collection="NO_OPS",
                   reuse=False)
    assert "L/spectral_norm/u" in store.vars            # conv2d_.py: no 'filters/' level
    assert callable(pixelnorm.Pixelnorm)                # model_nvidia.py:63 resolves
    assert resnet_block.get_dim(1) == 512 and isinstance(resnet_block.get_dim(3), int) and resnet_block.get_dim(3) == 256


def test_numpy_rng_stream_and_initial_values_match_the_oracle(host):
    """Variables are created in the reference's graph-construction order, including the draws that every reuse
    call makes and discards (conv2d.py:124-144); names, shapes and values must equal the oracle's."""
    store, _ = host
    from gan_lib_tensorflow_b200.SNGAN import gan_cifar_resnet as P
    from oracle import sngan_cifar as O

    P.Trainer(batch_size=64, seed=0, store=store)
    state_after_product = np.random.get_state()[1].copy()
    np.random.seed(0)
    om = O.SNGANCifar(dtype=torch.float32, u_seed=2)
    om.build()
    assert np.array_equal(np.random.get_state()[1], state_after_product)       # same number of draws consumed
    assert set(om.g.vars) == set(store.vars)
    for name, ov in om.g.vars.items():
        pv = store.vars[name]
        assert tuple(pv.data.shape) == tuple(ov.shape), name
        assert torch.equal(pv.data, ov.detach()), name
        assert pv.trainable == om.g.trainable[name], name
    flat = store.flat["Generator"]
    assert sum(v.data.numel() for v in flat.variables) == 7875587
    assert sum(v.data.numel() for v in store.flat["Discriminator"].variables) == 1701689
    for v in flat.variables:                                                    # views into the flat buffers
        off = v.data.data_ptr() - flat.params.data_ptr()
        assert off >= 0 and off % 256 == 0                                     # 16-byte rule of the TMA path
        assert v.grad.data_ptr() - flat.grads.data_ptr() == off


def test_spectral_norm_update_collection_bookkeeping(host):
    """update_collection=None: ONE grouped power iteration per forward pass of the network, with assign;
    NO_OPS: re-evaluated only when weights or u changed; a named collection defers the assignment."""
    store, rec = host
    from gan_lib_tensorflow_b200.common.ops import conv2d, sn
    from gan_lib_tensorflow_b200.framework import Var

    x = Var(torch.zeros(1, 4, 4, 8))

    def net(mode):
        with store.variable_scope("D", reuse=len(store.vars) > 0):
            conv2d.Conv2D(x, 8, 8, 3, 1, "a", spectral_normed=True, update_collection=mode)
            conv2d.Conv2D(x, 8, 8, 3, 1, "b", spectral_normed=True, update_collection=mode)

    net("NO_OPS")                       # build: two registrations
    rec.calls.clear()
    net("NO_OPS")
    assert rec.names().count("ganb_sn_power_iter") == 0          # cached: nothing changed
    net(None)
    calls = [c for c in rec.calls if c[0] == "ganb_sn_power_iter"]
    assert len(calls) == 1 and calls[0][1][4] == 1                # one grouped launch for both layers, assign=1
    assert calls[0][1][1] == 2
    net(None)
    assert rec.names().count("ganb_sn_power_iter") == 2           # a new forward pass iterates again
    net("NO_OPS")
    calls = [c for c in rec.calls if c[0] == "ganb_sn_power_iter"]
    assert len(calls) == 3 and calls[-1][1][4] == 0               # u changed -> fresh evaluation, no assign
    store.bump("D")
    net("NO_OPS")
    assert rec.names().count("ganb_sn_power_iter") == 4           # weights changed
    net("my_update_ops")
    assert len(sn.get_collection("my_update_ops")) == 2
    sn.run_update_collection("my_update_ops")
    assert sn.get_collection("my_update_ops") == []
    with pytest.raises(NotImplementedError):
        sn.spectral_normed_weight(store.vars["D/a/Filters"], num_iters=2)


def test_training_step_call_sequence_and_freezing(host):
    """D-step: generator runs without a tape (var_list=disc_params) -> no wgrad for G; G-step: D is frozen."""
    store, rec = host
    from gan_lib_tensorflow_b200.SNGAN import gan_cifar_resnet as P

    tr = P.Trainer(batch_size=64, seed=0, store=store)
    rec.calls.clear()
    tr.d_step(0)
    d_names = rec.names()
    n_wgrad_d = d_names.count("ganb_conv2d_wgrad")
    assert n_wgrad_d == 10                                        # the 10 D convolutions (SN weights)
    assert d_names.count("ganb_sn_bwd") == 1 and d_names.count("ganb_adam") == 1
    assert d_names.count("ganb_bn_stats") == 7                    # G forward only
    assert "ganb_norm_act_bwd" in d_names                         # D's own activations
    rec.calls.clear()
    tr.g_step(1)
    g_names = rec.names()
    assert g_names.count("ganb_sn_bwd") == 0                      # D frozen: no SN backward, no D wgrad
    assert g_names.count("ganb_conv2d_wgrad") == 11               # 10 G convolutions + G.Input
    assert g_names.count("ganb_adam") == 1
    stats = [c for c in rec.calls if c[0] == "ganb_bn_stats"]
    assert len(stats) == 7 and all(c[1][5] == 2 for c in stats)   # two statistic towers of 64 (reference towers)
    assert tr.disc_opt.t == 1 and tr.gen_opt.t == 1
    assert P.lr_decay(0) == 1.0 and P.lr_decay(60000) == 0.5


def test_critic_blocks_fuse_their_mid_block_relu_into_the_two_convolutions(host):
    """No normalisation sits between Conv1 and Conv2 of a critic block: Conv1 is launched with act = relu, no activation
    pass follows it, and Conv2's data gradient is the gated entry; with the switch off the separate passes come back."""
    store, rec = host
    from gan_lib_tensorflow_b200 import cabi, functional as F
    from gan_lib_tensorflow_b200.SNGAN import gan_cifar_resnet as P

    tr = P.Trainer(batch_size=64, seed=0, store=store)
    rec.calls.clear()
    tr.d_step(0)
    names = rec.names()
    relu = cabi.act_code("relu")
    igemm = [c for c in rec.calls if c[0] == "ganb_conv2d_igemm"]
    fused_fwd = [c for c in igemm if c[1][20] == relu]           # (..., residual_up2, act, out_dtype, stream)
    assert len(fused_fwd) == 4                                    # Conv1 of D.Block.1 .. D.Block.4
    assert all(c[1][21] == cabi.BF16 for c in fused_fwd)          # act(y) is stored in bf16 only
    gated = [c for c in rec.calls if c[0] == "ganb_conv2d_igemm_gated"]
    assert len(gated) == 4 and all(c[1][18] == relu for c in gated)
    n_act_fwd, n_act_bwd = names.count("ganb_norm_act_fwd"), names.count("ganb_norm_act_bwd")
    old = F.FUSED_CONV_ACT
    F.FUSED_CONV_ACT = False
    try:
        rec.calls.clear()
        tr.d_step(1)
        names = rec.names()
        assert "ganb_conv2d_igemm_gated" not in names
        assert not [c for c in rec.calls if c[0] == "ganb_conv2d_igemm" and c[1][20] == relu]
        assert names.count("ganb_norm_act_fwd") == n_act_fwd + 4 and names.count("ganb_norm_act_bwd") == n_act_bwd + 4
    finally:
        F.FUSED_CONV_ACT = old


def test_pair_schedule_issues_the_same_calls_as_the_two_steps(host):
    """Trainer.pair_step = d_step + g_step with the generator step's G forward issued first (next to the critic step on
    the GPU): the same multiset of kernel calls, one Adam per optimiser, G's statistics in two towers of 64."""
    store, rec = host
    from collections import Counter

    from gan_lib_tensorflow_b200.SNGAN import gan_cifar_resnet as P

    tr = P.Trainer(batch_size=64, seed=0, store=store)
    tr.d_step(1)
    tr.g_step(1)
    rec.calls.clear()
    tr.d_step(2)
    tr.g_step(2)
    two_steps = Counter(rec.names())
    rec.calls.clear()
    tr.pair_step(3)
    pair = rec.names()
    assert Counter(pair) == two_steps
    # the generator forward of the G-step (two towers of 64: groups == 2 at n == 128) comes before the critic's wgrads
    first_wgrad = pair.index("ganb_conv2d_wgrad")
    g_stats = [i for i, c in enumerate(rec.calls) if c[0] == "ganb_bn_stats" and c[1][2] == 128]
    assert len(g_stats) == 7 and max(g_stats) < first_wgrad
    assert tr.disc_opt.t == 3 and tr.gen_opt.t == 3 and store.tape is None


def test_optimistic_restore_matches_name_and_shape_only(host):
    store, _ = host
    from gan_lib_tensorflow_b200.common.ops import linear
    from gan_lib_tensorflow_b200.framework import Var

    with store.variable_scope("Generator"):
        linear.Linear(Var(torch.zeros(2, 16)), 16, 8, "G.Input")
    state = {"Generator/G.Input/W": np.ones((16, 8), "float32"), "Generator/G.Input/b": np.ones((9,), "float32"),
             "Generator/Unknown": np.ones((3,), "float32")}
    restored = store.load_state_dict(state)
    assert restored == ["Generator/G.Input/W"]                    # common/misc.py:275-307 semantics
    assert float(store.vars["Generator/G.Input/W"].data.sum()) == 128.0


def test_same_padding_helper():
    from gan_lib_tensorflow_b200 import kernels as K

    assert K.same_pads(32, 3, 1) == (1, 1, 32)
    assert K.same_pads(256, 4, 1) == (1, 2, 256)
    assert K.same_pads(256, 4, 2) == (1, 1, 128)
    assert K.same_pads(8, 3, 2) == (0, 1, 4)


# ------------------------------------------------------------------------------------------------ 2 ranks, gloo
_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
from tests import hostlogic
hostlogic.install()
import torch, torch.distributed as dist
from gan_lib_tensorflow_b200 import framework
from gan_lib_tensorflow_b200.SNGAN import gan_cifar_resnet as P
rank = int(os.environ["RANK"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=2)
store = framework.reset_default_graph("cpu")
seen = []
def allreduce(g):
    seen.append(float(g[0]))
    dist.all_reduce(g)
tr = P.Trainer(batch_size=64, seed=0, store=store, world_size=2, grad_allreduce=allreduce)
orig = tr._d_compute
def fake_compute():
    orig()
    store.flat["Discriminator"].grads.fill_(rank + 1.0)     # stand-in for the rank's local gradient
tr._d_compute = fake_compute
tr.d_step(0)
g = store.flat["Discriminator"].grads
assert torch.all(g == 3.0), g[:4]                           # 1 + 2 summed over the two ranks
assert seen == [rank + 1.0]
# replicated weights and u need no exchange: identical initial values on both ranks
w = store.vars["Discriminator/D.Output/W"].data.clone()
ws = [torch.zeros_like(w) for _ in range(2)]
dist.all_gather(ws, w)
assert torch.equal(ws[0], ws[1])
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_gradient_allreduce_gloo():
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    code = _WORKER.format(root=ROOT, port=port)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out.decode()[-2000:]


_WORKER_LOOP = r'''
import os, sys, tempfile
sys.path.insert(0, {root!r})
from tests import hostlogic
hostlogic.install()
import numpy as np, torch, torch.distributed as dist
from gan_lib_tensorflow_b200 import framework
from gan_lib_tensorflow_b200.ACGAN import train as AT
rank = int(os.environ["RANK"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=2)
store = framework.reset_default_graph("cpu")
calls = []
def allreduce(g):
    calls.append(g.numel())
    g.fill_(rank + 1.0)                      # stand-in for the rank's local gradient
    dist.all_reduce(g)
    assert torch.all(g == 3.0)
tr = AT.Trainer(batch_size=4, gradient_penalty=False, seed=0, world_size=2, grad_allreduce=allreduce)
rs = np.random.RandomState(10 + rank)        # every rank reads its own shard
data = rs.randint(0, 256, size=(3, 4, 3072)).astype("int32")
labels = rs.randint(0, 10, size=(3, 4)).astype("int32")
out = tempfile.mkdtemp()
tr.train(2, lambda: ((data[i], labels[i]) for i in range(3)), None, n_dis=2, out_dir=out, display_interval=10,
         out_image_interval=10, capture_after=None, log=lambda *a: None)
# 4 critic steps + 1 generator step, one collective each, over the whole flat gradient buffer of that network
sizes = [store.flat["d_net"].grads.numel()] * 2 + [store.flat["g_net"].grads.numel()] + [store.flat["d_net"].grads.numel()] * 2
assert calls == sizes, (calls, sizes)
assert tr.players.opt["d"].t == 4 and tr.players.opt["g"].t == 1 and tr.players.world_size == 2
w = store.vars["d_net/D.Output/W"].data.clone()
ws = [torch.zeros_like(w) for _ in range(2)]
dist.all_gather(ws, w)
assert torch.equal(ws[0], ws[1])             # replicated parameters stay identical (same initial values, summed gradients)
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_training_loop_gloo():
    """ACGAN Trainer.train (training.reference_loop over TwoPlayer) on two gloo ranks in host-logic mode: one gradient
    collective per optimiser step over that network's flat buffer, in step order (D, D, G, D, D), replicated
    parameters identical afterwards."""
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    code = _WORKER_LOOP.format(root=ROOT, port=port)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out.decode()[-2000:]
