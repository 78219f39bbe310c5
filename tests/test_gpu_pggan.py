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
