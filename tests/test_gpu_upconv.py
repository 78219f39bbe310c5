"""GPU parity tests for the sub-pixel form of UpsampleConv (common/resnet_block.py:83-97; ganb_upconv_* in
include/ganb200.h): four 2x2 convolutions over the low-resolution tensor instead of a 3x3 convolution over the nearest-2x
upsampled one.  The fp32 oracle is the reference formula (upsample, then convolve); the bf16-operand oracle mirrors the
product's rounding points (effective filters rounded, oracle.ops.subpixel_upconv_rounded)."""
import numpy as np
import pytest
import torch

from tests.test_gpu_ops import TOL_BLOCK_FP32, TOL_BLOCK_IMPL, _bf16_repr, check, env, rel, run_pair  # noqa: F401

pytestmark = pytest.mark.gpu


def _from_quad(t, n, h, w, c):
    """quad layout [n, h, w, 2i+j, c] -> NHWC [n, 2h, 2w, c]"""
    return t.reshape(n, h, w, 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(n, 2 * h, 2 * w, c)


def _to_quad(t, n, h, w, c):
    return t.reshape(n, h, 2, w, 2, c).permute(0, 1, 3, 2, 4, 5).contiguous().reshape(n, 2 * h, 2 * w, c)


@pytest.mark.parametrize("n,h,w,cin,cout", [(10, 16, 16, 128, 128), (6, 16, 8, 256, 256), (5, 32, 16, 192, 128),
                                            (3, 16, 16, 128, 320)])
def test_upconv_op_matches_oracle(env, n, h, w, cin, cout):
    store, tfshim = env
    from gan_lib_tensorflow_b200 import functional as F
    from gan_lib_tensorflow_b200 import kernels as K
    from gan_lib_tensorflow_b200.common.ops import conv2d as P
    from oracle import ops as O
    from oracle import resnet_block as ORB

    assert K.upconv_supported(n, h, w, cin, cout)
    rs = np.random.RandomState(31)
    x = _bf16_repr(rs.standard_normal((n, h, w, cin)).astype("float32"))
    cot = rs.standard_normal((n, 2 * h, 2 * w, cout)).astype("float32")
    # ---- product: low-resolution input, quad-layout output and cotangent
    np.random.seed(0)
    xv = F.Var(torch.from_numpy(x).cuda().to(torch.bfloat16), requires_grad=True)
    xv.grad_dtype = torch.float32
    with store.gradient_tape() as tape:
        out = P.Conv2D(xv, cin, cout, 3, 1, "G.Up.Conv1", he_init=True, subpixel_up2=True, out_dtype=torch.float32)
        assert out.quad
        for v in store.vars.values():
            v.grad = torch.zeros_like(v.data)
        cq = _to_quad(torch.from_numpy(cot), n, h, w, cout).cuda()
        tape.backward(out, grad=cq)
    torch.cuda.synchronize()
    got = {"out": _from_quad(out.data.float().cpu(), n, h, w, cout).numpy(), "dx": xv.grad.float().cpu().numpy(),
           "params": {k: v.grad.cpu().numpy() for k, v in store.vars.items()}}
    # ---- oracles
    refs = {}
    ORB.SUBPIXEL_RULE = lambda *a: True
    try:
        for mode in (True, False):
            O.BF16_OPERANDS = mode
            np.random.seed(0)
            g = tfshim.Graph(dtype=torch.float32, u_seed=2)
            xt = torch.from_numpy(x).clone().requires_grad_(True)
            yo = ORB.UpsampleConv(g, xt, cout, 3, name="G.Up.Conv1", he_init=True)
            params = g.trainable_variables()
            grads = torch.autograd.grad(yo, [xt] + [p for _, p in params], torch.from_numpy(cot))
            refs["bf16" if mode else "fp32"] = {"out": yo.detach().numpy(), "dx": grads[0].numpy(),
                                                "params": {nm: gr.numpy() for (nm, _), gr in zip(params, grads[1:])}}
    finally:
        O.BF16_OPERANDS = False
        ORB.SUBPIXEL_RULE = None
    check(got, refs, tag=f"upconv {n}x{h}x{w} {cin}->{cout}")


def test_up_block_subpixel_matches_oracle(env):
    """G.Block.3 of SNGAN-CIFAR at the benchmark shape (64 x 16x16 x 256 -> 32x32 x 256, conditional batch norm):
    ResidualBlock picks the sub-pixel path by itself (functional.upconv_eligible)."""
    store, tfshim = env
    from gan_lib_tensorflow_b200 import functional as F
    from gan_lib_tensorflow_b200.common import resnet_block as P
    from oracle import resnet_block as ORB

    n, h, c = 64, 16, 256
    assert F.upconv_eligible(n, h, h, c, c, 3)
    labels = np.random.RandomState(20).randint(0, 10, size=n).astype("int32")
    lab_p, lab_t = torch.from_numpy(labels).cuda(), torch.from_numpy(labels).long()
    x = _bf16_repr(np.random.RandomState(21).standard_normal((n, h, h, c)).astype("float32"))
    ORB.SUBPIXEL_RULE = lambda n_, h_, w_, ci, co, k: k == 3
    try:
        prod, refs = run_pair(
            store, tfshim,
            lambda xv: P.ResidualBlock(xv, c, c, 3, "G.Block.X", resample="up", labels=lab_p),
            lambda g, xt: ORB.ResidualBlock(g, xt, c, c, 3, "G.Block.X", resample="up", labels=lab_t), x)
    finally:
        ORB.SUBPIXEL_RULE = None
    check(prod, refs, tol_impl=TOL_BLOCK_IMPL, tol_fp32=TOL_BLOCK_FP32, tag="up block sub-pixel")
