#!/usr/bin/env python
"""Freezes the CPU oracle's outputs at the BASELINE.json shapes of configs 3-5 (tests/fullsize_cases.py) as compact
summaries tests/golden/fullsize_<case>.npz (norm + 4096-element sample of every output / gradient tensor, for the fp32
oracle and for the bf16-operand oracle).  The oracle needs minutes per case on 8 cores, which is why the GPU tests read
these files instead of re-running it.

  python tests/golden/make_fullsize.py [case ...]        # default: all cases
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ops as O  # noqa: E402
from oracle import tfshim  # noqa: E402
from tests import fullsize_cases as FC  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def _layer_case(orc_fn, x_np, cot_np):
    """The oracle half of tests.test_gpu_ops.run_pair for both operand modes."""
    res = {}
    for mode in ("bf16", "fp32"):
        O.BF16_OPERANDS = mode == "bf16"
        try:
            np.random.seed(0)
            g = tfshim.Graph(dtype=torch.float32, u_seed=2)
            xt = torch.from_numpy(x_np).clone().requires_grad_(True)
            yo = orc_fn(g, xt)
            params = g.trainable_variables()
            grads = torch.autograd.grad(yo, [xt] + [p for _, p in params], torch.from_numpy(cot_np), allow_unused=True)
            out = {}
            FC.flatten(mode + "/out", FC.summarize("out", yo.detach().numpy()), out)
            FC.flatten(mode + "/dx", FC.summarize("dx", grads[0].numpy()), out)
            for (n, _), gr in zip(params, grads[1:]):
                if gr is not None:
                    FC.flatten(mode + "/param/" + n, FC.summarize(n, gr.numpy()), out)
            res.update(out)
        finally:
            O.BF16_OPERANDS = False
    return res


def pggan_g():
    from oracle import pggan as OP
    i = FC.pggan_inputs()
    om = OP.PGGAN(i["bc"], i["trans"], True)
    return _layer_case(lambda g, zt: om.get_generator(g, zt, i["alpha"]), i["z"], i["cot_g"])


def pggan_d():
    from oracle import pggan as OP
    i = FC.pggan_inputs()
    om = OP.PGGAN(i["bc"], i["trans"], False)
    return _layer_case(lambda g, xt: om.get_discriminator(g, xt, i["alpha"], spectral_normed=True,
                                                          update_collection=O.NO_OPS), i["x"], i["cot_d"])


def pix2pix_g():
    from oracle import pix2pix as OX
    i = FC.pix2pix_inputs()
    return _layer_case(lambda g, xt: OX.unet_g(g, xt, 3, i["ngf"], keep_masks=[torch.from_numpy(m) for m in i["masks"]]),
                       i["x"], i["cot_g"])


def pix2pix_d():
    from oracle import pix2pix as OX
    i = FC.pix2pix_inputs()
    tgt = torch.from_numpy(i["tgt"])
    rs = np.random.RandomState(604)
    np.random.seed(0)
    g0 = tfshim.Graph(dtype=torch.float32, u_seed=2)
    with torch.no_grad():
        shape = tuple(OX.unet_d(g0, torch.from_numpy(i["x"]), tgt, i["ndf"], True, O.NO_OPS).shape)
    cot = rs.standard_normal(shape).astype("float32")
    res = _layer_case(lambda g, xt: OX.unet_d(g, xt, tgt, i["ndf"], True, O.NO_OPS), i["x"], cot)
    res["cot"] = cot
    return res


def imagenet_step():
    from oracle import sngan_imagenet as OI
    i = FC.imagenet_inputs()
    b, h = i["batch"], i["batch"] // 2
    res = {}
    for mode in ("bf16", "fp32"):
        O.BF16_OPERANDS = mode == "bf16"
        try:
            np.random.seed(0)
            om = OI.SNGANImageNet(dtype=torch.float32, u_seed=2)
            om.build()
            lab = torch.from_numpy(i["labels"]).long()
            dc, dp, dg = om.disc_grads(torch.from_numpy(i["data"]), lab,
                                       [torch.from_numpy(i["z_d"][:h]), torch.from_numpy(i["z_d"][h:])],
                                       torch.from_numpy(i["deq"]), None)
            u = {n: v.detach().numpy().copy() for n, v in om.g.vars.items() if n.endswith("/u")}
            gc, gp, gg = om.gen_grads([torch.from_numpy(i["z_g"][:b]), torch.from_numpy(i["z_g"][b:])],
                                      [torch.from_numpy(i["fl"][:b]).long(), torch.from_numpy(i["fl"][b:]).long()])
            res[mode + "/d_cost"] = np.asarray(dc.item())
            res[mode + "/g_cost"] = np.asarray(gc.item())
            for (n, _), t in list(zip(dp, dg)) + list(zip(gp, gg)):
                if t is not None:
                    FC.flatten(mode + "/param/" + n, FC.summarize(n, t.numpy()), res)
            for n, v in u.items():
                FC.flatten(mode + "/u/" + n, FC.summarize(n, v), res)
        finally:
            O.BF16_OPERANDS = False
    return res


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    cases = sys.argv[1:] or list(FC.CASES)
    for c in cases:
        t0 = time.time()
        res = globals()[c]()
        path = os.path.join(OUT, "fullsize_%s.npz" % c)
        np.savez_compressed(path, **res)
        print("%s: %d arrays, %.1f s, %.1f KB" % (c, len(res), time.time() - t0, os.path.getsize(path) / 1e3), flush=True)


if __name__ == "__main__":
    main()
