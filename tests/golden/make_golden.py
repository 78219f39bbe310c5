"""Generates tests/golden/oracle_golden.json: frozen outputs of the CPU oracle on fixed seeded inputs.

The reference itself cannot be executed (TensorFlow 1.5 is not installable in this image and no golden
vectors ship with it), so these vectors are produced by the ORACLE, not by the reference: they do not pin parity
with the reference (that stays "unpinned", oracle/__init__.py) but they freeze the oracle so that any later edit
that changes its arithmetic is caught, and they travel to the GPU box where /root/reference does not exist.

    python -m tests.golden.make_golden        # rewrites the JSON
"""
import json
import os

import numpy as np
import torch

from oracle import ops as O
from oracle import resnet_block as RB
from oracle import sngan_cifar as S
from oracle import tfshim

HERE = os.path.dirname(os.path.abspath(__file__))


def _sig(t, k=6):
    """A short, order-sensitive signature of a tensor: a few entries plus weighted sums."""
    a = t.detach().double().reshape(-1).numpy()
    w = np.cos(np.arange(a.size) * 0.37)
    return [float(a[0]), float(a[a.size // 2]), float(a[-1]), float(a.sum()), float((a * w).sum()),
            float(np.abs(a).max())][:k]


def compute():
    out = {}
    O.BF16_OPERANDS = False
    rs = np.random.RandomState(0)
    x = torch.from_numpy(rs.standard_normal((2, 8, 8, 16)).astype("float32"))
    labels = torch.tensor([3, 7])

    np.random.seed(0)
    g = tfshim.Graph(dtype=torch.float32, u_seed=2)
    y = O.Conv2D(g, x, 16, 8, 3, 1, "D.c", spectral_normed=True, update_collection=None)
    out["conv_sn_same"] = _sig(y)
    out["conv_sn_u_after_assign"] = _sig(g.vars["D.c/filters/spectral_norm/u"])
    y = O.Conv2D(g, x, 16, 8, 4, 1, "c4", he_init=False)
    out["conv_4x4_same_asymmetric_pad"] = _sig(y)
    y = O.Linear(g, x.reshape(2, -1), 1024, 5, "lin", spectral_normed=True, update_collection=O.NO_OPS)
    out["linear_sn"] = _sig(y)
    y = O.cond_batchnorm(g, "G.n", [0, 1, 2], x, labels=labels, n_labels=10)
    out["cond_batchnorm"] = _sig(y)
    out["pixel_norm"] = _sig(O.pixel_norm(x))
    y = O.embed_y(g, labels, 10, 12)
    out["embed_y"] = _sig(y)

    np.random.seed(1)
    g = tfshim.Graph(dtype=torch.float32, u_seed=2)
    for resample in ("up", "down", None):
        name = "G.B.%s" % resample
        y = RB.ResidualBlock(g, x, 16, 8, 3, name, resample=resample, labels=labels)
        out["resblock_%s" % resample] = _sig(y)
    y = RB.OptimizedResBlockDisc1(g, x[..., :3].contiguous(), 8, spectral_normed=True, update_collection=O.NO_OPS)
    out["optimized_first_block"] = _sig(y)

    # one small critic step + generator step of SNGAN-CIFAR (batch 4), losses and two gradient signatures
    np.random.seed(0)
    S.BATCH_SIZE = 4
    try:
        m = S.SNGANCifar(dtype=torch.float32, u_seed=2)
        m.build()
        data, lab = S.synthetic_batch(seed=0, batch=4)
        r2 = np.random.RandomState(1)
        z = [torch.from_numpy(r2.standard_normal((2, 128)).astype("float32")) for _ in range(2)]
        deq = torch.from_numpy(r2.uniform(0, 1 / 128, size=(4, 3072)).astype("float32"))
        cost, params, grads = m.disc_grads(torch.from_numpy(data), torch.from_numpy(lab).long(), z, deq,
                                           update_collection=None)
        out["sngan_d_cost_b4"] = [float(cost)]
        gd = dict((n, gr) for (n, _), gr in zip(params, grads))
        out["sngan_dgrad_D.Output.W"] = _sig(gd["Discriminator/D.Output/W"])
        out["sngan_dgrad_D.Block.1.Conv1"] = _sig(gd["Discriminator/D.Block.1.Conv1/Filters"])
        zg = [torch.from_numpy(r2.standard_normal((4, 128)).astype("float32")) for _ in range(2)]
        fl = [torch.from_numpy(r2.randint(0, 10, size=4)).long() for _ in range(2)]
        cost, params, grads = m.gen_grads(zg, fl)
        out["sngan_g_cost_b4"] = [float(cost)]
        gg = dict((n, gr) for (n, _), gr in zip(params, grads))
        out["sngan_ggrad_G.Output"] = _sig(gg["Generator/G.Output/Filters"])
    finally:
        S.BATCH_SIZE = 64

    # ---- secondary variants and the other model families (appended: the entries above keep their RNG streams)
    from oracle import pggan as PG
    from oracle import pix2pix as P2

    rs = np.random.RandomState(7)
    x = torch.from_numpy(rs.standard_normal((2, 8, 8, 16)).astype("float32"))
    np.random.seed(3)
    g = tfshim.Graph(dtype=torch.float32, u_seed=2)
    out["conv_weightnorm_sn"] = _sig(O.Conv2D(g, x, 16, 8, 3, 1, "wn", weightnorm=True, spectral_normed=True,
                                              update_collection=O.NO_OPS))
    out["conv_mask_b3"] = _sig(O.Conv2D(g, x[..., :12].contiguous(), 12, 9, 3, 1, "mk", mask_type=("b", 3)))
    out["conv_separable_s2"] = _sig(O.Conv2D(g, x, 16, 8, 4, 2, "sep", conv_type="separable_conv2d",
                                             channel_multiplier=2))
    out["conv_depthwise_valid"] = _sig(O.Conv2D(g, x, 16, 32, 3, 1, "dw", conv_type="depthwise_conv2d",
                                                channel_multiplier=2, padding="VALID"))
    out["deconv_weightnorm"] = _sig(O.Deconv2D(g, x, 16, 8, 4, name="dc", weight_norm=True))
    out["linear_weightnorm_inputs_norm"] = _sig(O.Linear(g, x.reshape(2, -1), 1024, 6, "lw", weightnorm=True,
                                                         inputs_norm=True))
    out["layer_norm"] = _sig(O.layer_norm(g, "ln", [1, 2, 3], x))
    out["resize_nearest_half"] = _sig(RB.resize_nearest(x, 4, 4))

    np.random.seed(4)
    g = tfshim.Graph(dtype=torch.float32, u_seed=2)
    z = torch.from_numpy(rs.standard_normal((2, 64)).astype("float32"))
    for cls, tag in ((PG.PGGAN, "nvidia"), (PG.PGGANResNet, "resnet")):
        m = cls(1, True, True)
        with torch.no_grad(), g.variable_scope(tag):
            fake = m.get_generator(g, z, 0.3)
            out["pggan_%s_fake" % tag] = _sig(fake)
            out["pggan_%s_logits" % tag] = _sig(m.get_discriminator(g, fake, 0.3, update_collection=O.NO_OPS))

    np.random.seed(5)
    g = tfshim.Graph(dtype=torch.float32, u_seed=2)
    a = torch.from_numpy(rs.uniform(-1, 1, size=(1, 64, 64, 3)).astype("float32"))
    b = torch.from_numpy(rs.uniform(-1, 1, size=(1, 64, 64, 3)).astype("float32"))
    with torch.no_grad(), g.variable_scope("d_net"):
        out["pix2pix_patchgan"] = _sig(P2.unet_d(g, a, b, 8, True, O.NO_OPS))
    with torch.no_grad(), g.variable_scope("d512"):
        out["pix2pix_patchgan_n_layers4"] = _sig(P2.unet_discriminator(g, a, b, 8, True, O.NO_OPS))
    return out


if __name__ == "__main__":
    data = compute()
    with open(os.path.join(HERE, "oracle_golden.json"), "w") as fh:
        json.dump(data, fh, indent=1)
    print("wrote", len(data), "golden entries")
