"""CUDA-graph capture of the two-optimiser trainers (training.TwoPlayer.capture): one captured step must reproduce
the eager step from the same state on the same inputs -- gradients, loss, updated parameters and the spectral-norm u
vectors -- for ACGAN (config 2, with the gradient penalty) and Pix2Pix (config 4, three D evaluations that each
re-assign u)."""
import numpy as np
import pytest
import torch

from tests.test_gpu_ops import env, rel  # noqa: F401

pytestmark = pytest.mark.gpu


class Snapshot:
    """Every variable (weights, u), both Adam slot buffers and the step counters of a TwoPlayer trainer."""

    def __init__(self, store, players, trainer=None):
        self.store, self.players = store, players
        self.vars = {k: v.data.clone() for k, v in store.vars.items()}
        self.slots = {r: (f.m.clone(), f.v.clone()) for r, f in store.flat.items()}
        self.t = {k: o.t for k, o in players.opt.items()}
        self.trainer, self.global_step = trainer, getattr(trainer, "global_step", 0)

    def restore(self):
        st = self.store
        for k, t in self.vars.items():
            st.vars[k].data.copy_(t)
        for r, (m, v) in self.slots.items():
            st.flat[r].m.copy_(m)
            st.flat[r].v.copy_(v)
        for k, t in self.t.items():
            self.players.opt[k].t = t
        if self.trainer is not None:
            self.trainer.global_step = self.global_step
        for r in st.flat:
            st.bump(r)
            st.bump_u(r)
            g = st.pack_groups.get(r)
            if g is not None and g.entries:
                g.refresh()


def _state(store, root):
    f = store.flat[root]
    u = {k: v.data.clone() for k, v in store.vars.items() if k.startswith(root) and k.endswith("/u")}
    return f.grads.clone(), f.params.clone(), u


def _compare(tag, eager, graphed, tol=2e-4):
    (ge, pe, ue), (gg, pg, ug) = eager, graphed
    r = rel(gg.cpu().numpy(), ge.cpu().numpy())
    print(f"{tag}: gradients rel {r:.2e}", end="")
    assert r < tol, (tag, r)
    # Adam with beta1 = 0 moves every weight by ~lr * sign(g): a gradient that differs in its last bits (fp32 atomics in
    # the reductions) can flip a near-zero element, so parameters are compared by the fraction that moved differently
    moved = float(((pe - pg).abs() > 1e-7).float().mean())
    print(f"  parameters differing {moved:.2e}", end="")
    assert moved < 2e-3, (tag, moved)
    for k in ue:
        ru = rel(ug[k].cpu().numpy(), ue[k].cpu().numpy())
        assert ru < 1e-5, (tag, k, ru)
    print(f"  u vectors {len(ue)} ok")


def test_acgan_captured_steps_match_eager(env):
    store, _ = env
    from gan_lib_tensorflow_b200.ACGAN import train as AT

    b = 16
    tr = AT.Trainer(batch_size=b, gradient_penalty=True, seed=0)
    rs = np.random.RandomState(5)
    real = tr.preprocess(torch.from_numpy(rs.randint(0, 256, size=(b, 3072)).astype("int32")).cuda(), None)
    labels = torch.from_numpy(rs.randint(0, 10, size=b).astype("int32")).cuda()
    z = torch.from_numpy(rs.standard_normal((b, 128)).astype("float32")).cuda()
    fl = torch.from_numpy(rs.randint(0, 10, size=b).astype("int32")).cuda()
    alpha = torch.from_numpy(rs.uniform(size=b).astype("float32")).cuda()
    tr.d_step(real, labels, z, fl, alpha)      # eager warm-up: workspaces, tables, operand copies
    tr.g_step(z, fl)
    snap = Snapshot(store, tr.players, tr)

    ld = float(tr.d_step(real, labels, z, fl, alpha).data.reshape(-1)[0])
    d_eager = _state(store, "d_net")
    lg = float(tr.g_step(z, fl).data.reshape(-1)[0])
    g_eager = _state(store, "g_net")

    snap.restore()
    tr.capture()
    snap.restore()
    assert tr.players.captured("d") and tr.players.captured("g")
    assert tr.players.launches("d") > 100 and tr.players.launches("g") > 50
    ld_g = float(tr.d_step(real, labels, z, fl, alpha)[0])
    d_graph = _state(store, "d_net")
    lg_g = float(tr.g_step(z, fl)[0])
    g_graph = _state(store, "g_net")
    print(f"acgan losses eager {ld:.6f} {lg:.6f} graph {ld_g:.6f} {lg_g:.6f}; launches d {tr.players.launches('d')} "
          f"g {tr.players.launches('g')}")
    assert abs(ld - ld_g) < 1e-4 * max(1.0, abs(ld)) and abs(lg - lg_g) < 1e-4 * max(1.0, abs(lg))
    _compare("acgan d_step", d_eager, d_graph)
    _compare("acgan g_step", g_eager, g_graph)
    # replays keep training: a second replay from the updated state changes the loss
    ld2 = float(tr.d_step(real, labels, z, fl, alpha)[0])
    assert np.isfinite(ld2) and ld2 != ld_g


def test_pix2pix_captured_steps_match_eager(env):
    store, _ = env
    from gan_lib_tensorflow_b200.Pix2Pix import train as PT

    b, size, ngf = 1, 512, 8      # 512x512 keeps the bottleneck instance norm well-posed (tests/test_gpu_wide.py)
    tr = PT.Trainer(ngf=ngf, ndf=8, size=size, seed=0)
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand(b, size, size, 3, device="cuda", generator=g) * 2 - 1
    t = torch.rand(b, size, size, 3, device="cuda", generator=g) * 2 - 1
    hw = [size // 128, size // 64, size // 32]
    masks = [(torch.rand(b, max(s, 1), max(s, 1), ngf * 8, device="cuda", generator=g) < 0.5).float() for s in hw]
    tr.d_step(x, t, masks)
    tr.g_step(x, t, masks)
    snap = Snapshot(store, tr.players, tr)

    ld = float(tr.d_step(x, t, masks).data.reshape(-1)[0])
    d_eager = _state(store, "d_net")
    lg = float(tr.g_step(x, t, masks).data.reshape(-1)[0])
    g_eager = _state(store, "g_net")

    snap.restore()
    tr.capture(x, t, masks)
    snap.restore()
    ld_g = float(tr.d_step(x, t, masks)[0])
    d_graph = _state(store, "d_net")
    lg_g = float(tr.g_step(x, t, masks)[0])
    g_graph = _state(store, "g_net")
    print(f"pix2pix losses eager {ld:.6f} {lg:.6f} graph {ld_g:.6f} {lg_g:.6f}")
    assert abs(ld - ld_g) < 1e-4 * max(1.0, abs(ld)) and abs(lg - lg_g) < 1e-4 * max(1.0, abs(lg))
    _compare("pix2pix d_step", d_eager, d_graph)
    _compare("pix2pix g_step", g_eager, g_graph)
    with pytest.raises(ValueError):
        tr.d_step(x, t, None)
