"""GPU bring-up probe for the full SNGAN-CIFAR step (not a pytest file; run under gpurun).

Builds the product model and the fp32 CPU oracle from the same NumPy seed, feeds both the same inputs and
compares variables after init, forward outputs, per-variable gradients of the D-step and the G-step.
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from gan_lib_tensorflow_b200 import framework  # noqa: E402
from gan_lib_tensorflow_b200 import functional as F  # noqa: E402
from gan_lib_tensorflow_b200.SNGAN import gan_cifar_resnet as P  # noqa: E402
from oracle import ops as O_ops  # noqa: E402
from oracle import sngan_cifar as O  # noqa: E402


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def main():
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    print(torch.cuda.get_device_name(0))

    # ---------------- oracle
    np.random.seed(0)
    om = O.SNGANCifar(dtype=torch.float32, u_seed=2)
    om.build()

    # ---------------- product
    store = framework.reset_default_graph("cuda", u_seed=2)
    t0 = time.time()
    tr = P.Trainer(batch_size=64, seed=0)
    torch.cuda.synchronize()
    print(f"product build {time.time() - t0:.2f}s; variables: {len(store.vars)}")

    # init parity (NumPy stream consumed in reference order)
    bad = 0
    for name, ov in om.g.vars.items():
        pv = store.vars.get(name)
        if pv is None:
            print("  MISSING in product:", name)
            bad += 1
            continue
        r = rel(pv.data.cpu().numpy().reshape(-1), ov.detach().numpy().reshape(-1))
        if r > 1e-6:
            print(f"  init mismatch {name}: rel={r:.3e}")
            bad += 1
    for name in store.vars:
        if name not in om.g.vars:
            print("  EXTRA in product:", name)
            bad += 1
    print("init parity:", "OK" if bad == 0 else f"{bad} problems")

    # ---------------- inputs
    data, labels = O.synthetic_batch(seed=0)
    rs = np.random.RandomState(1)
    z_d = rs.standard_normal((64, 128)).astype("float32")
    deq = rs.uniform(0, 1 / 128, size=(64, 3072)).astype("float32")
    z_g = rs.standard_normal((128, 128)).astype("float32")
    fl = rs.randint(0, 10, size=128).astype("int32")

    tr.set_real_batch(data, labels)
    tr.z_d.copy_(torch.from_numpy(z_d))
    tr.deq_noise.copy_(torch.from_numpy(deq))
    tr.z_g.copy_(torch.from_numpy(z_g))
    tr.fake_labels.copy_(torch.from_numpy(fl))

    # ---------------- forward parity: G and D
    with store.stat_towers(2):
        fake_p = P.Generator(64, tr.real_labels, noise=tr.z_d, reuse=True)
    lab_t = torch.tensor(labels).long()
    with torch.no_grad():
        f0 = O.Generator(om.g, 32, lab_t[:32], torch.from_numpy(z_d[:32]), reuse=True)
        f1 = O.Generator(om.g, 32, lab_t[32:], torch.from_numpy(z_d[32:]), reuse=True)
        fake_o = torch.cat([f0, f1]).numpy()
    print(f"G forward rel err: {rel(fake_p.data.cpu().numpy(), fake_o):.3e}")
    with torch.no_grad():
        d_o, _ = O.Discriminator(om.g, torch.from_numpy(fake_o), lab_t, update_collection=O_ops.NO_OPS, reuse=True)
    d_p, _ = P.Discriminator(F.Var(torch.from_numpy(fake_o).to(dev)), tr.real_labels, update_collection="NO_OPS",
                             reuse=True)
    print(f"D forward rel err: {rel(d_p.data.cpu().numpy(), d_o.numpy()):.3e}  (oracle |d| max {np.abs(d_o.numpy()).max():.3e})")
    for key, e in store.sn_groups["Discriminator"].entries.items():
        pass
    sig_p = {k: e.scal[0].item() for k, e in store.sn_groups["Discriminator"].entries.items()}
    print("sigma (product) first 4:", list(sig_p.items())[:4])

    # ---------------- D-step gradients
    cost_o, params_o, grads_o = om.disc_grads(torch.tensor(data), lab_t, [torch.from_numpy(z_d[:32]), torch.from_numpy(z_d[32:])],
                                              torch.from_numpy(deq), update_collection=None)
    # product: run the body but stop before Adam -> emulate by lr 0
    tr.disc_opt.set_lr(0.0)
    tr._d_body()
    torch.cuda.synchronize()
    print(f"D loss product {tr.d_loss.item():.6f} oracle {cost_o.item():.6f}")
    worst = 0.0
    for (name, _), g in zip(params_o, grads_o):
        pv = store.vars[name]
        if g is None:
            continue
        r = rel(pv.grad.cpu().numpy(), g.numpy())
        worst = max(worst, r)
        print(f"  dD {name:55s} rel={r:.3e} |g|={float(g.norm()):.3e}")
    print(f"D-step worst grad rel err {worst:.3e}")
    # u parity after the assign
    worst_u = 0.0
    for name, ov in om.g.vars.items():
        if name.endswith("/u"):
            worst_u = max(worst_u, rel(store.vars[name].data.cpu().numpy(), ov.detach().numpy()))
    print(f"u after D-step: worst rel err {worst_u:.3e}")

    # ---------------- G-step gradients
    cost_g, params_g, grads_g = om.gen_grads([torch.from_numpy(z_g[:64]), torch.from_numpy(z_g[64:])],
                                             [torch.from_numpy(fl[:64]).long(), torch.from_numpy(fl[64:]).long()])
    tr.gen_opt.set_lr(0.0)
    tr._g_body()
    torch.cuda.synchronize()
    print(f"G loss product {tr.g_loss.item():.6f} oracle {cost_g.item():.6f}")
    worst = 0.0
    for (name, _), g in zip(params_g, grads_g):
        pv = store.vars[name]
        if g is None:
            continue
        r = rel(pv.grad.cpu().numpy(), g.numpy())
        worst = max(worst, r)
        print(f"  dG {name:55s} rel={r:.3e} |g|={float(g.norm()):.3e}")
    print(f"G-step worst grad rel err {worst:.3e}")

    # ---------------- timing, eager then graphs
    def run_pairs(k):
        for it in range(k):
            tr.sample_noise()
            tr.d_step(1)
            tr.g_step(1)

    run_pairs(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    run_pairs(5)
    e1.record()
    torch.cuda.synchronize()
    print(f"eager: {e0.elapsed_time(e1) / 5:.3f} ms per D+G pair; d_loss {tr.d_loss.item():.4f} g_loss {tr.g_loss.item():.4f}")
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            run_pairs(2)
            tr.capture()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        run_pairs(3)
        torch.cuda.synchronize()
        e0.record()
        run_pairs(20)
        e1.record()
        torch.cuda.synchronize()
        print(f"graphs: {e0.elapsed_time(e1) / 20:.3f} ms per D+G pair; d_loss {tr.d_loss.item():.4f} g_loss {tr.g_loss.item():.4f}")
    except Exception as ex:  # noqa: BLE001
        print("graph capture failed:", repr(ex))
    return 0


if __name__ == "__main__":
    sys.exit(main())
