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


def run_oracle(bf16, data, labels, z_d, deq, z_g, fl):
    O_ops.BF16_OPERANDS = bf16
    np.random.seed(0)
    om = O.SNGANCifar(dtype=torch.float32, u_seed=2)
    om.build()
    lab_t = torch.tensor(labels).long()
    res = {"om": om}
    with torch.no_grad():
        f0 = O.Generator(om.g, 32, lab_t[:32], torch.from_numpy(z_d[:32]), reuse=True)
        f1 = O.Generator(om.g, 32, lab_t[32:], torch.from_numpy(z_d[32:]), reuse=True)
        res["fake"] = torch.cat([f0, f1]).numpy()
    cost_o, params_o, grads_o = om.disc_grads(torch.tensor(data), lab_t,
                                              [torch.from_numpy(z_d[:32]), torch.from_numpy(z_d[32:])],
                                              torch.from_numpy(deq), update_collection=None)
    res["d_cost"] = cost_o.item()
    res["d_grads"] = {n: g.numpy() for (n, _), g in zip(params_o, grads_o) if g is not None}
    res["u"] = {n: v.detach().numpy().copy() for n, v in om.g.vars.items() if n.endswith("/u")}
    cost_g, params_g, grads_g = om.gen_grads([torch.from_numpy(z_g[:64]), torch.from_numpy(z_g[64:])],
                                             [torch.from_numpy(fl[:64]).long(), torch.from_numpy(fl[64:]).long()])
    res["g_cost"] = cost_g.item()
    res["g_grads"] = {n: g.numpy() for (n, _), g in zip(params_g, grads_g) if g is not None}
    O_ops.BF16_OPERANDS = False
    return res


def compare(tag, prod, ref, floor):
    worst = 0.0
    lines = []
    for name, g in ref.items():
        gn = float(np.linalg.norm(g))
        if gn < floor:  # gradients that are analytically zero (biases in front of a batch norm)
            continue
        r = rel(prod[name], g)
        worst = max(worst, r)
        lines.append(f"    {name:58s} rel={r:.3e} |g|={gn:.3e}")
    print(f"  {tag}: worst rel err {worst:.3e}")
    return worst, lines


def main():
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    print(torch.cuda.get_device_name(0))
    data, labels = O.synthetic_batch(seed=0)
    rs = np.random.RandomState(1)
    z_d = rs.standard_normal((64, 128)).astype("float32")
    deq = rs.uniform(0, 1 / 128, size=(64, 3072)).astype("float32")
    z_g = rs.standard_normal((128, 128)).astype("float32")
    fl = rs.randint(0, 10, size=128).astype("int32")

    t0 = time.time()
    ref32 = run_oracle(False, data, labels, z_d, deq, z_g, fl)
    ref16 = run_oracle(True, data, labels, z_d, deq, z_g, fl)
    print(f"oracles done in {time.time() - t0:.1f}s")

    store = framework.reset_default_graph("cuda", u_seed=2)
    tr = P.Trainer(batch_size=64, seed=0)
    torch.cuda.synchronize()
    om = ref32["om"]
    bad = [n for n in om.g.vars if n not in store.vars] + [n for n in store.vars if n not in om.g.vars]
    print("variable name parity:", "OK" if not bad else bad)

    tr.set_real_batch(data, labels)
    tr.z_d.copy_(torch.from_numpy(z_d))
    tr.deq_noise.copy_(torch.from_numpy(deq))
    tr.z_g.copy_(torch.from_numpy(z_g))
    tr.fake_labels.copy_(torch.from_numpy(fl))

    with store.stat_towers(2):
        fake_p = P.Generator(64, tr.real_labels, noise=tr.z_d, reuse=True)
    fp = fake_p.data.cpu().numpy()
    print(f"G forward rel err: vs fp32 oracle {rel(fp, ref32['fake']):.3e}   vs bf16-operand oracle {rel(fp, ref16['fake']):.3e}")

    tr.disc_opt.set_lr(0.0)
    tr._d_body()
    torch.cuda.synchronize()
    print(f"D loss product {tr.d_loss.item():.6f}  fp32 oracle {ref32['d_cost']:.6f}  bf16 oracle {ref16['d_cost']:.6f}")
    prod = {n: v.grad.cpu().numpy().copy() for n, v in store.vars.items() if v.trainable and v.grad is not None}
    compare("D-step grads vs fp32 oracle", prod, ref32["d_grads"], 1e-7)
    w16, lines = compare("D-step grads vs bf16-operand oracle", prod, ref16["d_grads"], 1e-7)
    print("\n".join(lines))
    wu = max(rel(store.vars[n].data.cpu().numpy(), u) for n, u in ref32["u"].items())
    print(f"  u after the D-step: worst rel err {wu:.3e}")

    tr.gen_opt.set_lr(0.0)
    tr._g_body()
    torch.cuda.synchronize()
    print(f"G loss product {tr.g_loss.item():.6f}  fp32 oracle {ref32['g_cost']:.6f}  bf16 oracle {ref16['g_cost']:.6f}")
    prod = {n: v.grad.cpu().numpy().copy() for n, v in store.vars.items() if v.trainable and v.grad is not None}
    compare("G-step grads vs fp32 oracle", prod, ref32["g_grads"], 1e-7)
    w16, lines = compare("G-step grads vs bf16-operand oracle", prod, ref16["g_grads"], 1e-7)
    print("\n".join(lines))

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
