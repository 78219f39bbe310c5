"""1000 reference iterations (1 G-step + 5 D-steps each, gan_cifar_resnet.py:599-620) on synthetic images; prints the
D / G loss trajectory (BASELINE.md band check needs real CIFAR-10, which is not available offline)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gan_lib_tensorflow_b200 import framework  # noqa: E402
from gan_lib_tensorflow_b200.SNGAN import gan_cifar_resnet as P  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    framework.reset_default_graph("cuda")
    tr = P.Trainer(batch_size=64, seed=0)
    rs = np.random.RandomState(0)
    # "dataset": 4096 smooth synthetic images in 10 classes (class-dependent colour ramps + noise)
    n = 4096
    labels = rs.randint(0, 10, size=n).astype("int32")
    yy, xx = np.mgrid[0:32, 0:32] / 31.0
    base = np.stack([np.sin((c + 1) * xx * 1.3) * 0.5 + 0.5 for c in range(10)])           # [10,32,32]
    imgs = np.stack([np.stack([base[l], base[(l + 3) % 10].T, yy * 0 + l / 9.0]) for l in labels])  # [n,3,32,32]
    imgs = np.clip(imgs + 0.05 * rs.standard_normal(imgs.shape), 0, 1)
    data = (imgs * 255).astype("int32").reshape(n, 3072)

    def batches():
        while True:
            idx = rs.randint(0, n, size=64)
            yield data[idx], labels[idx]

    gen = batches()
    for it in range(2):
        tr.train_iteration(it, gen)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        tr.capture()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    traj = []
    t0 = time.time()
    for it in range(2, iters):
        d, g = tr.train_iteration(it, gen)
        if it % 10 == 0 or it < 20:
            traj.append((it, float(d.item()), float(g.item())))
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(f"{iters - 2} reference iterations (1 G + 5 D) in {dt:.1f}s = {(iters - 2) / dt:.1f} it/s")
    for it, d, g in traj[:: max(1, len(traj) // 40)]:
        print(f"  it {it:5d}  d_cost {d:8.4f}  g_cost {g:8.4f}")
    ds = np.array([t[1] for t in traj]); gs = np.array([t[2] for t in traj])
    print(json.dumps({"iters": iters, "its_per_s": (iters - 2) / dt, "d_last100_mean": float(ds[-10:].mean()),
                      "g_last100_mean": float(gs[-10:].mean()), "d_min": float(ds.min()), "d_max": float(ds.max()),
                      "finite": bool(np.isfinite(ds).all() and np.isfinite(gs).all())}))


if __name__ == "__main__":
    main()
