"""Runs SNGAN.gan_cifar_resnet.train (the reference loop, gan_cifar_resnet.py:599-658) for a few hundred iterations on
a structured synthetic dataset with CUDA-graph capture; leaves the sample grids, log.pkl and checkpoints in the given
directory and prints the logged scalars.  Not a pytest file:  python tests/probe_train_loop.py OUT_DIR [ITERS]"""
import os
import pickle
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gan_lib_tensorflow_b200 import framework  # noqa: E402
from gan_lib_tensorflow_b200.SNGAN import gan_cifar_resnet as P  # noqa: E402
from tests import trajectory_feeds as TF_  # noqa: E402


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/train_loop"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    os.makedirs(out, exist_ok=True)
    data, labels = TF_.dataset()

    def epochs(seed):
        def get_epoch():
            rs = np.random.RandomState(seed)
            perm = rs.permutation(len(data))
            for i in range(len(data) // 64):
                idx = perm[i * 64:(i + 1) * 64]
                yield data[idx], labels[idx]
        return get_epoch

    framework.reset_default_graph("cuda")
    t0 = time.time()
    tr = P.train(iters=iters, train_gen=epochs(0), dev_gen=epochs(1), out_dir=out, batch_size=64, capture=True,
                 sample_every=100, flush_until=0, flush_every=100)
    torch.cuda.synchronize()
    dt = time.time() - t0
    with open(os.path.join(out, "log.pkl"), "rb") as fh:
        log = pickle.load(fh)
    print(f"{iters} reference iterations (1 G + {P.N_CRITIC} D steps, dev cost + sample grid + checkpoint every 100) in "
          f"{dt:.1f} s wall = {iters / dt:.1f} it/s incl. eager warm-up, capture and host I/O")
    for k in ("d_cost", "g_cost", "dev_cost"):
        its = sorted(log[k])
        print(k, " ".join(f"{i}:{log[k][i]:.3f}" for i in its[:: max(1, len(its) // 12)]))
    print("files:", sorted(os.listdir(out)), sorted(os.listdir(os.path.join(out, "checkpoint"))))
    print("optimiser steps:", tr.gen_opt.t, tr.disc_opt.t)


if __name__ == "__main__":
    main()
