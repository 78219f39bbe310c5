"""Generates tests/golden/trajectory_500pairs_<mode>.json: the D / G loss trajectory of 500 D+G training pairs (1000
optimiser steps, Adam included) of the CPU oracle at batch 16 on the seeded feeds of tests/trajectory_feeds.py.

  python tests/golden/make_trajectory.py fp32     # the reference arithmetic
  python tests/golden/make_trajectory.py bf16     # the same graph with bf16 operand rounding at the B200 path's
                                                  # rounding points (how far ANY bf16 implementation drifts from fp32)
About 25 minutes per mode on 4 CPU threads.  The reference itself (TensorFlow 1.5) cannot run here."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
from tests import trajectory_feeds as TF_  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    pairs = int(sys.argv[2]) if len(sys.argv) > 2 else TF_.PAIRS
    torch.set_num_threads(int(os.environ.get("ORACLE_THREADS", "4")))
    from gan_lib_tensorflow_b200 import functional as F
    from oracle import ops as O_ops
    from oracle import resnet_block as ORB
    from oracle import sngan_cifar as O

    batch = TF_.BATCH
    O_ops.BF16_OPERANDS = (mode == "bf16")
    O.BATCH_SIZE = batch
    ORB.SUBPIXEL_RULE = lambda n, h, w, ci, co, k: F.upconv_eligible(batch, h, w, ci, co, k)
    np.random.seed(0)
    om = O.SNGANCifar(dtype=torch.float32, u_seed=2)
    om.build()
    data, labels = TF_.dataset()
    h = batch // 2
    traj = []
    t0 = time.time()
    for s, f in enumerate(TF_.feeds(pairs, batch)):
        x = torch.from_numpy(data[f["idx"]])
        lab = torch.from_numpy(labels[f["idx"]]).long()
        z = [torch.from_numpy(f["z_d"][:h]), torch.from_numpy(f["z_d"][h:])]
        d = om.disc_train_op(s, x, lab, z, torch.from_numpy(f["deq"])).item()
        g = om.gen_train_op(s, [torch.from_numpy(f["z_g"][:batch]), torch.from_numpy(f["z_g"][batch:])],
                            [torch.from_numpy(f["fl"][:batch]).long(), torch.from_numpy(f["fl"][batch:]).long()]).item()
        traj.append([d, g])
        if s % 25 == 0:
            print(f"{mode} pair {s}: d {d:.4f} g {g:.4f}  ({time.time() - t0:.0f}s)", flush=True)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"trajectory_{pairs}pairs_{mode}.json")
    with open(out, "w") as fh:
        json.dump({"mode": mode, "pairs": pairs, "batch": batch, "feeds": "tests/trajectory_feeds.py (dataset seed 11, "
                   "feed seed 12)", "init": "np.random.seed(0), u_seed=2", "d_g": traj}, fh)
    print("wrote", out)


if __name__ == "__main__":
    main()
