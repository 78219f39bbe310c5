"""One eager D+G pair bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off` launch lists."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gan_lib_tensorflow_b200 import framework  # noqa: E402
from gan_lib_tensorflow_b200.SNGAN import gan_cifar_resnet as P  # noqa: E402


def main():
    framework.reset_default_graph("cuda")
    tr = P.Trainer(batch_size=64, seed=0)
    rs = np.random.RandomState(0)
    tr.set_real_batch(rs.randint(0, 256, size=(64, 3072)).astype("int32"), rs.randint(0, 10, size=64).astype("int32"))
    for it in range(3):
        tr.sample_noise()
        tr.d_step(1)
        tr.g_step(1)
    torch.cuda.synchronize()
    tr.sample_noise()
    torch.cuda.cudart().cudaProfilerStart()
    tr.d_step(1)
    tr.g_step(1)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("d_loss", tr.d_loss.item(), "g_loss", tr.g_loss.item())


if __name__ == "__main__":
    main()
