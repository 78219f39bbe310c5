"""Seeded inputs of the long loss-trajectory comparison (tests/golden/make_trajectory.py on the CPU oracle,
tests/test_gpu_step.py::test_loss_trajectory_1k_steps on the B200 path): a 256-image synthetic CIFAR-shaped dataset
and, per D+G pair, the batch indices, generator noise, dequantisation noise and fake labels.  Pure NumPy, so both sides
regenerate identical feeds from the seeds instead of shipping them."""
import numpy as np

BATCH = 16
PAIRS = 500          # 500 critic steps + 500 generator steps = 1000 optimiser steps
DATASET = 256


def dataset(seed=11):
    """256 smooth class-dependent images (colour ramps + noise), CHW-flattened int32 pixels in [0, 255], labels 0..9."""
    rs = np.random.RandomState(seed)
    labels = rs.randint(0, 10, size=DATASET).astype("int32")
    yy, xx = np.mgrid[0:32, 0:32] / 31.0
    base = np.stack([np.sin((c + 1) * xx * 1.3) * 0.5 + 0.5 for c in range(10)])
    imgs = np.stack([np.stack([base[l], base[(l + 3) % 10].T, yy * 0 + l / 9.0]) for l in labels])
    imgs = np.clip(imgs + 0.05 * rs.standard_normal(imgs.shape), 0, 1)
    return (imgs * 255).astype("int32").reshape(DATASET, 3072), labels


def feeds(pairs=PAIRS, batch=BATCH, seed=12):
    """Yields one dict per D+G pair: idx (batch indices), z_d, deq, z_g, fl -- the tensors TF draws with its own RNG
    (gan_cifar_resnet.py:240, 335, 467) and the loader's batch."""
    rs = np.random.RandomState(seed)
    for _ in range(pairs):
        yield dict(idx=rs.randint(0, DATASET, size=batch),
                   z_d=rs.standard_normal((batch, 128)).astype("float32"),
                   deq=rs.uniform(0, 1 / 128, size=(batch, 3072)).astype("float32"),
                   z_g=rs.standard_normal((2 * batch, 128)).astype("float32"),
                   fl=rs.randint(0, 10, size=2 * batch).astype("int32"))
