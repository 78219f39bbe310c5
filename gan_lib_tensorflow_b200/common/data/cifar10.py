"""CIFAR-10 (Python version) batch reader with the reference's interface (common/data/cifar10.py:9-45):
`load(batch_size, data_dir)` returns (train epoch generator factory, dev epoch generator factory); an epoch yields
(uint8 pixels [batch, 3072] CHW-flattened, labels [batch]) and drops the last partial batch; images and labels are
reshuffled at the start of every epoch with ONE shared NumPy RNG state (get_state / set_state around the two shuffles,
:30-33).  Host code: the hot path starts at the uint8 -> float dequantisation kernel (ganb_preprocess_real)."""
from __future__ import annotations

import os
import pickle

import numpy as np


def unpickle(file):
    with open(file, 'rb') as fo:
        d = pickle.load(fo, encoding='bytes')
    return d[b'data'], d[b'labels']


def cifar_generator(filenames, batch_size, data_dir):
    all_data, all_labels = [], []
    for filename in filenames:
        data, labels = unpickle(os.path.join(data_dir, filename))
        all_data.append(data)
        all_labels.append(labels)
    images = np.concatenate(all_data, axis=0)
    labels = np.concatenate(all_labels, axis=0)

    def get_epoch():
        rng_state = np.random.get_state()
        np.random.shuffle(images)
        np.random.set_state(rng_state)
        np.random.shuffle(labels)
        for i in range(int(len(images) / batch_size)):
            yield (images[i * batch_size:(i + 1) * batch_size], labels[i * batch_size:(i + 1) * batch_size])

    return get_epoch


def load(batch_size, data_dir):
    return (
        cifar_generator(['data_batch_1', 'data_batch_2', 'data_batch_3', 'data_batch_4', 'data_batch_5'], batch_size,
                        data_dir),
        cifar_generator(['test_batch'], batch_size, data_dir),
    )
