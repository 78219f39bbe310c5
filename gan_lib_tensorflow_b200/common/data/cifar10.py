"""CIFAR-10 (Python version) batches behind the call the reference scripts make:
`train_gen, dev_gen = lib.data.cifar10.load(batch_size, data_dir)`; calling either object starts an epoch that yields
(uint8 pixels [batch, 3072] CHW-flattened, int labels [batch]) and drops the trailing partial batch
(reference: common/data/cifar10.py:9-45).

Design: each split is ONE contiguous uint8 matrix read once, plus a persistent ORDER vector.  An epoch draws one
permutation from NumPy's global RNG and composes it onto the order -- the same sequence of batches, and the same RNG
consumption, as the reference's pair of in-place shuffles under a saved / restored RNG state (a shuffle of the rows is
the row-gather by `permutation(n)` of the same stream) -- and every batch is a fresh gather, so a consumer may keep it
across epochs (the reference hands out views of the array it reshuffles).  Host code; the hot path starts at the
dequantisation kernel (ganb_preprocess_real)."""
from __future__ import annotations

import os
import pickle

import numpy as np

TRAIN_FILES = tuple('data_batch_%d' % i for i in range(1, 6))
TEST_FILES = ('test_batch',)


class CifarSplit:
    def __init__(self, files, batch_size: int, data_dir: str):
        pixels, labels = [], []
        for name in files:
            with open(os.path.join(data_dir, name), 'rb') as fh:
                record = pickle.load(fh, encoding='bytes')
            pixels.append(np.asarray(record[b'data'], dtype=np.uint8))
            labels.append(np.asarray(record[b'labels'], dtype=np.int64))
        self.pixels = np.ascontiguousarray(np.concatenate(pixels, axis=0))
        self.labels = np.concatenate(labels, axis=0)
        self.batch_size = int(batch_size)
        self.order = np.arange(len(self.labels))

    def __len__(self):
        return len(self.labels) // self.batch_size          # full batches per epoch

    def __call__(self):
        """One epoch.  The permutation is drawn when the epoch starts to be consumed (generator semantics, like the
        reference's get_epoch)."""
        self.order = self.order[np.random.permutation(len(self.order))]
        b = self.batch_size
        for k in range(len(self)):
            rows = self.order[k * b:(k + 1) * b]
            yield self.pixels[rows], self.labels[rows]


def load(batch_size, data_dir):
    return CifarSplit(TRAIN_FILES, batch_size, data_dir), CifarSplit(TEST_FILES, batch_size, data_dir)
