"""Metric logger with the reference's interface (common/plot.py): plot(name, value) records a scalar for the current
iteration, tick() advances the iteration, flush() prints the means since the last flush, extends log.pkl and -- when
matplotlib is installed -- redraws one <name>.jpg curve per metric.

Values may be device scalars (loss tensors / Vars): they are kept as they are and only read back inside flush(), so
logging a loss never synchronises the training stream (the reference fetches every scalar with session.run each
iteration, SNGAN/gan_cifar_resnet.py:605-632)."""
from __future__ import annotations

import collections
import os
import pickle

import numpy as np

_since_beginning = collections.defaultdict(lambda: {})
_since_last_flush = collections.defaultdict(lambda: {})

_iter = [0]
_out_dir = ['.']


def set_output_dir(path):
    _out_dir[0] = path


def reset():
    _since_beginning.clear()
    _since_last_flush.clear()
    _iter[0] = 0


def tick():
    _iter[0] += 1


def plot(name, value):
    if hasattr(value, 'data') and not isinstance(value, np.ndarray):   # framework.Var
        value = value.data
    if hasattr(value, 'detach'):
        value = value.detach().reshape(-1)[:1].clone()    # a device-side copy: later steps may overwrite static buffers
    _since_last_flush[name][_iter[0]] = value


def _host(v):
    return float(v.cpu().item()) if hasattr(v, 'cpu') else float(v)


def flush():
    prints = []
    for name, vals in _since_last_flush.items():
        vals = {k: _host(v) for k, v in vals.items()}
        prints.append("{}: {}".format(name, np.mean(list(vals.values()))))
        _since_beginning[name].update(vals)
        try:
            import matplotlib
            matplotlib.use('Agg')
            import matplotlib.pyplot as plt
        except ImportError:
            continue
        x_vals = np.sort(list(_since_beginning[name].keys()))
        y_vals = [_since_beginning[name][x] for x in x_vals]
        plt.clf()
        plt.plot(x_vals, y_vals)
        plt.xlabel('iteration')
        plt.ylabel(name)
        plt.savefig(os.path.join(_out_dir[0], name.replace(' ', '_') + '.jpg'))
    print("iter {}\n{}".format(_iter[0], ", ".join(prints)))
    _since_last_flush.clear()
    with open(os.path.join(_out_dir[0], 'log.pkl'), 'wb') as f:
        pickle.dump({k: dict(v) for k, v in _since_beginning.items()}, f, pickle.HIGHEST_PROTOCOL)
