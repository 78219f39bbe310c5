"""Metric logger behind the three calls of the reference's loops (`lib.plot.plot(name, value)`, `lib.plot.tick()`,
`lib.plot.flush()`; reference: common/plot.py): plot records a scalar under the current iteration, tick advances the
iteration, flush prints the per-metric means since the previous flush as `iter N` + `name: mean, ...`, merges them into
the history, rewrites `log.pkl` ({name: {iteration: value}}) and -- when matplotlib is importable -- one `<name>.jpg`
curve per metric.

Design: values may be device scalars (loss tensors / Vars).  They are stored as detached one-element device copies and
only read back inside flush(), so logging a loss never synchronises the training stream; the reference fetches every
scalar with session.run on every iteration (SNGAN/gan_cifar_resnet.py:605-632)."""
from __future__ import annotations

import os
import pickle
from collections import OrderedDict

import numpy as np


class MetricLog:
    def __init__(self, out_dir: str = '.'):
        self.out_dir = out_dir
        self.iteration = 0
        self.history: "OrderedDict[str, dict]" = OrderedDict()     # everything flushed so far
        self.pending: "OrderedDict[str, dict]" = OrderedDict()     # recorded since the last flush

    @staticmethod
    def _hold(value):
        data = getattr(value, 'data', value) if not isinstance(value, np.ndarray) else value   # framework.Var -> tensor
        if hasattr(data, 'detach'):
            return data.detach().reshape(-1)[:1].clone()     # device-side copy: static graph buffers get overwritten
        return data

    @staticmethod
    def _read(value) -> float:
        return float(value.cpu().item()) if hasattr(value, 'cpu') else float(value)

    def record(self, name, value):
        self.pending.setdefault(name, {})[self.iteration] = self._hold(value)

    def advance(self):
        self.iteration += 1

    def _draw(self, name):
        try:
            import matplotlib
            matplotlib.use('Agg')
            import matplotlib.pyplot as plt
        except ImportError:
            return
        series = self.history[name]
        xs = sorted(series)
        plt.clf()
        plt.plot(xs, [series[x] for x in xs])
        plt.xlabel('iteration')
        plt.ylabel(name)
        plt.savefig(os.path.join(self.out_dir, name.replace(' ', '_') + '.jpg'))

    def flush(self):
        summary = []
        for name, values in self.pending.items():
            host = {it: self._read(v) for it, v in values.items()}
            summary.append("{}: {}".format(name, np.mean(list(host.values()))))
            self.history.setdefault(name, {}).update(host)
            self._draw(name)
        print("iter {}\n{}".format(self.iteration, ", ".join(summary)))
        self.pending = OrderedDict()
        with open(os.path.join(self.out_dir, 'log.pkl'), 'wb') as fh:
            pickle.dump({k: dict(v) for k, v in self.history.items()}, fh, pickle.HIGHEST_PROTOCOL)


_log = MetricLog()


def set_output_dir(path):
    _log.out_dir = path


def reset():
    global _log
    _log = MetricLog(_log.out_dir)


def plot(name, value):
    _log.record(name, value)


def tick():
    _log.advance()


def flush():
    _log.flush()
