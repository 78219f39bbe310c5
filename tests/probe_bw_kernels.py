"""Algorithmic bandwidth of the streaming (HBM-bound) kernels of the headline step, each timed alone with CUDA events
over inputs that rotate through buffers larger than the 126 MB L2 (so reads come from HBM): batch statistics, normalise +
activation forward, the two kernels of its backward pass, bias-gradient column sums, a plain torch copy as the yardstick.
Prints GB/s = algorithmic bytes / time and the fraction of MEASURED_PEAKS.json's copy bandwidth.  Not a pytest file.
Tuning hooks (read once per process): GANB_STATS_BPS, GANB_V8_FWD_BPS, GANB_V8_BWD_BPS = blocks per SM."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gan_lib_tensorflow_b200 import kernels as K  # noqa: E402

BF16 = torch.bfloat16


def peak():
    try:
        return float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        return 6556.0


def timed(fn, sets, reps=24):
    """us per call, GPU time: the calls are captured into ONE CUDA graph (eager launches of these two- and three-kernel
    ops are CPU-bound at ~25 us per call) and the replay is timed with events."""
    for i in range(4):
        fn(*sets[i % len(sets)])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(*sets[i % len(sets)])
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (2 * reps)     # us


def main():
    pk = peak()
    tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("GANB_"))
    print(f"# peak {pk:.0f} GB/s  {tag}")
    for (n, h, c) in ((128, 32, 256), (128, 16, 256), (128, 8, 256), (256, 32, 128)):
        elems = n * h * h * c
        nbuf = max(3, int(400e6 // (elems * 2)) + 1)
        xs = [torch.randn(n, h, h, c, device="cuda").to(BF16) for _ in range(nbuf)]
        dzs = [torch.randn(n, h, h, c, device="cuda").to(BF16) for _ in range(nbuf)]
        outs = [torch.empty(n, h, h, c, device="cuda", dtype=BF16) for _ in range(2)]
        mean, rstd = K.bn_stats(xs[0], n, h * h, c, 2, 1e-5)
        mb = elems * 2 / 1e6
        rows = []
        t = timed(lambda x, o: o.copy_(x), [(x, outs[i % 2]) for i, x in enumerate(xs)])
        rows.append(("torch copy (read + write)", t, 2 * mb))
        t = timed(lambda x: K.bn_stats(x, n, h * h, c, 2, 1e-5), [(x,) for x in xs])
        rows.append(("bn_stats (partial + finalize)", t, mb))
        t = timed(lambda x, o: K.norm_act_fwd(x, n, h, h, c, mean, rstd, 2, None, None, None, "relu", False, BF16, out=o),
                  [(x, outs[i % 2]) for i, x in enumerate(xs)])
        rows.append(("norm_act_fwd", t, 2 * mb))
        t = timed(lambda x, dz: K.norm_act_bwd(x, dz, 0, n, h, h, c, mean, rstd, 2, None, None, None, "relu", False, None,
                                               None, None, BF16), list(zip(xs, dzs)))
        rows.append(("norm_act_bwd (reduce + finalize + apply)", t, 5 * mb))
        bias = torch.zeros(c, device="cuda")
        t = timed(lambda dz: K.colsum(dz.view(-1, c), n * h * h, c, bias, 0.0), [(dz,) for dz in dzs])
        rows.append(("colsum (partial + finalize)", t, mb))
        print(f"[{n} x {h} x {h} x {c}] bf16, {mb:.1f} MB per tensor, {nbuf} rotating buffers")
        for name, t, b in rows:
            print(f"  {name:44s} {t:8.1f} us  {b / t * 1e3:7.0f} GB/s  {b / t * 1e3 / pk:5.2f} of peak")
        del xs, dzs, outs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
