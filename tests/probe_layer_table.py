"""Per-layer table of every tensor-core launch in one SNGAN-CIFAR D+G pair: records the arguments of the conv / wgrad /
sub-pixel calls of one eager pair, then times each distinct call alone (20 launches in a CUDA graph, CUDA events) and prints
time, algorithmic TFLOP/s and the time lost against the measured bf16 peak.  Not a pytest file; run under gpurun."""
import collections
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gan_lib_tensorflow_b200 import framework  # noqa: E402
from gan_lib_tensorflow_b200 import kernels as K  # noqa: E402
from gan_lib_tensorflow_b200.SNGAN import gan_cifar_resnet as P  # noqa: E402

PEAK = 1685.6e12
calls = collections.OrderedDict()


def _key(name, args, kwargs):
    parts = []
    for a in list(args) + [kwargs[k] for k in sorted(kwargs)]:
        if isinstance(a, torch.Tensor):
            parts.append("T%s%s" % (tuple(a.shape), str(a.dtype)[-4:]))
        else:
            parts.append(repr(a))
    return name + "|" + "|".join(parts)


def hook(name, flops_fn):
    orig = getattr(K, name)

    def wrapped(*args, **kwargs):
        k = _key(name, args, kwargs)
        if k not in calls:
            calls[k] = dict(name=name, args=args, kwargs=kwargs, count=0, flops=flops_fn(*args, **kwargs), fn=orig)
        calls[k]["count"] += 1
        return orig(*args, **kwargs)
    setattr(K, name, wrapped)


def main():
    framework.reset_default_graph("cuda")
    rs = np.random.RandomState(0)
    if len(sys.argv) > 1 and sys.argv[1] == "imagenet":      # config 3: SNGAN ImageNet-128, batch 32 per GPU
        from gan_lib_tensorflow_b200.SNGAN import gan_imagNet_resnet as PI
        tr = PI.Trainer(batch_size=32, seed=0)
        tr.set_real_batch(rs.randint(0, 256, size=(32, 49152)).astype("int32"),
                          rs.randint(0, 1000, size=32).astype("int32"))
    else:
        tr = P.Trainer(batch_size=64, seed=0)
        tr.set_real_batch(rs.randint(0, 256, size=(64, 3072)).astype("int32"),
                          rs.randint(0, 10, size=64).astype("int32"))
    tr.sample_noise()
    tr.d_step(1)
    tr.g_step(1)
    # x, wp, n, h, w, cin, ho, wo, cout, kh, kw, ...
    hook("conv_igemm", lambda x, wp, n, h, w, cin, ho, wo, cout, kh, kw, *a, **k: 2.0 * n * ho * wo * cin * cout * kh * kw)
    hook("conv_wgrad", lambda x, dy, dw, n, h, w, cin, ho, wo, cout, kh, kw, *a, **k: 2.0 * n * ho * wo * cin * cout * kh * kw)
    # sub-pixel calls are credited with the ALGORITHMIC flops of the 3x3 conv over the upsampled tensor
    hook("upconv_fprop", lambda x, we, n, h, w, cin, cout, *a, **k: 2.0 * n * 4 * h * w * cin * cout * 9)
    hook("upconv_dgrad", lambda dy, we, n, h, w, cin, cout, *a, **k: 2.0 * n * 4 * h * w * cin * cout * 9)
    hook("upconv_wgrad", lambda x, dy, dw, n, h, w, cin, cout, *a, **k: 2.0 * n * 4 * h * w * cin * cout * 9)
    tr.sample_noise()
    tr.d_step(1)
    tr.g_step(1)
    torch.cuda.synchronize()
    rows = []
    for k, c in calls.items():
        fn, args, kwargs = c["fn"], c["args"], c["kwargs"]
        for _ in range(3):
            fn(*args, **kwargs)
        torch.cuda.synchronize()
        # 20 back-to-back launches inside one CUDA graph: device time without the Python / ctypes launch cost
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph):
                for _ in range(20):
                    fn(*args, **kwargs)
        torch.cuda.current_stream().wait_stream(side)
        graph.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / 20
        del graph
        ideal = c["flops"] / PEAK * 1e6
        ints = [a for a in args if isinstance(a, int)]
        rows.append(dict(name=c["name"], dims=ints[:11], count=c["count"], us=us, tflops=c["flops"] / us / 1e6,
                         lost_us=(us - ideal) * c["count"], total_us=us * c["count"]))
    rows.sort(key=lambda r: -r["lost_us"])
    tot = sum(r["total_us"] for r in rows)
    print(f"distinct calls {len(rows)}  launches {sum(r['count'] for r in rows)}  sum of standalone time {tot:.0f} us per pair")
    print(f"{'call':13s} {'n,h,w,cin,ho,wo,cout,kh,kw':42s} {'x':>2s} {'us':>8s} {'TF/s':>7s} {'total':>8s} {'lost':>8s}")
    for r in rows:
        print(f"{r['name']:13s} {str(r['dims']):42s} {r['count']:2d} {r['us']:8.1f} {r['tflops']:7.0f} {r['total_us']:8.0f} {r['lost_us']:8.0f}")
    for r in rows:
        r.pop("fn", None)
    json.dump(rows, open("gpurun_out/layer_table.json", "w"))


if __name__ == "__main__":
    main()
