"""Kernel timeline of graph-replayed D+G pairs (torch.profiler / CUPTI): dumps (name, stream, start_us, dur_us) per
kernel so that gaps and cross-stream overlap can be analysed offline. Not a pytest file; run under gpurun."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gan_lib_tensorflow_b200 import framework  # noqa: E402
from gan_lib_tensorflow_b200.SNGAN import gan_cifar_resnet as P  # noqa: E402


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.json"
    pair = "--steps" not in sys.argv      # default: the D+G pair schedule (Trainer.pair_step); --steps: d_step + g_step

    def run_pair():
        if pair:
            tr.pair_step(1)
        else:
            tr.d_step(1)
            tr.g_step(1)
    framework.reset_default_graph("cuda")
    tr = P.Trainer(batch_size=64, seed=0)
    rs = np.random.RandomState(0)
    tr.set_real_batch(rs.randint(0, 256, size=(64, 3072)).astype("int32"), rs.randint(0, 10, size=64).astype("int32"))
    for it in range(2):
        tr.sample_noise()
        tr.d_step(1)
        tr.g_step(1)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        tr.capture()
    torch.cuda.current_stream().wait_stream(s)
    for it in range(5):
        run_pair()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(50):
        run_pair()
    e1.record()
    torch.cuda.synchronize()
    print("unprofiled: %.1f us per pair (50 replays)" % (e0.elapsed_time(e1) * 1e3 / 50))
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for it in range(3):
            run_pair()
        torch.cuda.synchronize()
    ev = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            ev.append({"name": e.name, "start": e.time_range.start, "dur": e.time_range.end - e.time_range.start,
                       "stream": getattr(e, "stream", None) if hasattr(e, "stream") else None})
    try:
        prof.export_chrome_trace(out.replace(".json", "_chrome.json"))
    except Exception as ex:  # noqa: BLE001
        print("chrome trace export failed:", ex)
    with open(out, "w") as fh:
        json.dump(ev, fh)
    print(len(ev), "device events")


if __name__ == "__main__":
    main()
