#!/usr/bin/env python
"""From a chrome trace of graph-replayed D+G pairs (tests/probe_timeline.py): the time during which NO tensor-core
convolution kernel is resident, attributed to the kernels that run in those windows (split evenly when several
do).  That "exposed" time is what separates the step from the sum of its tensor-core kernels.

  python tools/exposed_time.py gpurun_out/timeline_chrome.json [pairs]
"""
import collections
import json
import re
import sys

TENSOR = re.compile(r"conv_pair|conv_igemm|conv_halo|conv_wgrad")


def main():
    d = json.load(open(sys.argv[1]))
    pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    ev = [e for e in d["traceEvents"] if e.get("cat") == "kernel"]
    pts = []
    for i, e in enumerate(ev):
        pts.append((e["ts"], 1, i))
        pts.append((e["ts"] + e["dur"], 0, i))
    pts.sort()
    active = set()
    exposed = collections.defaultdict(float)
    idle = 0.0
    n_tensor = 0
    last = pts[0][0]
    for t, kind, i in pts:
        dt = t - last
        if dt > 0:
            if n_tensor == 0:
                if active:
                    for j in active:
                        exposed[short(ev[j]["name"])] += dt / len(active)
                else:
                    idle += dt
        if kind == 1:
            active.add(i)
            n_tensor += bool(TENSOR.search(ev[i]["name"]))
        else:
            active.discard(i)
            n_tensor -= bool(TENSOR.search(ev[i]["name"]))
        last = t
    tot = sum(exposed.values())
    print(f"per pair: exposed (no tensor kernel resident) {tot / pairs:.1f} us + idle {idle / pairs:.1f} us")
    for k, v in sorted(exposed.items(), key=lambda kv: -kv[1]):
        print(f"{v / pairs:9.1f}  {k}")


def short(name):
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*", "", name)[:70]


if __name__ == "__main__":
    main()
