#!/usr/bin/env python
"""Summarises tests/probe_timeline.py's kernel list: per-kernel warm (in-graph) time per D+G pair, split by stream,
plus the busy / overlapped time of the step.

  python tools/timeline_summary.py gpurun_out/timeline.json [pairs]
"""
import collections
import json
import re
import sys


def main():
    ev = json.load(open(sys.argv[1]))
    pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    ev = [e for e in ev if e["dur"] > 0 and "Memcpy" not in e["name"] and "Memset" not in e["name"]]
    ev.sort(key=lambda e: e["start"])
    t0, t1 = ev[0]["start"], max(e["start"] + e["dur"] for e in ev)
    span = (t1 - t0) / pairs
    # union of busy intervals, and time with >= 2 kernels resident
    pts = []
    for e in ev:
        pts.append((e["start"], 1))
        pts.append((e["start"] + e["dur"], -1))
    pts.sort()
    busy = over = 0.0
    depth, last = 0, pts[0][0]
    for t, d in pts:
        if depth >= 1:
            busy += t - last
        if depth >= 2:
            over += t - last
        depth += d
        last = t
    streams = collections.Counter(e["stream"] for e in ev)
    main_stream = streams.most_common(1)[0][0]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for e in ev:
        name = re.sub(r"^void ", "", e["name"])
        name = re.sub(r"\(.*", "", name)[:64]
        a = agg[name]
        a[0] += 1
        a[1] += e["dur"]
        if e["stream"] == main_stream:
            a[2] += e["dur"]
    tot = sum(a[1] for a in agg.values())
    print(f"pairs {pairs}  span/pair {span:.1f} us  busy/pair {busy / pairs:.1f} us  overlapped/pair {over / pairs:.1f} us  "
          f"sum of kernel time/pair {tot / pairs:.1f} us  streams {dict(streams)}")
    print(f"{'us/pair':>9} {'main-stream':>11} {'n/pair':>7}  kernel")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{a[1] / pairs:9.1f} {a[2] / pairs:11.1f} {a[0] / pairs:7.1f}  {k}")


if __name__ == "__main__":
    main()
