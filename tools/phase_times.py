#!/usr/bin/env python
"""Phase durations of the D+G pair from a kernel timeline (tests/probe_timeline.py JSON: name / start / dur per kernel).
Markers (one per pair unless noted): preprocess_real (critic pass of the D step starts), gan_loss #1 (its forward ends),
adam #1 (its backward ends), gan_loss #2 (critic forward of the G step ends), conv_halo_narrow #3 / the first kernel
after it (generator backward starts), adam #2 (generator backward ends).

  python tools/phase_times.py gpurun_out/timeline.json
"""
import json
import sys


def main():
    ev = sorted(json.load(open(sys.argv[1])), key=lambda e: e["start"])
    starts = [i for i, e in enumerate(ev) if "preprocess_real" in e["name"]]
    rows = []
    for k in range(len(starts) - 1):
        # a pair "starts" at the kernel after the previous pair's last adam/pack; use the marker-to-marker period instead
        seg = ev[starts[k]:starts[k + 1]]
        t0 = seg[0]["start"]
        loss = [e for e in seg if "gan_loss" in e["name"]]
        adam = [e for e in seg if "adam_kernel" in e["name"]]
        narrow = [e for e in seg if "conv_halo_narrow" in e["name"]]
        if len(loss) < 2 or len(adam) < 2:
            continue
        marks = [("critic fwd (D step)", loss[0]["start"]), ("critic bwd + update", adam[0]["start"] + adam[0]["dur"]),
                 ("critic fwd (G step)", loss[1]["start"])]
        nb = [e for e in narrow if e["start"] > loss[1]["start"]]
        if nb:
            marks.append(("critic bwd (G step)", nb[0]["start"]))
        marks.append(("generator bwd + update", adam[1]["start"] + adam[1]["dur"]))
        marks.append(("packs + both generator fwd (next pair)", seg[-1]["start"] + seg[-1]["dur"]))
        prev = t0
        row = []
        for name, t in marks:
            row.append((name, t - prev))
            prev = t
        row.append(("period", ev[starts[k + 1]]["start"] - t0))
        rows.append(row)
    if not rows:
        print("no complete pair between markers")
        return
    for j, (name, _) in enumerate(rows[0]):
        print(f"{sum(r[j][1] for r in rows) / len(rows):9.1f} us  {name}")


if __name__ == "__main__":
    main()
