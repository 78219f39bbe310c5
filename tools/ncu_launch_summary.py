#!/usr/bin/env python
"""Share of kernel time by kernel from an ncu launch list
(ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv python bench.py --steps 2 --warmup 1).
Per-launch times under ncu are cold-cache and serialised: the SHARES are what should agree with the live measurement of
bench.py (roofline.step.kernel_time_share).

  python tools/ncu_launch_summary.py gpurun_out/launches.csv > profiles/rNN_launches_bench_summary.txt
"""
import collections
import csv
import re
import sys

TENSOR = re.compile(r"conv_pair|conv_igemm|conv_halo|conv_wgrad|upconv_(fprop|dgrad|wgrad)")


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    H = rows[hdr]
    per = collections.OrderedDict()
    n = 0
    scale = {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3}
    for r in rows[hdr + 1:]:
        if len(r) < len(H):
            continue
        rec = dict(zip(H, r))
        if rec.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", rec["Kernel Name"]).replace("void ", "")
        t = float(rec["Metric Value"].replace(",", "")) * scale.get(rec.get("Metric Unit", "us"), 1.0)
        a = per.setdefault(name, [0.0, 0])
        a[0] += t
        a[1] += 1
        n += 1
    total = sum(v[0] for v in per.values())
    tensor = sum(v[0] for k, v in per.items() if TENSOR.search(k))
    ours = sum(v[0] for k, v in per.items() if "ganb::" in k)
    print(f"total {total:.0f} us over {n} launches; libganb200 kernels {100 * ours / total:.1f} %; "
          f"tensor-core convolution kernels {100 * tensor / total:.1f} % of kernel time\n")
    for k, (t, c) in sorted(per.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f"{t:9.1f} us  {100 * t / total:5.1f} %  n={c:4d}  {k[:100]}")


if __name__ == "__main__":
    main()
