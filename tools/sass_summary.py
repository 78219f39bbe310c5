#!/usr/bin/env python
"""Instruction-count summary of the shipped library's SASS (cuobjdump -sass): which kernels use the 5th-generation
tensor cores (UTCHMMA / UTCHMMA.2CTA = tcgen05.mma, cta_group::2), TMEM loads (LDTM), TMA loads / stores
(UTMALDG / UTMASTG), multicast commits (UTCBAR).  Runs without a GPU.

  python tools/sass_summary.py [path/to/libganb200.so] > profiles/rNN_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCCP", "UBLKCP",
             "SYNCS", "ELECT", "LDGSTS", "HMMA", "REDUX", "SHFL", "ATOM", "RED"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gan_lib_tensorflow_b200", "libganb200.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        per[cur]["_total"] += 1
        if op.startswith("STG.") and ".256" in op:      # 256-bit stores: one whole 32-byte sector per thread and instruction
            per[cur]["STG.256"] += 1
        if op.startswith("LDG.") and ".256" in op:
            per[cur]["LDG.256"] += 1
        for mn in MNEMONICS:
            if op == mn or op.startswith(mn + "."):
                if mn == "UTCHMMA" and op.startswith("UTCHMMA.2CTA"):
                    continue
                per[cur][mn] += 1
                break
    demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    tot = collections.Counter()
    print(f"cuobjdump -sass {os.path.relpath(lib, ROOT)}: {len(per)} kernels for sm_100a")
    print("%-86s %7s  %s" % ("kernel", "instrs", "tensor / TMA / TMEM instructions"))
    for (name, cnt), dm in zip(per.items(), demangle):
        short = re.sub(r"\(.*", "", dm).replace("void ", "").replace("ganb::", "")[:86]
        hits = " ".join(f"{k}={v}" for k, v in cnt.items() if k != "_total" and k in
                        ("UTCHMMA.2CTA", "UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "UBLKCP", "LDGSTS", "STG.256", "LDG.256"))
        for k, v in cnt.items():
            tot[k] += v
        if hits:
            print("%-86s %7d  %s" % (short, cnt["_total"], hits))
    print("\nlibrary totals: " + "  ".join(f"{k}={tot[k]}" for k in MNEMONICS + ["STG.256", "LDG.256"] if tot[k]))
    print("kernels without tensor-core / TMA instructions (CUDA-core, bandwidth-bound): %d" %
          sum(1 for c in per.values() if not any(c[k] for k in ("UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "LDTM"))))


if __name__ == "__main__":
    main()
