#!/usr/bin/env python
"""Summarises an .ncu-rep (ncu --set full) into the few metrics the roofline in DESIGN.md / bench.py cites.

  python tools/ncu_summary.py gpurun_out/prof_pair_r01.ncu-rep profiles/r01_ncu_conv_pair.txt [--json out.json]
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "gpc__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.max",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "l1tex__m_l1tex2xbar_write_bytes.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines, summary = [], []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines.append(f"kernel: {d.get('Kernel Name')}")
        rec = {"kernel": d.get("Kernel Name")}
        for k in KEYS:
            if k in d and d[k] != "":
                lines.append(f"  {k} [{u[k]}] = {d[k]}")
                rec[k] = (d[k], u[k])
        try:
            rd = float(d["dram__bytes_read.sum"].replace(",", "")) * SCALE[u["dram__bytes_read.sum"]]
            wr = float(d["dram__bytes_write.sum"].replace(",", "")) * SCALE[u["dram__bytes_write.sum"]]
            rec["dram_bytes_total"] = rd + wr
            lines.append(f"  => DRAM traffic per launch = {(rd + wr) / 1e6:.1f} MB")
        except Exception:  # noqa: BLE001
            pass
        summary.append(rec)
        lines.append("")
    with open(out, "w") as fh:
        fh.write(f"# ncu --set full --clock-control none summary of {rep}\n" + "\n".join(lines))
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as fh:
            json.dump(summary, fh, indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
