"""CPU tests of the measurement tools under tools/ and of what the shipped library contains (cuobjdump runs without a GPU):
the tensor-core / TMA / TMEM / 256-bit-store claims of DESIGN.md are checked against the SASS of libganb200.so itself."""
import csv
import json
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gan_lib_tensorflow_b200", "libganb200.so")


def _run(tool, *args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), *args], capture_output=True, text=True,
                          check=True).stdout


@pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="needs cuobjdump and the built library")
def test_shipped_library_uses_tcgen05_tma_tmem_and_256_bit_stores():
    out = _run("sass_summary.py", LIB)
    totals = next(line for line in out.splitlines() if line.startswith("library totals:"))
    counts = dict(tok.split("=") for tok in totals.split(":", 1)[1].split())
    for mnemonic in ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "STG.256", "LDG.256"):
        assert int(counts.get(mnemonic, 0)) > 0, (mnemonic, totals)
    # the dominant kernel: CTA-pair MMA fed by TMA, accumulators read from TMEM, 32-byte epilogue stores
    pair = next(line for line in out.splitlines() if line.startswith("conv_pair_kernel<256, 3, 8, false>"))
    for mnemonic in ("UTCHMMA.2CTA", "UTMALDG", "LDTM", "STG.256"):
        assert mnemonic + "=" in pair, pair
    assert "HMMA" not in counts                      # no mma.sync / wmma path anywhere in the library


def test_phase_times_on_a_synthetic_timeline(tmp_path):
    ev, t = [], 0.0

    def k(name, dur):
        nonlocal t
        ev.append({"name": name, "start": t, "dur": dur, "stream": 7})
        t += dur

    for _pair in range(3):
        k("ganb::preprocess_real_kernel", 4)
        k("ganb::conv_igemm_kernel<64, 8>", 96)
        k("ganb::gan_loss_kernel", 2)              # critic forward of the D step: 100 us
        k("ganb::conv_wgrad_kernel<128, 6>", 190)
        k("ganb::adam_kernel", 8)                   # critic backward + update: 200 us
        k("ganb::conv_igemm_kernel<64, 8>", 50)
        k("ganb::gan_loss_kernel", 2)              # critic forward of the G step: 50 us ... to the second loss
        k("ganb::conv_igemm_kernel<64, 8>", 68)
        k("ganb::conv_halo_narrow_kernel<16, 4>", 30)
        k("ganb::conv_pair_kernel<256, 3, 8>", 270)
        k("ganb::adam_kernel", 30)
        k("ganb::conv_pair_kernel<256, 3, 8>", 250)  # both generator forwards of the next pair
    path = tmp_path / "timeline.json"
    path.write_text(json.dumps(ev))
    out = _run("phase_times.py", str(path))
    rows = {line.split("us", 1)[1].strip(): float(line.split("us", 1)[0]) for line in out.splitlines() if " us " in line}
    assert rows["critic fwd (D step)"] == pytest.approx(100.0)
    assert rows["critic bwd + update"] == pytest.approx(200.0)
    assert rows["critic fwd (G step)"] == pytest.approx(50.0)
    assert rows["critic bwd (G step)"] == pytest.approx(70.0)
    assert rows["generator bwd + update"] == pytest.approx(330.0)
    assert rows["period"] == pytest.approx(1000.0)


def test_ncu_launch_summary_shares(tmp_path):
    path = tmp_path / "launches.csv"
    with open(path, "w", newline="") as fh:
        fh.write("==PROF== Connected to process 1\n")
        w = csv.writer(fh, quoting=csv.QUOTE_ALL)
        w.writerow(["ID", "Process ID", "Process Name", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"])
        w.writerow(["0", "1", "python", "void ganb::conv_pair_kernel<256, 3, 8, 0>(CUtensorMap_st)", "gpu__time_duration.sum", "us", "60.0"])
        w.writerow(["1", "1", "python", "ganb::bn_stats_partial_v8p_kernel(const __nv_bfloat16 *)", "gpu__time_duration.sum", "us", "30.0"])
        w.writerow(["2", "1", "python", "void at::native::vectorized_elementwise_kernel<4>(int)", "gpu__time_duration.sum", "ns", "10000"])
    out = _run("ncu_launch_summary.py", str(path))
    assert "total 100 us over 3 launches" in out
    assert "libganb200 kernels 90.0 %" in out and "tensor-core convolution kernels 60.0 %" in out
