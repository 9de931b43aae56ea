"""Markdown table of the per-kernel headline metrics of an `ncu --set full` report.
    python tools/ncu_summary.py report.ncu-rep > profiles/<name>.md      (needs ncu on PATH; reads, never profiles)"""
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time", 1.0),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1.0),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor (DMMA) pipe %", 1.0),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 CUDA-core %", 1.0),
    ("dram__bytes_read.sum", "DRAM rd", 1.0),
    ("dram__bytes_write.sum", "DRAM wr", 1.0),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1.0),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", 1.0),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %", 1.0),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %", 1.0),
    ("launch__registers_per_thread", "regs", 1.0),
    ("launch__grid_size", "grid", 1.0),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics",
                          ",".join(m for m, _, _ in METRICS)], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {name: hdr.index(name) for name, _, _ in METRICS if name in hdr}
    ki = hdr.index("Kernel Name")
    print("| # | kernel | " + " | ".join(f"{label} [{units[col[m]]}]" if units[col[m]] else label
                                        for m, label, _ in METRICS if m in col) + " |")
    print("|---|---|" + "---|" * len(col))
    for i, r in enumerate(rows[2:]):
        name = r[ki].replace("oo::", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("void ", "")
        name = name.split("(")[0]
        vals = []
        for m, _, _ in METRICS:
            if m not in col:
                continue
            v = r[col[m]]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.3g}" if abs(f) < 1e4 else f"{f:.0f}"
            except ValueError:
                pass
            vals.append(v)
        print(f"| {i} | `{name}` | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
