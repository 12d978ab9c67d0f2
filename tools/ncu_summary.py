"""Summarise an ncu report (read on the CPU box): one row per launch with the metrics the
judge asks for.  usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/x.csv"""
import csv
import subprocess
import sys

WANT = [
    "ID", "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in WANT if c in idx]
    w = csv.writer(sys.stdout)
    w.writerow([f"{c} [{units[idx[c]]}]" if units[idx[c]] else c for c in cols])
    for r in data:
        row = [r[idx[c]] for c in cols]
        row[1] = row[1].split("(")[0][-40:]
        w.writerow(row)


if __name__ == "__main__":
    main()
