#!/usr/bin/env python3
"""Turn an ncu report (ncu --set full ... -o X) into the per-kernel table kept under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof_r1a.ncu-rep > profiles/r1a_ncu_full_summary.md
"""
import csv
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def main(path):
    raw = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print("| # | kernel | " + " | ".join(n for _, n in COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    for n, r in enumerate(rows[2:]):
        cells = []
        for m, _ in COLS:
            if m in hdr:
                i = hdr.index(m)
                v = r[i]
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {units[i]}".strip())
            else:
                cells.append("-")
        print(f"| {n} | `{r[ki][:48]}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
