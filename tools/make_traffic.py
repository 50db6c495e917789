"""profiles/traffic.json from an `ncu --set full` capture of one step (tools/prof_step.py): DRAM bytes (dram__bytes_read.sum +
dram__bytes_write.sum) per kernel launch, keyed the way bench.py looks them up ("<precision>_<streams>streams_<stems>stems").
    python tools/make_traffic.py gpurun_out/r2_full.ncu-rep compensated 32 4 "profiles/r2_ncu_full_summary.md"
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(rep, precision, streams, stems, source):
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    ki, ri, wi, ti = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
    units = rows[1]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}

    def val(r, i):
        return float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
    launches = [(r[ki], val(r, ri) + val(r, wi), float(r[ti].replace(",", ""))) for r in rows[2:]]
    names = ["meta", "stft", "down1", "down2", "down3", "down4", "down5", "down6", "up1", "up2", "up3", "up4", "up5", "up6"]
    per = {}
    k = 0
    for name, b, _ in launches:
        if "up7" in name:
            per["up7"] = per.get("up7", 0.0) + b
        elif "istft" in name:
            per["istft"] = per.get("istft", 0.0) + b
        elif "conv_rp" in name and "8, 16>" in name.replace(" ", "") and "down1" in per:
            per["down1"] += b                      # more than one down1 launch (stems in groups of 4)
        else:
            per[names[k]] = per.get(names[k], 0.0) + b
            k += 1
    per["up6+up7"] = per.get("up6", 0.0) + per.get("up7", 0.0)
    tc = sum(per[n] for n in ("down2", "down3", "down4", "down5", "down6", "up1", "up2", "up3", "up4", "up5"))
    path = os.path.join(ROOT, "profiles", "traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[f"{precision}_{streams}streams_{stems}stems"] = {"source": source, "tc_layers_dram_bytes_per_step": tc, "per_kernel_dram_bytes": per,
                                                         "step_dram_bytes": sum(b for _, b, _ in launches)}
    json.dump(data, open(path, "w"), indent=1)
    print(json.dumps(data[f"{precision}_{streams}streams_{stems}stems"], indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5])
