#!/usr/bin/env python
"""Summarise an `ncu --set full` report into the small JSON files kept under profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python profiles/summarize_ncu.py /tmp/raw.csv profiles/<name>.json "<how it was captured>"

Per launch: kernel, duration, DRAM bytes read / written, tensor-pipe activity, issue activity, warps active."""
import csv
import json
import sys

KEYS = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__inst_executed.avg": "inst_per_smsp",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__registers_per_thread": "regs",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}


def main():
    raw, out, how = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    launches = []
    for d in data:
        e = {"kernel": d[idx["Kernel Name"]][:110]}
        for k, name in KEYS.items():
            if k not in idx:
                continue
            try:
                v = float(d[idx[k]].replace(",", ""))
            except ValueError:
                continue
            u = units[idx[k]]
            if name.startswith("dram_") and name != "dram_throughput_pct":
                v *= UNIT_SCALE.get(u, 1.0)
                name_out = name + "_bytes"
            elif name == "time_us":
                v *= UNIT_SCALE.get(u, 1.0)
                name_out = name
            else:
                name_out = name
            e[name_out] = v
        launches.append(e)
    fam = {}
    for e in launches:
        base = e["kernel"].split("(")[0].strip()
        f = fam.setdefault(base, {"launches": 0, "time_us": 0.0, "dram_bytes": 0.0})
        f["launches"] += 1
        f["time_us"] += e.get("time_us", 0.0)
        f["dram_bytes"] += e.get("dram_read_bytes", 0.0) + e.get("dram_write_bytes", 0.0)
    json.dump({"how": how, "families": fam, "launches": launches}, open(out, "w"), indent=1)
    print(out, len(launches), "launches")


if __name__ == "__main__":
    main()
