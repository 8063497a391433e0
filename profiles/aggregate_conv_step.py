#!/usr/bin/env python
"""Label and aggregate the 43 `conv_gemm_kernel` launches of ONE train step from an `ncu --set full` raw CSV.

    ncu --set full --clock-control none -k regex:conv_gemm_kernel -s 172 -c 43 -o /tmp/x \
        python bench.py --steps 2 --warmup 3 --no-latency --no-cpu-baseline --no-side-configs --no-graph
    ncu -i /tmp/x.ncu-rep --page raw --csv > gpurun_out/full_raw.csv
    python profiles/aggregate_conv_step.py gpurun_out/full_raw.csv gpurun_out/per_op_b256.json profiles/<out>.json "<how>"

Launch order of a step (net.cu forward_impl / backward): 18 conv3x3 forward launches with the four ConvTranspose2d
forwards after bottleneck.3 / dec4.3 / dec3.3 / dec2.3, then the data gradients from dec1.3 down to enc1.3.
`families.conv_fprop` is what bench.py reads for `roofline.traffic`."""
import csv
import json
import sys

sys.path.insert(0, __file__.rsplit("/", 1)[0])
from summarize_ncu import KEYS, UNIT_SCALE  # noqa: E402

CONV = ["enc1.0", "enc1.3", "enc2.0", "enc2.3", "enc3.0", "enc3.3", "enc4.0", "enc4.3", "bott.0", "bott.3",
        "dec4.0", "dec4.3", "dec3.0", "dec3.3", "dec2.0", "dec2.3", "dec1.0", "dec1.3"]
FWD = CONV[:10] + ["up4"] + CONV[10:12] + ["up3"] + CONV[12:14] + ["up2"] + CONV[14:16] + ["up1"] + CONV[16:18]
BWD = ["dec1.3", "dec1.0", "up1", "dec2.3", "dec2.0", "up2", "dec3.3", "dec3.0", "up3", "dec4.3", "dec4.0", "up4",
       "bott.3", "bott.0", "enc4.3", "enc4.0", "enc3.3", "enc3.0", "enc2.3", "enc2.0", "enc1.3"]
EXTRA = {
    "derived__lts__lts2xbar_bytes.sum.per_second": "l2_to_sm_TBps",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
}


def main():
    raw, perop, out, how = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    assert len(data) == 43, len(data)
    labels = [n + " fprop" for n in FWD] + [n + " dgrad" for n in BWD]
    launches = []
    for d, lab in zip(data, labels):
        e = {"layer": lab, "kernel": d[idx["Kernel Name"]].split("(")[0].replace("void ", "")}
        for k, name in list(KEYS.items()) + list(EXTRA.items()):
            if k not in idx:
                continue
            try:
                v = float(d[idx[k]].replace(",", ""))
            except ValueError:
                continue
            u = units[idx[k]]
            if name in ("dram_read", "dram_write"):
                e[name + "_bytes"] = v * UNIT_SCALE.get(u, 1.0)
            elif name == "time_us":
                e[name] = v * UNIT_SCALE.get(u, 1.0)
            else:
                e[name] = v
        launches.append(e)
    # algorithmic bytes of the same ops (SURVEY 8d figures as the per-op table of bench.py carries them)
    po = json.load(open(perop))
    alg = {}
    for r in po["rows"]:
        alg.setdefault(r["name"], 0.0)
        alg[r["name"]] += r["bytes"] / po["prof_steps"]

    def fam(sel, algo):
        ls = [e for e in launches if sel(e["layer"])]
        t = sum(e["time_us"] for e in ls)
        return {"launches": len(ls), "time_us": t,
                "dram_bytes": sum(e["dram_read_bytes"] + e["dram_write_bytes"] for e in ls),
                "tensor_pipe_active_pct_time_weighted": sum(e.get("tensor_pipe_active_pct", 0.0) * e["time_us"] for e in ls) / t,
                "algorithmic_bytes": algo}

    families = {
        "conv_fprop": fam(lambda l: l.endswith("fprop") and not l.startswith("up"), alg.get("conv_fprop", 0.0)),
        "forward_conv(fprop+convT_fprop)": fam(lambda l: l.endswith("fprop"), alg.get("conv_fprop", 0.0) + alg.get("convT_fprop", 0.0)),
        "backward_conv(dgrad+convT_dgrad)": fam(lambda l: l.endswith("dgrad"), alg.get("conv_dgrad", 0.0) + alg.get("convT_dgrad", 0.0)),
    }
    json.dump({"how": how, "families": families, "launches": launches}, open(out, "w"), indent=1)
    for k, v in families.items():
        print(k, {a: round(b, 1) if isinstance(b, float) else b for a, b in v.items()})


if __name__ == "__main__":
    main()
