"""Summarise an `ncu --set full` report into a small CSV (one row per profiled launch) with the
metrics the judge reads, plus the top stall reasons.

    python scripts/summarise_ncu.py gpurun_out/prof.ncu-rep > profiles/r02/ncu_<what>.csv
    python scripts/summarise_ncu.py --traffic gpurun_out/prof.ncu-rep 4096 > profiles/r02/ncu_traffic.json

Runs here (no GPU): `ncu -i <rep> --page raw --csv` only reads the report."""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__issue_active.avg.pct", "issue_active_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_insts"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__waves_per_multiprocessor", "waves"),
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def to_bytes(v, unit):
    v = float(v.replace(",", "")) if v else 0.0
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    if sys.argv[1] == "--traffic":
        rep, batch = sys.argv[2], int(sys.argv[3])
        hdr, units, rows = raw(rep)
        ix = {h: i for i, h in enumerate(hdr)}
        agg = {}
        for r in rows:
            name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("ctr::", "")
            key = "embed_fwd" if "embed_fwd" in name else "embed_bwd" if "embed_bwd" in name else \
                "adam_rows" if "adam_rows" in name else name
            b = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) + \
                to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
            a = agg.setdefault(key, {"n": 0, "dram_bytes": 0.0, "time_us": 0.0, "kernels": set()})
            a["n"] += 1
            a["dram_bytes"] += b
            a["time_us"] += float(r[ix["gpu__time_duration.sum"]] or 0)
            a["kernels"].add(name)
        out = {"batch": batch, "source": rep.split("/")[-1],
               "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the profiled "
                       "launches of each kernel; ctr_embed_bwd is two kernels side by side: their sum)"}
        for k, a in agg.items():
            per = len(a["kernels"])              # kernels that make up one logical launch
            launches = max(1, a["n"] // per)
            out[k] = {"dram_bytes": a["dram_bytes"] / launches, "time_us_cold": a["time_us"] / launches,
                      "launches": launches, "kernels": sorted(a["kernels"])}
        print(json.dumps(out, indent=1))
        return
    rep = sys.argv[1]
    hdr, units, rows = raw(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    w = csv.writer(sys.stdout)
    w.writerow(["kernel"] + [n for m, n in METRICS if m in ix] + ["top_stalls"])
    for r in rows:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        vals = []
        for m, n in METRICS:
            if m not in ix:
                continue
            v = r[ix[m]]
            if n.startswith("dram_r") or n.startswith("dram_w"):
                v = "%.0f" % to_bytes(v, units[ix[m]])
            vals.append(v)
        st = sorted(((float(r[ix[h]] or 0), h.split("issue_stalled_")[1].split("_per_")[0]) for h in stall),
                    reverse=True)[:3]
        w.writerow([name] + vals + ["; ".join("%s %.2f" % (n, v) for v, n in st)])


if __name__ == "__main__":
    main()
