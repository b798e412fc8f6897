"""Turn `ncu -i X.ncu-rep --page raw --csv` exports into the small tracked summaries under
profiles/: one CSV row per captured kernel (selected metrics) and r01_ncu_traffic.json with the
DRAM bytes per launch that bench.py reports as roofline.traffic."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v.replace(",", "")) * m.get(unit, 1)


def main(rep, out_csv, batch=None, traffic_json=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    stall = [k for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")]
    cols = ["Kernel Name"] + [k for k in KEYS if k in hdr]
    traffic = {}
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols + ["units", "top_stalls"])
        for r in body:
            vals = [r[hdr.index(c)] for c in cols]
            st = sorted(((float(r[hdr.index(k)]), k.split("stalled_")[1].replace("_per_issue_active.ratio", ""))
                         for k in stall), reverse=True)[:4]
            w.writerow(vals + [" ".join(units[hdr.index(c)] or "-" for c in cols[1:]),
                               "; ".join("%s %.1f" % (n, v) for v, n in st)])
            name = r[hdr.index("Kernel Name")]
            rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
            wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
            for key in ("embed_fwd", "embed_bwd", "adam_rows", "tower_mid", "tc_gemm"):
                if key in name and key not in traffic:
                    traffic[key] = {"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr,
                                    "duration_us": float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")),
                                    "kernel": name[:80]}
    if traffic_json:
        traffic["batch"] = int(batch)
        traffic["source"] = os.path.basename(rep) + " (ncu --set full --clock-control none; caches flushed per kernel)"
        json.dump(traffic, open(traffic_json, "w"), indent=1)
    print("wrote", out_csv, traffic_json or "")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], *(sys.argv[3:5]))
