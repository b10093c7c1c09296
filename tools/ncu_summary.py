#!/usr/bin/env python
"""Turn an `ncu --set full` report into the markdown table kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--pick N] [--label name=workload ...] > profiles/rNN_ncu.md

Reads the report with `ncu -i <rep> --page raw --csv` (works without a GPU).  One column per profiled launch;
--pick N keeps every N-th launch per kernel name (the capture profiles REPS launches of each workload)."""
import argparse
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 (lts) throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit rate %"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 sectors read by L1TEX"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("smsp__cycles_active.avg", "SMSP active cycles (avg)"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--pick", type=int, default=1, help="keep the last of every N consecutive launches")
    ap.add_argument("--select", default="", help="comma separated launch indices to keep (applied before --pick)")
    ap.add_argument("--labels", nargs="*", default=[], help="column labels, in launch order after --pick")
    ap.add_argument("--traffic-json", default="", help="update this JSON file (profiles/ncu_traffic.json) with the DRAM bytes per "
                    "launch of every kept launch, keyed '<workload>|<kernel>'; bench.py reads roofline.traffic from it")
    ap.add_argument("--workload", default="", help="workload name for --traffic-json (bench.py's config.workload)")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    if args.select:
        data = [data[int(i)] for i in args.select.split(",")]
    if args.pick > 1:
        data = [r for i, r in enumerate(data) if i % args.pick == args.pick - 1]
    name_i = hdr.index("Kernel Name")
    cols = []
    for i, r in enumerate(data):
        kname = r[name_i].split("(")[0].replace("void ", "")
        label = args.labels[i] if i < len(args.labels) else f"launch {i}"
        cols.append(f"{label}: `{kname}`")
    print("| metric | " + " | ".join(cols) + " |")
    print("|---|" + "---|" * len(cols))
    for key, title in METRICS:
        if key not in hdr:
            continue
        j = hdr.index(key)
        unit = units[j]
        cells = []
        for r in data:
            v = r[j]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.4g}" if abs(f) < 1e6 else f"{f:.6g}"
            except ValueError:
                pass
            cells.append(v)
        print(f"| {title} ({key}{', ' + unit if unit else ''}) | " + " | ".join(cells) + " |")
    # derived: DRAM traffic per launch in bytes
    if "dram__bytes_read.sum" in hdr and "dram__bytes_write.sum" in hdr:
        jr, jw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        cells = []
        for r in data:
            t = float(r[jr]) * scale.get(units[jr], 1.0) + float(r[jw]) * scale.get(units[jw], 1.0)
            cells.append(f"{t:.6g}")
        print("| **DRAM traffic per launch (read + written, bytes)** | " + " | ".join(cells) + " |")
        if args.traffic_json and args.workload:
            import json
            import re
            from pathlib import Path
            path = Path(args.traffic_json)
            table = json.loads(path.read_text()) if path.exists() else {}
            for r, cell in zip(data, cells):
                kernel = r[name_i].split("(")[0] if "<" not in r[name_i] else r[name_i][: r[name_i].rfind(">") + 1]
                kernel = re.sub(r"\(int\)|void |spmv::|\s", "", kernel)
                table[f"{args.workload}|{kernel}"] = {"dram_bytes": int(float(cell)), "source": str(Path(args.report).name)}
            path.write_text(json.dumps(table, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    sys.exit(main())
