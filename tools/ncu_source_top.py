#!/usr/bin/env python
"""Top stall lines of one kernel from an ncu report's source page (needs -lineinfo + --import-source on).
    python tools/ncu_source_top.py report.ncu-rep kernel_regex [N]"""
import csv
import io
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                         capture_output=True, text=True).stdout
    blocks = raw.split('"Kernel Name",')
    if len(blocks) < 2:
        raise SystemExit("no kernel matched")
    body = blocks[1].split("\n", 1)[1]
    rows = list(csv.reader(io.StringIO(body)))
    hdr = rows[0]
    data = [r for r in rows[1:] if len(r) == len(hdr)]
    si = hdr.index("# Samples")
    src = hdr.index("Source")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(r[si] or 0) for r in data)
    print(f"{blocks[1].splitlines()[0][:100]}  total samples {total}")
    data.sort(key=lambda r: -int(r[si] or 0))
    for r in data[:top]:
        n = int(r[si] or 0)
        stalls = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:3]
        s = ", ".join(f"{name} {v}" for v, name in stalls if v)
        print(f"{100.0*n/total:5.1f}%  {r[src][:110]:110s} | {s}")


if __name__ == "__main__":
    main()
