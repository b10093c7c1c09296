"""Survey of the multi-row forms of the row kernels (csr_rowm_kernel / hll_rowm_kernel) in the built library: registers,
spill bytes (csrc/build/*.ptxas.log) and how many of a step's loads ptxas issues ahead of the first multiply (cuobjdump
-sass).  A form whose loads are split by a DMUL stalls on its first gather before the rest is requested.  No GPU needed.

    python tools/rowm_survey.py            # every rowm instantiation and, for comparison, the one-row kernels
"""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "sparsematrixvectormultiplication_b200"


def ptxas_info():
    info, cur = {}, None
    for log in (PKG / "csrc" / "build").glob("cuda_*.ptxas.log"):
        for line in log.read_text().splitlines():
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                cur = m.group(1)
            m = re.search(r"(\d+) bytes spill stores", line)
            if m and cur:
                info.setdefault(cur, {})["spill"] = int(m.group(1))
            m = re.search(r"Used (\d+) registers", line)
            if m and cur:
                info.setdefault(cur, {})["regs"] = int(m.group(1))
    return info


def main():
    info = ptxas_info()
    sass = subprocess.run(["cuobjdump", "-sass", str(PKG / "libspmv_b200.so")], capture_output=True, text=True).stdout
    rows = []
    for body in re.split(r"Function : ", sass)[1:]:
        name = body.split()[0]
        m = re.search(r"(csr|hll)_row(m|u?)_kernelILi(\d)E(?:Li(\d)ELi(\d)E)?([fd])E", name)
        if not m:
            continue
        fmt, multi, batch, rows_per, ctas, v = m.groups()
        ops = re.findall(r"\b(LDG|DMUL)\b", body)
        ahead = 0
        for op in ops:
            if op == "DMUL":
                break
            ahead += 1
        total = ops.count("LDG")
        rec = info.get(name, {})
        rows.append((fmt, v, int(rows_per or 1), int(batch), int(ctas or 8), rec.get("regs"), rec.get("spill"), ahead, total,
                     "row" + multi))
    print("format storage rows batch ctas/SM regs spill_bytes loads_ahead_of_first_DMUL/loads rows_in_flight/SM kernel")
    for r in sorted(rows):
        fmt, v, rp, b, c, regs, spill, ahead, total, kind = r
        print(f"{fmt} {'f32' if v == 'f' else 'f64'} {rp} {b} {c} {regs} {spill} {ahead}/{total} {256 * c * rp} {kind}")


if __name__ == "__main__":
    sys.exit(main())
