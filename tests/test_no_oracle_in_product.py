"""The oracle is the checker.  The product package must never import, link or execute it."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "sparsematrixvectormultiplication_b200"


def test_product_sources_never_touch_the_oracle():
    pat = re.compile(r"liboracle|libspmv_ref|from\s+oracle|import\s+oracle|oracle/|orc_\w+")
    offenders = []
    for path in PKG.rglob("*"):
        if path.suffix in {".py", ".c", ".cu", ".cuh", ".h"} or path.name == "Makefile":
            for n, line in enumerate(path.read_text(errors="replace").splitlines(), 1):
                code = line.split("#")[0] if path.suffix == ".py" else line
                if pat.search(code) and "oracle can check" not in line and "CPU oracle" not in line and "oracle's layout" not in line:
                    offenders.append(f"{path.relative_to(ROOT)}:{n}: {line.strip()}")
    assert not offenders, "\n".join(offenders)


def test_include_headers_never_touch_the_oracle():
    for path in (ROOT / "include").glob("*.h"):
        assert "orc_" not in path.read_text() and "oracle.h" not in path.read_text()
