"""Regenerate every fixture of tests/golden/ from the UNMODIFIED reference in a scratch copy of the repo and compare the arrays
byte for byte with the committed files (build container only: needs /root/reference).

    python tools/check_golden.py        # prints one line per fixture; exit code 1 if any array differs
"""
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent


def same_bytes(a: np.ndarray, b: np.ndarray) -> bool:
    return a.shape == b.shape and a.dtype == b.dtype and np.ascontiguousarray(a).tobytes() == np.ascontiguousarray(b).tobytes()


def main() -> int:
    with tempfile.TemporaryDirectory() as tmp:
        scratch = Path(tmp) / "repo"
        shutil.copytree(REPO, scratch, ignore=shutil.ignore_patterns(".git", "gpurun_out", "profiles", "build", "*.so", "_ref"))
        for f in (scratch / "tests" / "golden").glob("*.npz"):
            f.unlink()                                      # nothing may survive from the committed set
        subprocess.run([sys.executable, "tools/make_golden.py"], cwd=scratch, check=True, capture_output=True)
        bad = 0
        for committed in sorted((REPO / "tests" / "golden").glob("*.npz")):
            fresh = scratch / "tests" / "golden" / committed.name
            if not fresh.exists():
                print(f"{committed.name}: NOT regenerated")
                bad += 1
                continue
            a, b = np.load(committed), np.load(fresh)
            diff = sorted(k for k in set(a.files) | set(b.files) if k not in a.files or k not in b.files or not same_bytes(a[k], b[k]))
            print(f"{committed.name}: {len(a.files)} arrays, " + ("byte-identical" if not diff else f"DIFFERENT in {diff[:8]}"))
            bad += bool(diff)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
