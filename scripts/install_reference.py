#!/usr/bin/env python
"""Install the UNMODIFIED reference into baseline/_ref (git-ignored; it travels to the GPU box with gpurun snapshots).

    python scripts/install_reference.py [--reference /root/reference]

The reference's pyproject declares `packages.find where=["src"]`, so a plain `pip install --target baseline/_ref` lands its
sub-packages as top-level `data/ models/ pipelines/ ...` while its own code imports `src.data`, `src.models`, ... - unusable.
Installing with `--target baseline/_ref/src` gives the layout its imports expect (`src` becomes a namespace package) with
`baseline/_ref` on sys.path.  The build needs a writable source tree, so pip runs on a copy under /tmp.  `--no-deps`: faiss-cpu
and matplotlib are not installable here (no network); both are optional at run time (training.py:29-32 degrades faiss to None,
matplotlib is stubbed by scripts/train_b200.py / tests/refenv.py).
Next to the package, the reference's `tests/`, `configs/` and `scripts/` directories (not part of its wheel) are copied verbatim
so that its own test files and `scripts/train.py` can be run against the B200 drop-in on the GPU box.
Nothing under baseline/_ref is product source, and nothing in the product package imports it.
"""
from __future__ import annotations

import argparse
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
DEST = ROOT / "baseline" / "_ref"


def install(reference: Path, dest: Path = DEST) -> bool:
    if not (reference / "pyproject.toml").exists():
        return False
    with tempfile.TemporaryDirectory() as tmp:
        src = Path(tmp) / "ref"
        shutil.copytree(reference, src)
        if dest.exists():
            shutil.rmtree(dest)
        (dest / "src").mkdir(parents=True)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", str(dest / "src"), str(src)]
        subprocess.run(cmd, check=True)
    for extra in ("tests", "configs", "scripts"):
        if (reference / extra).is_dir():
            shutil.copytree(reference / extra, dest / extra)
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    a = ap.parse_args()
    ok = install(Path(a.reference))
    print(f"baseline/_ref {'installed' if ok else 'NOT installed (reference absent)'}")
