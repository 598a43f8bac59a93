#!/usr/bin/env python
"""Place the UNMODIFIED reference (THUwangcy/ReChorus) under baseline/_ref/ so that it travels to the GPU box
(git-ignored, not gpurun-ignored): the reference arm of bench.py (`--impl reference`, cpu_baseline kind "reference") and
the overlay tests run the reference's own classes there.  `/root/reference` does not exist on the GPU box.

The contract's recipe is `pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref
/root/reference`; ReChorus is a source tree without setup.py / pyproject.toml ("Directory is not installable"), so the
fallback is a verbatim copy of its `src/` (580 KB of Python) and of the one csv dataset it ships.  Nothing is edited; a
manifest with per-file SHA-256 is written next to it so "unmodified" can be checked.  Run by __graft_entry__.build()
whenever /root/reference is present."""
import hashlib
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("B2R_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def install(verbose=True) -> str:
    if not os.path.isdir(os.path.join(SRC, "src")):
        return ""
    note = ""
    if not os.path.isdir(os.path.join(DST, "src")):
        os.makedirs(DST, exist_ok=True)
        res = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links",
                              "/opt/wheelhouse", "--target", DST, SRC], capture_output=True, text=True)
        note = "pip: " + (res.stderr.strip().splitlines() or ["ok"])[-1]
        if res.returncode != 0:
            shutil.copytree(os.path.join(SRC, "src"), os.path.join(DST, "src"),
                            ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
            data = os.path.join(SRC, "data", "Grocery_and_Gourmet_Food")
            if os.path.isdir(data):
                os.makedirs(os.path.join(DST, "data", "Grocery_and_Gourmet_Food"), exist_ok=True)
                for f in ("train.csv", "dev.csv", "test.csv"):
                    shutil.copy2(os.path.join(data, f), os.path.join(DST, "data", "Grocery_and_Gourmet_Food", f))
            for f in ("LICENSE", "requirements.txt"):
                if os.path.exists(os.path.join(SRC, f)):
                    shutil.copy2(os.path.join(SRC, f), os.path.join(DST, f))
    manifest = {}
    for base, _, files in os.walk(os.path.join(DST, "src")):
        for f in sorted(files):
            if f.endswith(".py"):
                p = os.path.join(base, f)
                rel = os.path.relpath(p, DST)
                manifest[rel] = _sha(p)
                ref = os.path.join(SRC, rel)
                if os.path.exists(ref) and _sha(ref) != manifest[rel]:
                    raise RuntimeError(f"baseline/_ref/{rel} differs from the reference checkout")
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "how": note or "already present", "files": manifest}, f, indent=1)
    if verbose:
        print(f"baseline/_ref: {len(manifest)} reference files ({note or 'already present'})")
    return DST


if __name__ == "__main__":
    install()
