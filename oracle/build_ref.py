"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference package as a build artefact.

    python -m oracle.build_ref          # build container only (needs /root/reference)

The reference is pure Python (plus the two source files of its upfirdn2d op, which torch JIT-compiles on first
import where CUDA is available).  ``/root/reference`` does not exist on the GPU box, so ``build()`` stages a
byte-for-byte copy of ``sgmse/`` and ``enhancement.py`` under ``oracle/_ref/`` -- git-ignored (the reference's
sources never enter this repository's history) but not gpurun-ignored, i.e. it travels with the snapshot exactly
like ``sgmse_b200/lib/*.so``.  A manifest with the sha256 of every staged file is written next to it, so a test
can prove the checker is the unmodified reference (``tests/test_oracle_vs_reference.py::test_staged_reference_is_unmodified``).

Consumers (test infrastructure only, see oracle/__init__.py): ``oracle/refshim.py`` imports the package from
``/root/reference`` when it exists and from ``oracle/_ref`` otherwise; ``bench.py --impl reference`` / ``cpu_baseline``
time the reference's own ``enhancement.py:75-96`` sequence on it (``kind: "reference"``); the ``-m gpu`` drop-in
tests build a live ``ScoreModel`` from it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SGMSE_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
# what the hot path and its callers import; preprocessing/, train.py, calc_metrics.py, logs/ are not needed
TREES = ["sgmse"]
FILES = ["enhancement.py", "LICENSE"]
KEEP_EXT = (".py", ".cpp", ".cu", ".h")


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def staged() -> bool:
    return os.path.isfile(os.path.join(DST, "MANIFEST.json")) and os.path.isdir(os.path.join(DST, "sgmse"))


def build(force: bool = False) -> str | None:
    """Stage the reference under oracle/_ref/ (no-op where /root/reference is absent: the GPU box uses the staged copy)."""
    if not os.path.isdir(os.path.join(SRC, "sgmse")):
        return DST if staged() else None
    files = []
    for tree in TREES:
        for dirpath, dirnames, filenames in os.walk(os.path.join(SRC, tree)):
            dirnames[:] = [d for d in dirnames if d != "__pycache__"]
            for f in filenames:
                if f.endswith(KEEP_EXT):
                    files.append(os.path.relpath(os.path.join(dirpath, f), SRC))
    files += [f for f in FILES if os.path.isfile(os.path.join(SRC, f))]
    manifest = {"source": SRC, "files": {}}
    sub = os.path.join(SRC, ".SUBMODULES.json")
    if os.path.isfile(sub):
        try:
            manifest["submodules"] = json.load(open(sub))
        except Exception:
            pass
    for rel in sorted(files):
        manifest["files"][rel] = _sha(os.path.join(SRC, rel))
    mpath = os.path.join(DST, "MANIFEST.json")
    if not force and os.path.isfile(mpath):
        try:
            old = json.load(open(mpath))
            if old.get("files") == manifest["files"] and all(
                    os.path.isfile(os.path.join(DST, r)) and _sha(os.path.join(DST, r)) == s for r, s in old["files"].items()):
                return DST
        except Exception:
            pass
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for rel in files:
        out = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), out)
    with open(mpath, "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    return DST


def verify() -> list:
    """Names of staged files whose bytes differ from the manifest (empty = the staged copy is intact)."""
    m = json.load(open(os.path.join(DST, "MANIFEST.json")))
    return [r for r, s in m["files"].items() if not os.path.isfile(os.path.join(DST, r)) or _sha(os.path.join(DST, r)) != s]


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
