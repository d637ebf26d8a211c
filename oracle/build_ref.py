"""
TEST / MEASUREMENT INFRASTRUCTURE -- recipe for `oracle/_ref/`.

The reference (rscalzo/riemann) is 100 % Python, so "building" it means staging an UNMODIFIED copy of its
`riemann/` package and `examples/` scripts (examples/test_changepoint.py holds the proposal BASELINE config 2
uses) where they can travel to the GPU box: `oracle/_ref/` is git-ignored (no reference source enters the
history) but not gpurun-ignored, exactly like the built `.so`.  `__graft_entry__.build()` calls `stage()` in
the build container, where `/root/reference` exists; on the GPU box the staged copy is used as is.

Only `oracle/refshim.py` reads the staged tree, and only `tests/` and `bench.py`'s CPU arm
(`--impl reference`, `cpu_baseline`) import refshim: the CPU arm then times riemann/samplers/sampler.py:44-90
itself (`cpu_baseline.kind == "reference"`).  A manifest with the sha256 of every staged file is written next
to the copy so a run can state which reference it timed.
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DST = os.path.join(HERE, "_ref")
REF_SRC = os.environ.get("RIEMANN_REFERENCE_SRC", "/root/reference")
PARTS = ("riemann", "examples")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def manifest(root):
    out = {}
    for part in PARTS:
        for dp, _, files in os.walk(os.path.join(root, part)):
            for fn in sorted(files):
                if fn.endswith(".py"):
                    p = os.path.join(dp, fn)
                    out[os.path.relpath(p, root)] = _sha(p)
    return out


def stage(force=False):
    """Copy the reference's python sources to oracle/_ref/ (no edits).  Returns the destination, or None when the
    reference tree is absent (GPU box: the copy staged in the build container is already there)."""
    if not os.path.isdir(os.path.join(REF_SRC, "riemann")):
        return REF_DST if os.path.isdir(os.path.join(REF_DST, "riemann")) else None
    want = manifest(REF_SRC)
    mf = os.path.join(REF_DST, "MANIFEST.json")
    if not force and os.path.exists(mf):
        try:
            with open(mf) as f:
                if json.load(f).get("files") == want:
                    return REF_DST
        except Exception:
            pass
    if os.path.isdir(REF_DST):
        shutil.rmtree(REF_DST)
    os.makedirs(REF_DST)
    for part in PARTS:
        shutil.copytree(os.path.join(REF_SRC, part), os.path.join(REF_DST, part),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(mf, "w") as f:
        json.dump({"source": REF_SRC, "files": want}, f, indent=1, sort_keys=True)
    return REF_DST


if __name__ == "__main__":
    print(stage(force=True))
