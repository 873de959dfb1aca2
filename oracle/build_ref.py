"""Make `oracle/_ref/sdnq`: a pristine, byte-for-byte copy of the reference's Python package (`/root/reference/src/sdnq`).

    python oracle/build_ref.py

The reference is pure Python (no compiled sources), so "building" it is copying its package where the GPU box can import it:
`/root/reference` exists only in the authoring container, `oracle/_ref/` is git-ignored (reference sources never enter this
repository's history) but not gpurun-ignored, so it travels with the snapshot like the built `.so` files do.  Used by the CPU
arm of bench.py (`kind: "reference"`), its `gpu_reference` leg and the level-B binding test -- never by the product path.
A manifest with the SHA-256 of every copied file is written next to it so that a run can show the copy is unmodified.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/sdnq"
DST = os.path.join(HERE, "_ref", "sdnq")


def build(verbose=False):
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None       # GPU box: use the copy that travelled
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for dirpath, dirnames, filenames in os.walk(SRC):
        dirnames[:] = [d for d in dirnames if d != "__pycache__"]
        rel = os.path.relpath(dirpath, SRC)
        os.makedirs(os.path.join(DST, rel), exist_ok=True)
        for f in filenames:
            if not f.endswith(".py"):
                continue
            s, d = os.path.join(dirpath, f), os.path.join(DST, rel, f)
            shutil.copyfile(s, d)
            manifest[os.path.normpath(os.path.join(rel, f))] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    with open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    if verbose:
        print(f"copied {len(manifest)} files -> {DST}")
    return DST


if __name__ == "__main__":
    out = build(verbose=True)
    sys.exit(0 if out else 1)
