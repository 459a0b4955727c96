"""Recipe that stages the UNMODIFIED reference modules for the CPU baseline -- TEST / BENCH INFRASTRUCTURE.

    python oracle/stage_ref.py            # /root/reference/src/{models,utils}/*.py  ->  oracle/_ref/src/

The reference (mahdi-shafiei/AIMNet-X2D) is pure Python with no build system (no setup.py / pyproject.toml), so there is
nothing to compile or pip-install: the files the hot path needs are copied byte for byte from where they lie under
``/root/reference`` into ``oracle/_ref/`` -- which is listed in ``.gitignore`` (never part of the history) but NOT in
``.gpurunignore``, so it travels to the GPU box with the snapshot, where ``/root/reference`` does not exist.
``bench.py --impl reference`` (and its ``cpu_baseline`` leg) then import ``models.gnn.GNN`` / ``models.losses`` from there
and drive the reference's own ``nn.Module`` API; ``torch_scatter`` (third-party, pinned 2.1.2, not installable here) is
provided by ``oracle/scatter_port.py``.  A manifest with the sha256 of every staged file is written next to them;
``verify()`` re-checks it so that a modified copy is never timed as "the reference".
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
DST = os.path.join(HERE, "_ref")
FILES = ["models/__init__.py", "models/gnn.py", "models/layers.py", "models/pooling.py", "models/losses.py",
         "utils/__init__.py", "utils/activation.py", "utils/distributed.py", "utils/optimization.py", "utils/random.py"]


def _sha(path: str) -> str:
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def stage() -> bool:
    """Copy the files; returns False (and leaves any existing staging alone) when /root/reference is absent."""
    if not os.path.isdir(REF_SRC):
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(DST, "src", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF_SRC, "sha256": manifest}, fh, indent=1)
    return True


def verify() -> bool:
    """True iff oracle/_ref holds every staged file with the recorded checksum."""
    path = os.path.join(DST, "MANIFEST.json")
    if not os.path.exists(path):
        return False
    with open(path) as fh:
        manifest = json.load(fh)["sha256"]
    return all(os.path.exists(os.path.join(DST, "src", rel)) and _sha(os.path.join(DST, "src", rel)) == h
               for rel, h in manifest.items()) and set(manifest) == set(FILES)


def src_dir() -> str:
    return os.path.join(DST, "src")


if __name__ == "__main__":
    print("staged" if stage() else "reference not present; nothing staged", "| verified:", verify())
