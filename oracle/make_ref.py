"""Recipe that populates ``oracle/_ref/`` with the UNMODIFIED reference modules of the hot path.

TEST / BENCH INFRASTRUCTURE.  The reference is pure Python (no native code to compile), so "building" it means copying
the handful of files the path consists of, byte for byte, from the read-only checkout at ``/root/reference`` into
``oracle/_ref/`` (git-ignored: reference sources never enter this repository's history; NOT gpurun-ignored: the copy
travels to the GPU box, where ``/root/reference`` does not exist).  ``bench.py --impl reference`` and the ``cpu_baseline``
/ ``eager_b200`` legs import them from there through ``oracle/ref_harness.py`` and run the reference's own
``DiscreteDiffusion.p_sample`` (RQC/diffusion.py:53) / ``linear_inversion`` (RQC/reconstruct.py:56) as they are.

    python oracle/make_ref.py            # copy + write MANIFEST.json (sha256 per file)
    python oracle/make_ref.py --check    # verify an existing copy against its manifest

``__graft_entry__.build()`` runs this whenever ``/root/reference`` is present.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

REF_ROOT = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = [
    "versions/RQC_dataset_building_phase/model.py",
    "versions/RQC_dataset_building_phase/diffusion.py",
    "versions/RQC_dataset_building_phase/reconstruct.py",
    "versions/RQC_dataset_building_phase/config.py",
    "versions/multi_qubit_special_states/model.py",
    "versions/multi_qubit_special_states/diffusion.py",
    "versions/multi_qubit_special_states/reconstruct.py",
    "versions/multi_qubit_special_states/config.py",
]


def _sha(path: str) -> str:
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def make() -> dict:
    if not os.path.isdir(os.path.join(REF_ROOT, "versions")):
        raise FileNotFoundError(f"{REF_ROOT} is not mounted: oracle/_ref can only be populated in the build container")
    manifest = {}
    for rel in FILES:
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF_ROOT, rel), dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_ROOT, "files": manifest}, f, indent=1)
    return manifest


def check() -> bool:
    try:
        manifest = json.load(open(os.path.join(DEST, "MANIFEST.json")))["files"]
    except (OSError, KeyError, ValueError):
        return False
    return all(os.path.exists(os.path.join(DEST, rel)) and _sha(os.path.join(DEST, rel)) == h for rel, h in manifest.items()) \
        and set(manifest) == set(FILES)


if __name__ == "__main__":
    if "--check" in sys.argv:
        ok = check()
        print("oracle/_ref", "ok" if ok else "missing or modified")
        sys.exit(0 if ok else 1)
    for rel, h in make().items():
        print(h[:16], rel)
