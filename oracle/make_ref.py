"""ORACLE tooling (TEST INFRASTRUCTURE): make the UNMODIFIED reference travel to the GPU box.

  python -m oracle.make_ref          # build container only (needs /root/reference)

The reference is pure Python with no installer, so "building" it is a file copy: the model
wrappers of the five BASELINE configs and the two operator packages they import are copied, byte for byte,
from `/root/reference/models` into `oracle/_ref/models/` together with a manifest of their SHA-256 sums.
`oracle/_ref/` is git-ignored (reference sources never enter the history) but NOT gpurun-ignored, so the
copy ships with the snapshot like the nvcc-built library does.  Consumers: `oracle/ref_loader.py` (the
parity tests and `bench.py --impl reference` / `cpu_baseline` -- checker and baseline only, never the
product path).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/models"
DST = os.path.join(HERE, "_ref", "models")

MODEL_FILES = (
    "equiformerv2_qm9.py",
    "equiformerv2_oc20.py",
    "equiformerv2_MatPES.py",
    "equiformerv2_MatPESv2.py",
    "equiformerv2_MatPES_GATAV2.py",
    "equiformerv2_MatPES_GATAV2_phi_at_every_iteration_like_gata.py",
    "equiformerv2_MatPES_GATAV2_GLOBALALLATTENTION_HTR_phi_at_every_iteration_like_gata_with_DISTANCE.py",
)
PACKAGES = (
    "EquiformerV2Functions",
    "NewFunctions/Gotennet_morethaninspired",
    "NewFunctions/Gotennets_GATA_phi_refined_every_layer",
    "NewFunctions/GATA_and_all2all",
)


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(force=False):
    """-> path of the copy, or None when the reference tree is absent (GPU box: the shipped copy is used as is)."""
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None
    files = list(MODEL_FILES)
    for p in PACKAGES:
        for name in sorted(os.listdir(os.path.join(SRC, p))):
            if name.endswith(".py"):
                files.append(p + "/" + name)
    manifest_path = os.path.join(DST, "MANIFEST.json")
    want = {f: _sha(os.path.join(SRC, f)) for f in files}
    if not force and os.path.exists(manifest_path):
        try:
            if json.load(open(manifest_path)) == want and all(os.path.exists(os.path.join(DST, f)) for f in files):
                return DST
        except ValueError:
            pass
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for f in files:
        out = os.path.join(DST, f)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), out)
    with open(manifest_path, "w") as fh:
        json.dump(want, fh, indent=1, sort_keys=True)
    return DST


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
