"""Recipe for oracle/_ref: the reference's own decoder, unmodified, where bench.py's CPU arm can import it on the GPU box.

    python oracle/make_ref.py            (run by __graft_entry__.build() in the build container)

/root/reference does not exist on the GPU box and is pure Python (nothing to compile), so "building" the reference
checker means placing the three torch-only modules its decoder needs -- src/models/components/{diinn,rdn,common}.py --
under oracle/_ref/ (git-ignored: reference sources never enter this repository's history; not gpurun-ignored, so the
directory travels with the snapshot like the built .so). Files are copied byte for byte; a manifest records their sha256.
bench.py --impl reference and bench.py's cpu_baseline import ImplicitDecoder from there (kind "reference") and fall back to
the oracle's port of the same algorithm (kind "port") only when the directory is absent.

Test infrastructure: nothing in the product package imports oracle/.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
DST = os.path.join(HERE, "_ref")
FILES = ["src/models/components/diinn.py", "src/models/components/rdn.py", "src/models/components/common.py"]
PACKAGES = ["src", "src/models", "src/models/components"]


def make_ref(verbose: bool = False) -> bool:
    """-> True if oracle/_ref holds the reference decoder afterwards."""
    if not os.path.isdir(REF_ROOT):
        return os.path.exists(os.path.join(DST, FILES[0]))
    manifest = {}
    for pkg in PACKAGES:
        os.makedirs(os.path.join(DST, pkg), exist_ok=True)
        init = os.path.join(DST, pkg, "__init__.py")
        if not os.path.exists(init):
            open(init, "w").close()
    for rel in FILES:
        src, dst = os.path.join(REF_ROOT, rel), os.path.join(DST, rel)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_ROOT, "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"oracle/_ref: {len(FILES)} files from {REF_ROOT}")
    return True


def import_reference_decoder():
    """-> the reference's ImplicitDecoder class from oracle/_ref, or None if the directory is absent."""
    if not os.path.exists(os.path.join(DST, FILES[0])):
        return None
    if DST not in sys.path:
        sys.path.insert(0, DST)
    import importlib
    mod = importlib.import_module("src.models.components.diinn")
    if not os.path.abspath(mod.__file__).startswith(DST):
        # another `src` package was imported first (in the build container: /root/reference itself, by the live
        # oracle-vs-reference tests). Accept it only if it is byte-identical to the copy the manifest describes.
        with open(os.path.join(DST, "MANIFEST.json")) as f:
            want = json.load(f)["sha256"][FILES[0]]
        with open(mod.__file__, "rb") as f:
            if hashlib.sha256(f.read()).hexdigest() != want:
                return None
    return mod.ImplicitDecoder


if __name__ == "__main__":
    ok = make_ref(verbose=True)
    print("reference decoder available:", ok)
