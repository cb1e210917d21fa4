"""Install the UNMODIFIED reference sources of the hot path's callers under baseline/_ref/ (git-ignored).

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference is a research code drop without setup.py / pyproject.toml, so
`pip install --target baseline/_ref /root/reference` has nothing to install; this script is the equivalent: it copies,
byte for byte, the FocusRAFT model folder (`core/models/ff-raft`: FF_RAFT_Core/, losses/, common.py and the one config
every experiment shares) and the FF-PWC model files to `baseline/_ref/`, so that the GPU box -- which has no
/root/reference mount -- can run `bench.py --impl reference` and the "reference on this GPU" measurement on the
reference's own code.  `baseline/_ref/` is listed in .gitignore (nothing of the reference enters the history) but not
in .gpurunignore (it travels with the snapshot).  Run by `__graft_entry__.build()` when /root/reference exists.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("FFCORR_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")

# (source directory relative to the reference root, file patterns)
WANT = [
    ("core/models/ff-raft", ("common.py", "__init__.py")),
    ("core/models/ff-raft/FF_RAFT_Core", None),                       # every .py of the model
    ("core/models/ff-raft/FF_RAFT_Core/utils", None),
    ("core/models/ff-raft/losses", None),
    ("core/models/ff-raft/config/experiment", ("ffraft_chairs_orb.yaml", "ffraft_kitti_orb.yaml", "ffraft_sintel_orb.yaml")),
    ("core/models/ff-pwcnet", ("common.py", "__init__.py")),
    ("core/models/ff-pwcnet/PWCNet_Core", None),
    ("core/models/ff-pwcnet/losses", None),
    ("core/models/ff-pwcnet/config", ("ffpwc_chairs.yaml", "ffpwc_sintel.yaml")),
]


def install(verbose: bool = True) -> str | None:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"{SRC} is absent: keeping whatever baseline/_ref already holds")
        return DST if os.path.isdir(DST) else None
    manifest = {}
    for rel, names in WANT:
        sdir = os.path.join(SRC, rel)
        if not os.path.isdir(sdir):
            continue
        ddir = os.path.join(DST, rel)
        os.makedirs(ddir, exist_ok=True)
        for fn in sorted(os.listdir(sdir)):
            sp = os.path.join(sdir, fn)
            if not os.path.isfile(sp):
                continue
            if names is None:
                if not fn.endswith(".py"):
                    continue
            elif fn not in names:
                continue
            shutil.copyfile(sp, os.path.join(ddir, fn))
            with open(sp, "rb") as f:
                manifest[os.path.join(rel, fn)] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"installed {len(manifest)} reference files under {DST}")
    return DST


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
