"""TEST / BENCH INFRASTRUCTURE ONLY -- stage the UNMODIFIED reference env for the CPU arm of bench.py.

    python oracle/make_ref.py            (also run by __graft_entry__.build() when /root/reference is present)

The reference (LahiruCooray/rl-aerial-manipulator) is a plain Python tree without packaging, so there is nothing to
`pip install`; and /root/reference does not exist on the GPU box.  This recipe copies the files of the hot path
(the env modules and the simul_files / utils2 packages they import) byte for byte into the
git-ignored `oracle/_ref/` (it travels to the GPU box with the snapshot like the built .so files, and never enters the
history), together with MANIFEST.json (sha256 of every file, so "unmodified" can be checked on the box).

bench.py's `--impl reference` arm and `cpu_baseline` then time the reference's own `WaypointQuadEnv.step`
(initial-implementation-v2/rl_env_scaledObs.py:9,123) under the gymnasium stub of oracle/ref_harness.py -- `kind:
"reference"`.  Without `oracle/_ref/` they fall back to the NumPy port (`kind: "port"`).  Nothing in the product package
imports any of this.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC_DEFAULT = "/root/reference"

# the hot path's files, relative to the reference root
FILES = [
    "initial-implementation-v2/rl_env_scaledObs.py",
    "initial-implementation-v2/simul_files/model/params.py",
    "initial-implementation-v2/simul_files/model/quadcopter.py",
    "initial-implementation-v2/simul_files/utils/quaternion.py",
    "initial-implementation-v2/simul_files/utils/utils.py",
    "initial-implementation-v2/utils2/utils.py",
    "initial-implementation-v1/rl_env_scaledObs.py",
    "initial-implementation-v1/rl_env.py",
    "initial-implementation-v1/simul_files/model/params.py",
    "initial-implementation-v1/simul_files/model/quadcopter.py",
    "initial-implementation-v1/simul_files/utils/quaternion.py",
    "initial-implementation-v1/simul_files/utils/utils.py",
]


def sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def staged() -> bool:
    return os.path.exists(os.path.join(DEST, "MANIFEST.json"))


def verify() -> bool:
    """True when every staged file still has the sha256 recorded at staging time."""
    if not staged():
        return False
    man = json.load(open(os.path.join(DEST, "MANIFEST.json")))
    return all(os.path.exists(os.path.join(DEST, f)) and sha256(os.path.join(DEST, f)) == d for f, d in man["files"].items())


def make(src: str = SRC_DEFAULT, verbose: bool = True) -> bool:
    if not os.path.isdir(os.path.join(src, "initial-implementation-v2")):
        if verbose:
            print(f"[make_ref] no reference tree at {src}; keeping whatever oracle/_ref holds")
        return staged()
    man = {"source": src, "files": {}}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        man["files"][rel] = sha256(d)
        assert man["files"][rel] == sha256(s)
    json.dump(man, open(os.path.join(DEST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"[make_ref] staged {len(FILES)} unmodified reference files under {DEST}")
    return True


if __name__ == "__main__":
    ok = make(sys.argv[1] if len(sys.argv) > 1 else SRC_DEFAULT)
    sys.exit(0 if ok else 1)
