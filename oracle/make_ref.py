#!/usr/bin/env python
"""Recipe for ``oracle/_ref``: the UNMODIFIED reference package, made importable next to the repo.

    python oracle/make_ref.py            # needs /root/reference (the build container)

jcoheur/pybitup is pure Python, so "building" it is copying the ``.py`` files of the ``pybitup`` package, byte for byte,
from where they lie under ``/root/reference`` into ``oracle/_ref/pybitup/``.  ``oracle/_ref/`` is git-ignored (no
reference source enters the history) but NOT gpurun-ignored, so it travels to the GPU box, where ``/root/reference``
does not exist.  A ``MANIFEST.json`` with the sha256 of every copied file is written beside it; ``verify()`` re-checks
the copy against the manifest (and against ``/root/reference`` when that is present).

Test / measurement infrastructure only: ``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` leg time
``pybitup.solve_problem.Sampling(json).sample(models)`` from this copy (through ``oracle/ref_shim.py``, which only
stubs the plotting / MPI imports the package makes at import time).  Nothing under ``pybitup_b200/`` touches it.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/pybitup"
DST = os.path.join(HERE, "_ref", "pybitup")
MANIFEST = os.path.join(HERE, "_ref", "MANIFEST.json")


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def _package_files(root):
    """The package's python modules (``pybitup/*.py``, ``pybitup/pce/*.py``); its test data and scripts are not needed."""
    out = []
    for dirpath, dirnames, filenames in os.walk(root):
        dirnames[:] = [d for d in dirnames if d not in ("test", "__pycache__")]
        for fn in sorted(filenames):
            if fn.endswith(".py"):
                out.append(os.path.relpath(os.path.join(dirpath, fn), root))
    return sorted(out)


def make(force=False):
    if not os.path.isdir(SRC):
        if os.path.exists(MANIFEST):
            return False                      # GPU box: the prebuilt copy is all there is
        raise SystemExit("oracle/make_ref.py: %s is absent and there is no prebuilt oracle/_ref" % SRC)
    files = _package_files(SRC)
    if not force and os.path.exists(MANIFEST):
        try:
            with open(MANIFEST) as f:
                man = json.load(f)
            if man.get("files") == {r: _sha(os.path.join(SRC, r)) for r in files} and all(
                    os.path.exists(os.path.join(DST, r)) and _sha(os.path.join(DST, r)) == h for r, h in man["files"].items()):
                return False                  # up to date
        except (OSError, ValueError):
            pass
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    man = {"source": SRC, "files": {}}
    for r in files:
        dst = os.path.join(DST, r)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, r), dst)
        man["files"][r] = _sha(dst)
    with open(MANIFEST, "w") as f:
        json.dump(man, f, indent=1, sort_keys=True)
    return True


def verify():
    """True when oracle/_ref holds exactly the files of the manifest (and they equal /root/reference when present)."""
    if not os.path.exists(MANIFEST):
        return False
    with open(MANIFEST) as f:
        man = json.load(f)
    for r, h in man["files"].items():
        p = os.path.join(DST, r)
        if not os.path.exists(p) or _sha(p) != h:
            return False
        if os.path.isdir(SRC) and _sha(os.path.join(SRC, r)) != h:
            return False
    return True


if __name__ == "__main__":
    changed = make(force="--force" in sys.argv)
    print("oracle/_ref %s (%d files, verified: %s)" % ("written" if changed else "up to date",
                                                        len(json.load(open(MANIFEST))["files"]), verify()))
