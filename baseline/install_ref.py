#!/usr/bin/env python
"""Installs the UNMODIFIED reference (BerkIGuler/AdaFortiTran) under ``baseline/_ref/`` for ``bench.py --impl reference``.

The reference is a plain script tree (no ``setup.py`` / ``pyproject.toml``), so
``pip install --no-index --target baseline/_ref /root/reference`` fails ("neither 'setup.py' nor 'pyproject.toml' found");
the outcome is recorded in ``baseline/_ref/INSTALL.json``.  What the reference's own README prescribes instead is running
from the checkout with the repository root on ``PYTHONPATH``; this script reproduces exactly that: it copies the
``src/`` package and the ``config/`` YAMLs byte for byte (no file is edited; sha256 of every file is written to
``baseline/_ref/MANIFEST.json``).  ``baseline/_ref/`` is git-ignored (the reference's sources never enter this repository's
history) but it is not gpurun-ignored, so it travels to the GPU box like the built ``.so`` files do.

usage: python baseline/install_ref.py [--reference /root/reference] [--try-pip]
"""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 16), b""):
            h.update(blk)
    return h.hexdigest()


def install(reference="/root/reference", try_pip=False):
    if not os.path.isdir(os.path.join(reference, "src", "models")):
        raise SystemExit(f"{reference}: not a checkout of the reference (src/models missing)")
    record = {"reference": reference, "pip": "not attempted"}
    if try_pip:
        tmp = os.path.join(HERE, "_pip_probe")
        r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                            "--find-links", "/opt/wheelhouse", "--target", tmp, reference], capture_output=True, text=True)
        record["pip"] = {"returncode": r.returncode, "tail": (r.stderr or r.stdout).strip().splitlines()[-3:]}
        shutil.rmtree(tmp, ignore_errors=True)
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    manifest = {}
    for sub in ("src", "config"):
        for root, dirs, files in os.walk(os.path.join(reference, sub)):
            dirs[:] = [d for d in dirs if d != "__pycache__"]
            for fn in files:
                if fn.endswith(".pyc"):
                    continue
                s = os.path.join(root, fn)
                rel = os.path.relpath(s, reference)
                d = os.path.join(DEST, rel)
                os.makedirs(os.path.dirname(d), exist_ok=True)
                shutil.copyfile(s, d)
                manifest[rel] = sha256(d)
    for name in ("LICENSE", "requirements.txt"):
        s = os.path.join(reference, name)
        if os.path.exists(s):
            shutil.copyfile(s, os.path.join(DEST, name))
            manifest[name] = sha256(os.path.join(DEST, name))
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    with open(os.path.join(DEST, "INSTALL.json"), "w") as f:
        json.dump(record, f, indent=1)
    return DEST, len(manifest)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--try-pip", action="store_true")
    a = ap.parse_args()
    dest, n = install(a.reference, a.try_pip)
    print(f"installed {n} reference files under {dest}")
