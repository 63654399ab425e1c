"""Install the UNMODIFIED reference into oracle/_ref/ (test / baseline infrastructure, never shipped).

    python oracle/build_ref.py            (build container only: needs /root/reference)

`pip install --no-index --no-deps --target oracle/_ref <copy of /root/reference>`: the reference is pure Python, so
this is a plain package install (its own setup.py, no build system of ours).  oracle/_ref/ is git-ignored (no
reference source enters the history) but travels to the GPU box with gpurun, where bench.py's CPU arm
(`--impl reference`, `cpu_baseline.kind = "reference"`) imports it; nothing under tests/ or the product does.
SOURCE_SHA256 records the digest of the installed package sources.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = "/root/reference"


def tree_sha256(root: str) -> str:
    h = hashlib.sha256()
    for d, _, files in sorted(os.walk(root)):
        for f in sorted(files):
            if f.endswith(".py"):
                p = os.path.join(d, f)
                h.update(os.path.relpath(p, root).encode())
                with open(p, "rb") as fh:
                    h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False) -> bool:
    """True if oracle/_ref holds the reference afterwards."""
    pkg = os.path.join(DEST, "continuum_robot")
    if not os.path.isdir(os.path.join(SRC, "src", "continuum_robot")):
        return os.path.isdir(pkg)  # GPU box: use what travelled
    want = tree_sha256(os.path.join(SRC, "src", "continuum_robot"))
    stamp = os.path.join(DEST, "SOURCE_SHA256")
    if not force and os.path.isdir(pkg) and os.path.exists(stamp) and open(stamp).read().split()[0] == want:
        return True
    shutil.rmtree(DEST, ignore_errors=True)
    with tempfile.TemporaryDirectory() as tmp:  # /root/reference is read-only and setuptools writes build files
        work = os.path.join(tmp, "reference")
        shutil.copytree(SRC, work, ignore=shutil.ignore_patterns(".git", "__pycache__"))
        subprocess.check_call([sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation",
                               "--no-deps", "--find-links", "/opt/wheelhouse", "--target", DEST, work])
    got = tree_sha256(pkg)
    if got != want:
        raise RuntimeError("installed reference differs from /root/reference/src/continuum_robot")
    with open(stamp, "w") as f:
        f.write(f"{want}  continuum_robot (*.py under /root/reference/src/continuum_robot)\n")
    return True


if __name__ == "__main__":
    print("oracle/_ref ready:", build(force="--force" in sys.argv))
