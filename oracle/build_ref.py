"""TEST INFRASTRUCTURE ONLY.  Recipe for ``oracle/_ref``: a verbatim, git-ignored copy of the reference's Python package
(``/root/reference/src`` -- pure Python, no build step) made in the build container so that the UNMODIFIED reference
travels to the GPU box with the repo snapshot (``oracle/_ref/`` is listed in .gitignore, not in .gpurunignore; nothing
under it is ever committed).  ``oracle/shim.py`` imports the reference from ``/root/reference`` when that exists and from
``oracle/_ref`` otherwise; ``bench.py --impl reference`` and the ``cpu_baseline`` leg then time the reference's own
``MultimodalTransformer`` (``cpu_baseline.kind = "reference"``) instead of the restatement (``"port"``).

    python -m oracle.build_ref          # copies when /root/reference is present, otherwise reports what is there
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("OMR_REFERENCE_ROOT", "/root/reference")
DEST = os.path.join(HERE, "_ref")


def have_ref() -> bool:
    return os.path.isfile(os.path.join(DEST, "src", "transformer", "model.py"))


def ensure_ref(verbose: bool = False) -> bool:
    """Copy <reference>/src -> oracle/_ref/src (only *.py; refreshed when a file differs).  Returns have_ref()."""
    src = os.path.join(REF_SRC, "src")
    if os.path.isdir(src) and os.path.abspath(REF_SRC) != os.path.abspath(DEST):
        n = 0
        for root, _dirs, files in os.walk(src):
            rel = os.path.relpath(root, src)
            for f in files:
                if not f.endswith(".py"):
                    continue
                a, b = os.path.join(root, f), os.path.join(DEST, "src", rel, f)
                if not (os.path.isfile(b) and filecmp.cmp(a, b, shallow=False)):
                    os.makedirs(os.path.dirname(b), exist_ok=True)
                    shutil.copyfile(a, b)
                    n += 1
        if verbose:
            print(f"oracle/_ref: {n} file(s) refreshed from {src}")
    elif verbose:
        print(f"oracle/_ref: {REF_SRC} not present; {'using the existing copy' if have_ref() else 'no copy available'}")
    return have_ref()


if __name__ == "__main__":
    sys.exit(0 if ensure_ref(verbose=True) else 1)
