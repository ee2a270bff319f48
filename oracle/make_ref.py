"""Recipe for ``oracle/_ref/``: the files of the UNMODIFIED reference that its hot path needs, copied from where they lie under
/root/reference so that the reference itself can be run (under the import stubs of ``oracle/harness.py``) where that tree
does not exist — the GPU box.  ``python -m oracle.make_ref``; ``__graft_entry__.build()`` runs it when the tree is present.

TEST / MEASUREMENT INFRASTRUCTURE.  ``oracle/_ref/`` is a build output: git-ignored (no reference source enters the history),
not gpurun-ignored (it travels with the snapshot like the built .so).  Used by ``bench.py --impl reference`` (``cpu_baseline.kind``
"reference") and by tests that compare with the reference when /root/reference is absent.
"""
from __future__ import annotations

import os
import shutil
import sys

REFERENCE = os.environ.get("STRYKE_REFERENCE_SRC", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["Stryke/__init__.py", "Stryke/stryke.py", "Stryke/barotrauma.py", "Data/fish_class.csv",
         "Data/bio_response_models_Pflugrath_2021.csv"]


def make(force=False):
    """Copy the files; returns the destination, or None when the reference tree is not there."""
    if not os.path.isfile(os.path.join(REFERENCE, "Stryke", "stryke.py")):
        return None
    for rel in FILES:
        src, dst = os.path.join(REFERENCE, rel), os.path.join(DEST, rel)
        if not force and os.path.isfile(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    return DEST


if __name__ == "__main__":
    print(make(force="--force" in sys.argv))
