"""Build ``libstryke_b200.so`` in-tree with nvcc for sm_100a:  ``python -m stryke_b200.build [--force]``."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "passage.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "passage_common.cuh"), os.path.join(HERE, "csrc", "passage_fast.cuh"),
        os.path.join(HERE, "csrc", "bootstrap.cuh"), os.path.join(HERE, "csrc", "specialise.inl"),
        os.path.join(HERE, "csrc", "passage_kernels.cuh"),
        os.path.join(os.path.dirname(HERE), "include", "stryke_b200.h")]
OUT = os.path.join(HERE, "libstryke_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-ldl"]


def needs_build():
    return not os.path.isfile(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS)


def build(force=False, verbose=False, out=None):
    global OUT
    if out:
        OUT = os.path.join(HERE, os.path.basename(out))
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = ["-DSTRYKE_BOUNDS_CHECK"] if os.environ.get("STRYKE_B200_BOUNDS_CHECK") == "1" else []   # debug build: checked table indices
    extra += os.environ.get("STRYKE_B200_NVCC_FLAGS", "").split()                                  # experiments: -DSTRYKE_COUNT_MINB=5 ...
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-o", OUT, SRC] + (["-Xptxas", "-v"] if verbose else [])
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv,
                out=sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None))
