"""In-tree build of libseqwin_b200.so: ``python -m seqwin_b200.build``."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB = HERE / "libseqwin_b200.so"


def build(verbose: bool = False, force: bool = False) -> Path:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... (see csrc/Makefile)."""
    cmd = ["make", "-C", str(HERE / "csrc"), f"-j{os.cpu_count() or 4}"]
    if force:
        subprocess.check_call(["make", "-C", str(HERE / "csrc"), "clean"], stdout=subprocess.DEVNULL)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode:
        sys.stdout.write(res.stdout)
    if res.returncode:
        raise RuntimeError("building libseqwin_b200.so failed")
    return LIB


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
