#!/usr/bin/env python3
"""The reference's OWN smoke tests against the drop-in `_core`.

    python tools/dropin_proof.py --assemble     here, where /root/reference exists: copy the reference's
                                                Python package and its tests/smoke into dropin/_pkg/
                                                (git-ignored scratch that travels to the GPU box) and drop
                                                _core.py / _lib.py / libseqwin_b200.so into seqwin/graph/
    python tools/dropin_proof.py --run [log]    on the GPU box: pytest the reference's test files with that
                                                package on PYTHONPATH; the log goes to profiles/

Nothing under dropin/_pkg is product code or committed: it is the unmodified reference package with its
native module swapped, which is exactly what a Seqwin user would run.
"""
import argparse
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "dropin" / "_pkg"
REF = Path(os.environ.get("SEQWIN_REF", "/root/reference"))


def assemble() -> None:
    if not (REF / "src" / "seqwin").is_dir():
        sys.exit(f"{REF} is not available here")
    if PKG.exists():
        shutil.rmtree(PKG)
    shutil.copytree(REF / "src" / "seqwin", PKG / "seqwin")
    shutil.copytree(REF / "tests" / "smoke", PKG / "tests_smoke")
    gdir = PKG / "seqwin" / "graph"
    for stale in gdir.glob("_core*.so"):
        stale.unlink()
    for name in ("_core.py", "_lib.py"):
        shutil.copy2(ROOT / "seqwin_b200" / name, gdir / name)
    # the library is built in-tree; link rather than copy so that the package always runs the current build
    so = gdir / "libseqwin_b200.so"
    if so.exists() or so.is_symlink():
        so.unlink()
    os.symlink(os.path.relpath(ROOT / "seqwin_b200" / "libseqwin_b200.so", gdir), so)
    print(f"assembled {PKG}")


def run(log: Path) -> int:
    if not (PKG / "seqwin").is_dir():
        sys.exit("dropin/_pkg is missing: run --assemble where the reference is available")
    env = dict(os.environ, PYTHONPATH=str(PKG))
    tests = [str(PKG / "tests_smoke" / t) for t in ("test_graph.py", "test_outputs.py")]
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--rootdir", str(PKG / "tests_smoke"),
           "-c", "/dev/null", *tests, "-rA"]
    probe = subprocess.run([sys.executable, "-c",
                            "import seqwin.graph as g, seqwin.graph._core as c; print(c.__file__); "
                            "from seqwin.graph._lib import lib; print(lib()._name)"],
                           env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=str(PKG))
    text = ("# reference tests/smoke/test_graph.py + test_outputs.py against the drop-in _core\n"
            "# native module in use:\n" + probe.stdout + "\n$ " + " ".join(cmd) + "\n" + res.stdout)
    log.parent.mkdir(parents=True, exist_ok=True)
    log.write_text(text)
    print(text[-3000:])
    return res.returncode


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--assemble", action="store_true")
    ap.add_argument("--run", nargs="?", const=str(ROOT / "gpurun_out" / "r2_dropin_reference_tests.txt"))
    a = ap.parse_args()
    if a.assemble:
        assemble()
    if a.run:
        sys.exit(run(Path(a.run)))
