"""Golden vectors for the edge / node filter, from the UNMODIFIED reference.

    python tests/golden/make_filter_golden.py          (authoring container only)

Imports the reference's own `_filter_edges_and_nodes` (src/seqwin/kmers.py:132-173) -- the package
is assembled in a scratch directory from /root/reference/src/seqwin plus the extension that
oracle/Makefile compiled (oracle/_ref) -- and runs it on the scored golden graphs that
tests/golden/arrays/ already holds.  Stores, per case and threshold, the filtered nodes / edges
(small cases) in tests/golden/arrays/filter_*.npz.  /root/reference does not exist on the GPU box;
tests only read what this script wrote.
"""
from __future__ import annotations

import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]


def load_reference_kmers_module():
    scratch = Path(tempfile.mkdtemp(prefix="seqwin_ref_pkg_"))
    shutil.copytree("/root/reference/src/seqwin", scratch / "seqwin")
    so = next((ROOT / "oracle" / "_ref" / "seqwin_ref").glob("_core*.so"))
    shutil.copy(so, scratch / "seqwin" / "graph" / so.name)
    sys.path.insert(0, str(scratch))
    from seqwin import kmers  # noqa: E402  (the reference's module, unmodified)
    return kmers


def main() -> None:
    K = load_reference_kmers_module()
    out_dir = HERE / "arrays"
    for case, kw in (("fixtures", (17, 10)), ("edge", (17, 10)), ("edge", (21, 200)), ("fixtures", (31, 50))):
        a = np.load(out_dir / f"{case}_{kw[0]}_{kw[1]}.npz", allow_pickle=False)
        nodes, edges = a["nodes_penalty"], a["edges"]
        res = {}
        for th in (0.0, 0.9, 1.0, 2.0, 3.5, 1e9):
            n2, e2, _ = K._filter_edges_and_nodes(nodes, edges, th)
            tag = str(th).replace(".", "p").replace("+", "")
            res[f"nodes_{tag}"] = n2
            res[f"edges_{tag}"] = e2
        np.savez_compressed(out_dir / f"filter_{case}_{kw[0]}_{kw[1]}.npz", **res)
        print(case, kw, {k: len(v) for k, v in res.items()})


if __name__ == "__main__":
    main()
