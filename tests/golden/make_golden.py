"""Regenerate tests/golden/ from the UNMODIFIED reference (run in the authoring container only).

    python tests/golden/make_golden.py

* copies the reference's own smoke fixtures (4 FASTA files + expected graph.npz) verbatim from
  /root/reference/tests/smoke/fixtures -- they are the golden vectors its tests pin,
* writes a set of hand-made edge-case FASTA files (tests/golden/cases/),
* runs the reference extension compiled by oracle/Makefile (oracle/_ref) on the fixtures, the edge
  cases and on deterministic synthetic sets (seqwin_b200.synth) and stores, per case and (k, w),
  the array sizes and SHA-256 digests of kmers / nodes / edges / record_offsets, before and after
  _get_penalty_native, in tests/golden/digests.json; small cases also keep the full arrays (.npz).

/root/reference does not exist on the GPU box, so tests only ever read what this script wrote.
"""
from __future__ import annotations

import gzip
import hashlib
import json
import shutil
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
HERE = Path(__file__).resolve().parent
REF_FIX = Path("/root/reference/tests/smoke/fixtures")

from oracle import oracle as O  # noqa: E402
from tests.cases import EDGE_FILES, GOLDEN_KW, SYNTH_CASES, edge_case_paths, synth_paths  # noqa: E402


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_ref(ref, paths, k, w, is_targets):
    kmers, nodes, edges, offsets, ids = ref._build_native([str(p) for p in paths], k, w, 2, False)
    out = {"n_kmers": len(kmers), "n_nodes": len(nodes), "n_edges": len(edges),
           "kmers": digest(kmers), "nodes": digest(nodes), "edges": digest(edges),
           "record_offsets": digest(offsets), "offsets": offsets.tolist(),
           "ids": hashlib.sha256(repr(ids).encode()).hexdigest()}
    arrays = {"kmers": kmers, "nodes_build": nodes.copy(), "edges": edges, "record_offsets": offsets}
    if is_targets is not None and any(is_targets) and not all(is_targets):
        ref._get_penalty_native(kmers, nodes, offsets, np.asarray(is_targets, dtype=np.bool_), 2)
        out["nodes_penalty"] = digest(nodes)
        arrays["nodes_penalty"] = nodes
    return out, arrays


def main() -> None:
    ref = O.load_reference()
    assert ref is not None, "reference not built (make -C oracle ref)"
    # 1. the reference's own fixtures, verbatim
    dst = HERE / "ref_fixtures"
    if dst.exists():
        shutil.rmtree(dst)
    (dst / "targets").mkdir(parents=True)
    (dst / "non-targets").mkdir()
    (dst / "expected").mkdir()
    for sub in ("targets", "non-targets"):
        for f in sorted((REF_FIX / sub).glob("*.fasta")):
            shutil.copy(f, dst / sub / f.name)
    shutil.copy(REF_FIX / "expected" / "graph.npz", dst / "expected" / "graph.npz")

    # 2. edge-case FASTA files
    cases = HERE / "cases"
    if cases.exists():
        shutil.rmtree(cases)
    cases.mkdir()
    for name, content in EDGE_FILES.items():
        if name.endswith(".gz"):
            with gzip.GzipFile(cases / name, "wb", mtime=0) as fh:
                fh.write(content)
        else:
            (cases / name).write_bytes(content)

    digests: dict = {}
    fix_paths = [dst / "targets" / "target-1.fasta", dst / "targets" / "target-2.fasta",
                 dst / "non-targets" / "non-target-1.fasta", dst / "non-targets" / "non-target-2.fasta"]
    arrays_dir = HERE / "arrays"
    if arrays_dir.exists():
        shutil.rmtree(arrays_dir)
    arrays_dir.mkdir()

    def record(case, paths, is_targets, keep_arrays):
        digests[case] = {}
        for (k, w) in GOLDEN_KW:
            d, arrs = run_ref(ref, paths, k, w, is_targets)
            digests[case][f"{k},{w}"] = d
            if keep_arrays:
                np.savez_compressed(arrays_dir / f"{case}_{k}_{w}.npz", **arrs)

    record("fixtures", fix_paths, [True, True, False, False], True)
    ep, et = edge_case_paths(cases)
    record("edge", ep, et, True)
    import tempfile
    for name, spec in SYNTH_CASES.items():
        with tempfile.TemporaryDirectory() as td:
            paths, is_t = synth_paths(spec, Path(td))
            record(name, paths, list(is_t), False)
    (HERE / "digests.json").write_text(json.dumps(digests, indent=1, sort_keys=True))
    print("wrote", HERE / "digests.json", {c: len(v) for c, v in digests.items()})


if __name__ == "__main__":
    main()
