from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def _have_gpu() -> bool:
    try:
        import ctypes as C
        from seqwin_b200 import _lib
        return _lib.lib().sw_device_info(None, None, None, None) == 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def have_gpu() -> bool:
    return _have_gpu()


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build what the tests load: the CUDA library (nvcc cross-compiles on CPU), the oracle, the emulator."""
    from seqwin_b200 import build as b
    if not b.LIB.exists():
        b.build()
    from oracle import oracle as O
    O.build()
    yield


@pytest.fixture(scope="session")
def digests() -> dict:
    return json.loads((GOLDEN / "digests.json").read_text())


@pytest.fixture(scope="session")
def fixture_paths() -> list[Path]:
    f = GOLDEN / "ref_fixtures"
    return [f / "targets" / "target-1.fasta", f / "targets" / "target-2.fasta",
            f / "non-targets" / "non-target-1.fasta", f / "non-targets" / "non-target-2.fasta"]


@pytest.fixture(scope="session")
def expected_graph():
    return np.load(GOLDEN / "ref_fixtures" / "expected" / "graph.npz", allow_pickle=False)


@pytest.fixture(scope="session")
def edge_paths():
    from tests.cases import edge_case_paths
    return edge_case_paths(GOLDEN / "cases")


@pytest.fixture(scope="session")
def synth_sets(tmp_path_factory):
    """name -> (paths, is_targets) of the deterministic synthetic cases, written once per session."""
    from tests.cases import SYNTH_CASES, synth_paths
    out = {}
    for name, spec in SYNTH_CASES.items():
        out[name] = synth_paths(spec, tmp_path_factory.mktemp(name))
    return out
