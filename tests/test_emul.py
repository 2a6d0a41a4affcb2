"""The sketch kernel's per-thread phase functions, executed serially on the host (tests/emul) and
compared with the oracle: tiling, halos, N gaps, ties, tiny tiles.  CPU only."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O

EMUL_DIR = Path(__file__).resolve().parent / "emul"


@pytest.fixture(scope="module")
def emul():
    so = EMUL_DIR / "libemul.so"
    srcs = [EMUL_DIR / "emul.cpp"] + list((EMUL_DIR.parents[1] / "seqwin_b200" / "csrc").glob("*.h")) + \
        [EMUL_DIR.parents[1] / "seqwin_b200" / "csrc" / "ingest.cpp"]
    if not so.exists() or so.stat().st_mtime < max(s.stat().st_mtime for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-fPIC", "-shared", "-o", str(so),
                               str(EMUL_DIR / "emul.cpp"), "-lz"])
    L = C.CDLL(str(so))
    L.emul_sketch.restype = C.c_long
    L.emul_sketch_sparse.restype = C.c_long
    L.emul_sketch_files.restype = C.c_long

    def run(seqs, k, w, nt, c1, force_generic=0):
        n = len(seqs)
        arrs = [np.frombuffer(s, dtype=np.uint8) for s in seqs]
        ptrs = (C.c_void_p * max(1, n))(*[a.ctypes.data for a in arrs])
        lens = np.array([len(s) for s in seqs], dtype=np.uint32)
        cap = sum(len(s) for s in seqs) + 16
        h1, pos, rec = np.empty(cap, np.uint64), np.empty(cap, np.uint32), np.empty(cap, np.uint32)
        nt_out = C.c_uint32()
        m = L.emul_sketch(ptrs, C.c_void_p(lens.ctypes.data), C.c_size_t(n), C.c_uint32(k), C.c_uint32(w), nt, c1, force_generic,
                          C.c_void_p(h1.ctypes.data), C.c_void_p(pos.ctypes.data), C.c_void_p(rec.ctypes.data),
                          C.c_size_t(cap), C.byref(nt_out))
        assert m >= 0, f"emulator failed ({m})"
        return h1[:m], pos[:m], rec[:m], nt_out.value

    def run_sparse(seqs, k, w, nt, c1, cand_per_window=0.0):
        n = len(seqs)
        arrs = [np.frombuffer(s, dtype=np.uint8) for s in seqs]
        ptrs = (C.c_void_p * max(1, n))(*[a.ctypes.data for a in arrs])
        lens = np.array([len(s) for s in seqs], dtype=np.uint32)
        cap = sum(len(s) for s in seqs) + 16
        h1, pos, rec = np.empty(cap, np.uint64), np.empty(cap, np.uint32), np.empty(cap, np.uint32)
        nt_out, nf_out = C.c_uint32(), C.c_uint32()
        m = L.emul_sketch_sparse(ptrs, C.c_void_p(lens.ctypes.data), C.c_size_t(n), C.c_uint32(k), C.c_uint32(w), nt, c1,
                                 C.c_double(cand_per_window), C.c_void_p(h1.ctypes.data), C.c_void_p(pos.ctypes.data),
                                 C.c_void_p(rec.ctypes.data), C.c_size_t(cap), C.byref(nt_out), C.byref(nf_out))
        assert m >= 0, f"emulator failed ({m})"
        return h1[:m], pos[:m], rec[:m], nt_out.value, nf_out.value
    def run_files(paths, k, w, sparse):
        arr = (C.c_char_p * max(1, len(paths)))(*[str(p).encode() for p in paths])
        cap = 4_000_000
        h1, pos, rec = np.empty(cap, np.uint64), np.empty(cap, np.uint32), np.empty(cap, np.uint32)
        m = L.emul_sketch_files(arr, C.c_size_t(len(paths)), C.c_uint32(k), C.c_uint32(w), int(sparse),
                                C.c_void_p(h1.ctypes.data), C.c_void_p(pos.ctypes.data), C.c_void_p(rec.ctypes.data),
                                C.c_size_t(cap))
        assert 0 <= m <= cap, f"emulator failed ({m})"
        return h1[:m], pos[:m], rec[:m]
    run.sparse = run_sparse
    run.files = run_files
    return run


def _oracle_stream(seqs, k, w):
    H, P, R = [], [], []
    for r, s in enumerate(seqs):
        h, p = O.minimize(s, k, w)
        H.append(h), P.append(p), R.append(np.full(len(h), r, np.uint32))
    return np.concatenate(H), np.concatenate(P), np.concatenate(R)


def _seqs(rng):
    def rand(n, p_n=0.0, low=False, alphabet=b"ACGT"):
        a = np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), n)].copy()
        if p_n:
            for s in rng.integers(0, max(1, n), max(1, int(n * p_n / 5))):
                a[s:s + rng.integers(1, 12)] = ord("N")
        if low:
            a[rng.random(n) < 0.2] |= 0x20
        return a.tobytes()
    return [rand(int(rng.integers(1, 4000))) for _ in range(3)] + [
        rand(3000, 0.01, True), rand(2500, alphabet=b"AC"), rand(1500, alphabet=b"A"), rand(2000, 0.05), b"", b"ACGT",
        rand(12000), rand(40000, 0.002), b"ACGTTGCA" * 400]


CONFIGS = [(8, 11), (32, 9), (4, 45), (128, 45), (128, 27), (128, 33), (256, 17), (256, 21), (256, 61)]
KW = [(21, 200), (21, 10), (17, 10), (7, 10), (5, 3), (4, 1), (6, 7), (3, 2), (31, 50), (21, 46), (21, 45), (21, 47),
      (9, 9), (11, 16), (15, 64), (12, 11), (33, 100), (127, 10)]


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: f"nt{c[0]}c{c[1]}")
def test_emulated_kernel_matches_oracle(emul, cfg):
    nt, c1 = cfg
    rng = np.random.default_rng(nt * 1000 + c1)
    n_multi_tile = 0
    for k, w in KW:
        if nt * c1 <= w + 1:
            continue
        seqs = _seqs(rng)
        oh, op, orr = _oracle_stream(seqs, k, w)
        # the specialised path (taken when w - 1 >= C1) and the generic path must both match
        for force_generic in ((0, 1) if w - 1 >= c1 else (0,)):
            eh, ep, er, n_tiles = emul(seqs, k, w, nt, c1, force_generic)
            assert len(eh) == len(oh), (cfg, k, w, force_generic, len(eh), len(oh))
            assert np.array_equal(eh, oh) and np.array_equal(ep, op) and np.array_equal(er, orr), (cfg, k, w, force_generic)
        n_multi_tile += n_tiles > len(seqs)
    assert n_multi_tile > 0


SPARSE_CONFIGS = [(128, 64), (8, 64), (16, 32), (4, 48)]
SPARSE_KW = [(21, 200), (21, 96), (31, 150), (15, 128), (17, 333), (7, 100), (21, 120)]


@pytest.mark.parametrize("cfg", SPARSE_CONFIGS, ids=lambda c: f"nt{c[0]}c{c[1]}")
def test_emulated_sparse_kernel_matches_oracle(emul, cfg):
    """sketch_sparse_kernel's phases (candidate lists, neighbour scans) plus the dense hand-over.
    Thresholds from 'a few candidates per window' (most tiles handed over for lack of coverage) to
    'every k-mer is a candidate' (private lists overflow) must all give the oracle's stream."""
    nt, c1 = cfg
    rng = np.random.default_rng(nt * 77 + c1)
    n_sparse_tiles = n_fallback_tiles = 0
    for k, w in SPARSE_KW:
        if nt * c1 < 4 * w // 3:
            continue
        seqs = _seqs(rng) + [b"ACGTTGCA" * 40 + bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 6000)) + b"AC" * 900]
        oh, op, orr = _oracle_stream(seqs, k, w)
        for cpw in (0.0, 3.0, 8.0, 40.0, 1e9):
            eh, ep, er, n_tiles, n_fb = emul.sparse(seqs, k, w, nt, c1, cpw)
            assert len(eh) == len(oh), (cfg, k, w, cpw, len(eh), len(oh))
            assert np.array_equal(eh, oh) and np.array_equal(ep, op) and np.array_equal(er, orr), (cfg, k, w, cpw)
            n_sparse_tiles += n_tiles - n_fb
            n_fallback_tiles += n_fb
    assert n_sparse_tiles > 0 and n_fallback_tiles > 0


def test_sparse_tie_rule_on_repeats(emul):
    """Equal hashes inside one window (tandem repeats longer than w): the rightmost minimum wins and
    the left / right scans must treat ties asymmetrically."""
    rng = np.random.default_rng(5)
    unit = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 37))
    seqs = [unit * 300, (unit + b"A") * 250, bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 500)) + unit * 200]
    for k, w in ((11, 100), (21, 200), (15, 120)):
        oh, op, orr = _oracle_stream(seqs, k, w)
        for cpw in (0.0, 1e9):
            for nt, c1 in ((8, 64), (128, 64)):
                eh, ep, er, _, _ = emul.sparse(seqs, k, w, nt, c1, cpw)
                assert np.array_equal(eh, oh) and np.array_equal(ep, op) and np.array_equal(er, orr), (k, w, cpw, nt)


def test_sparse_kernel_property_random_parameters(emul):
    """Random (k, w, candidate density, sequence composition) draws: the sparse phases + hand-over must
    reproduce the oracle for every draw (seeded, so failures reproduce)."""
    rng = np.random.default_rng(20261017)
    alphabets = [b"ACGT", b"ACGT", b"ACGTN", b"AC", b"ACGTacgt", b"AAAAAAAC"]
    for trial in range(40):
        k = int(rng.integers(3, 40))
        w = int(rng.integers(64, 420))
        nt, c1 = [(8, 64), (16, 32), (4, 48)][trial % 3]
        if nt * c1 < 4 * w // 3:
            nt, c1 = 8, 64
        cpw = float(rng.choice([0.0, 2.0, 6.0, 11.0, 25.0, 1e9]))
        seqs = []
        for _ in range(int(rng.integers(1, 5))):
            alpha = np.frombuffer(alphabets[int(rng.integers(0, len(alphabets)))], dtype=np.uint8)
            n = int(rng.integers(1, 6000))
            s = alpha[rng.integers(0, len(alpha), n)].copy()
            if rng.random() < 0.4 and n > 300:      # a tandem repeat longer than the window
                unit = s[:int(rng.integers(1, 60))]
                rep = np.tile(unit, 1 + (w + 80) // len(unit))
                at = int(rng.integers(0, n - 1))
                s = np.concatenate([s[:at], rep, s[at:]])
            seqs.append(s.tobytes())
        oh, op, orr = _oracle_stream(seqs, k, w)
        eh, ep, er, _, _ = emul.sparse(seqs, k, w, nt, c1, cpw)
        assert np.array_equal(eh, oh) and np.array_equal(ep, op) and np.array_equal(er, orr), (trial, k, w, nt, c1, cpw)


@pytest.mark.parametrize("kw", [(17, 10), (21, 200), (5, 3)], ids=lambda kw: f"k{kw[0]}w{kw[1]}")
def test_fasta_ingest_then_emulated_sketch_matches_oracle(emul, kw, fixture_paths, edge_paths):
    """FASTA / .gz files -> host parser + 2-bit packer (csrc/ingest.cpp) -> emulated sketch kernels, against
    the oracle's graph built from the same files: the parser's line rules (blank lines, CR LF, inner
    whitespace, IUPAC codes, lower case, empty records, gzip) checked without a GPU."""
    k, w = kw
    for paths in (fixture_paths, edge_paths[0]):
        paths = [str(p) for p in paths]
        kmers, nodes, _, _, _ = O._build_native(paths, k, w)
        hashes = np.repeat(nodes["hash"], (nodes["stop"] - nodes["start"]).astype(np.int64))
        order = np.lexsort((kmers["pos"], kmers["record_idx"]))
        want = (hashes[order], kmers["pos"][order], kmers["record_idx"][order])
        for sparse in ((0, 1) if w >= 96 else (0,)):
            h1, pos, rec = emul.files(paths, k, w, sparse)
            assert len(h1) == len(want[0]), (kw, sparse, len(h1), len(want[0]))
            assert np.array_equal(h1, want[0]) and np.array_equal(pos, want[1]) and np.array_equal(rec, want[2]), (kw, sparse)


def test_fasta_parser_line_width_fast_path(emul, tmp_path):
    """The parser predicts where a sequence line ends from the width of the previous one.  Files built
    to defeat the prediction -- CR LF, widths that change, short last lines, blank lines, an N inside a
    full-width line, leading blanks, a header right after a short line -- must still give the oracle's
    minimizers."""
    rng = np.random.default_rng(3)

    def seq(n, p_n=0.0):
        a = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].copy()
        if p_n:
            a[rng.random(n) < p_n] = ord("N")
        return a.tobytes().decode()

    files = []

    def write(name, recs, width, eol="\n", lower=False, blank=False):
        p = tmp_path / name
        with open(p, "w", newline="") as f:
            for rid, s in recs:
                f.write(f">{rid} desc{eol}")
                s = s.lower() if lower else s
                for i in range(0, len(s), width):
                    f.write(s[i:i + width] + eol)
                    if blank and i % (7 * width) == 0:
                        f.write(eol)
        files.append(str(p))

    write("a.fa", [("r1", seq(5000)), ("r2", seq(3333, 0.01))], 80)
    write("b.fa", [("r1", seq(4000)), ("r2", seq(1000))], 60, eol="\r\n")
    write("c.fa", [("r1", seq(2500, 0.002))], 70, lower=True, blank=True)
    write("d.fa", [("r1", seq(100)), ("r2", seq(81)), ("r3", seq(80)), ("r4", seq(79)), ("r5", "")], 80)
    write("e.fa", [("r1", seq(6000))], 33)
    s1 = seq(80)
    with open(tmp_path / "f.fa", "w") as f:
        f.write(">x\n" + s1 + "\n" + s1[:40] + "N" + s1[41:] + "\n" + seq(30) + "\n>y\n" + seq(80) + "\n" + seq(80) + "\n"
                + "   " + seq(77) + "\n" + seq(80))   # no newline at the end of the file
    files.append(str(tmp_path / "f.fa"))
    for k, w in ((5, 3), (17, 10), (21, 50)):
        kmers, nodes, _, _, _ = O._build_native(files, k, w)
        hashes = np.repeat(nodes["hash"], (nodes["stop"] - nodes["start"]).astype(np.int64))
        order = np.lexsort((kmers["pos"], kmers["record_idx"]))
        h1, pos, rec = emul.files(files, k, w, 0)
        assert len(h1) == len(order), (k, w, len(h1), len(order))
        assert np.array_equal(h1, hashes[order]) and np.array_equal(pos, kmers["pos"][order]), (k, w)
        assert np.array_equal(rec, kmers["record_idx"][order]), (k, w)
