"""Shared definitions of the parity cases (used by tests/ and tests/golden/make_golden.py)."""
from __future__ import annotations

from pathlib import Path

import numpy as np

from seqwin_b200.synth import SynthSpec, write_set

# (k, w) pairs every golden case is pinned on: the reference test settings (17,10), (7,10), the
# defaults (21,200), the survey's stress pairs, w=1, a window as long as a tile chunk, large k.
GOLDEN_KW = [(17, 10), (7, 10), (21, 200), (5, 3), (4, 1), (6, 7), (31, 50), (21, 46), (63, 20), (127, 10), (3, 2)]

_rng = np.random.default_rng(7)


def _rand(n, alphabet=b"ACGT"):
    return np.frombuffer(alphabet, dtype=np.uint8)[_rng.integers(0, len(alphabet), n)].tobytes()


def _wrap(seq: bytes, width=60, eol=b"\n") -> bytes:
    return eol.join(seq[i:i + width] for i in range(0, len(seq), width)) + eol


_s1 = _rand(3000)
_s2 = bytearray(_rand(2500))
_s2[100:130] = b"N" * 30
_s2[1000:1001] = b"R"
_s2[1500:1700] = bytes(_s2[1500:1700]).lower()
_s2[2000:2010] = b"-" * 10
_s3 = _rand(1200).replace(b"T", b"U")
_s4 = b"ACGT" * 300                       # period-4 repeats: long runs of tied hashes
_s5 = b"A" * 700 + _rand(300) + b"C" * 500  # homopolymers
_s6 = _rand(40)                           # shorter than most k+w-1
_s7 = bytearray(_rand(900))
for _i in range(0, 900, 37):
    _s7[_i] = ord("N")                    # N every 37 bases: many short runs

EDGE_FILES: dict[str, bytes] = {
    # plain, two records, header with description
    "a_plain.fasta": b">rec1 some description\n" + _wrap(_s1) + b">rec2\tother\n" + _wrap(_s4),
    # CRLF line ends, blank lines, inner spaces, lower case, IUPAC, gaps
    "b_crlf.fasta": b">crlf_1 x\r\n" + _wrap(bytes(_s2), 70, b"\r\n") + b"\r\n   \r\n>crlf_2\r\n"
                    + b"ACGT ACGT\tACGT\r\n" + _wrap(_s3, 50, b"\r\n"),
    # empty-id header, record with no sequence, tiny record, homopolymers, no trailing newline
    "c_odd.fasta": b">\n" + _wrap(_s5) + b">empty_record\n>tiny\n" + _s6 + b"\n>dense_n\n" + _wrap(bytes(_s7))[:-1],
    # gzip input (detected by suffix only)
    "d_gz.fasta.gz": b">gz1\n" + _wrap(_rand(2000)) + b">gz2\n" + _wrap(_s1[500:2500]),
    # empty file: zero records, still an assembly slot
    "e_empty.fasta": b"",
    # shares sequence with a_plain so that nodes / edges span assemblies
    "f_shared.fasta": b">sh1\n" + _wrap(_s1[:1500] + _rand(500) + _s1[1500:]) + b">sh2\n" + _wrap(_s4[:600]),
}


def edge_case_paths(cases_dir: Path) -> tuple[list[Path], list[bool]]:
    names = sorted(EDGE_FILES)
    return [Path(cases_dir) / n for n in names], [True, True, False, False, True, False]


SYNTH_CASES = {
    "synth_small": SynthSpec(n_genomes=6, n_targets=2, genome_len=20_000, n_contigs=3, seed=11, n_runs_every=2),
    "synth_medium": SynthSpec(n_genomes=12, n_targets=4, genome_len=120_000, n_contigs=7, seed=5, n_runs_every=3),
    "synth_skew": SynthSpec(n_genomes=10, n_targets=3, genome_len=60_000, n_contigs=4, seed=3, n_runs_every=5,
                            skew=True),
}


def synth_paths(spec: SynthSpec, out_dir: Path):
    return write_set(spec, out_dir)
