"""Deterministic synthetic bacterial-genome sets.

Root ancestor = iid uniform ACGT; two clade ancestors (target / non-target) = root with 3 % iid
substitutions; genome g = its clade ancestor with ``mu`` iid substitutions (seed = base_seed + g),
cut into ``n_contigs`` records at fixed breakpoints.  Every 10th genome gets ten 100-bp ``N`` runs
and lower-case soft-masking over 5 % to exercise the validity path.  Targets come first.

Genomes are produced as ASCII ``uint8`` arrays; :func:`write_fasta` lays them out as 80-column
plain FASTA (what the reference CPU path reads) and :func:`seqwin_b200.pack.pack_records` turns
the same arrays into the 2-bit device layout.
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclass(frozen=True)
class SynthSpec:
    n_genomes: int = 8
    n_targets: int = 2
    genome_len: int = 5_000_000
    n_contigs: int = 50
    mu: float = 0.01
    clade_div: float = 0.03
    seed: int = 42
    n_runs_every: int = 10          # every Nth genome gets N runs + soft masking (0 = never)
    skew: bool = False              # config C5: near-clonal + repeated element + low-complexity


def _substitute(codes: np.ndarray, rate: float, rng: np.random.Generator) -> np.ndarray:
    """iid substitutions: each chosen site moves to one of the 3 other bases."""
    out = codes.copy()
    n_sub = rng.binomial(len(codes), rate)
    if n_sub:
        sites = rng.integers(0, len(codes), n_sub)
        out[sites] = (out[sites] + rng.integers(1, 4, n_sub).astype(np.uint8)) & 3
    return out


class SynthSet:
    """Lazily generates genome ``g`` of a :class:`SynthSpec` (codes 0..3, then ASCII)."""

    def __init__(self, spec: SynthSpec):
        self.spec = spec
        rng = np.random.default_rng(spec.seed)
        root = rng.integers(0, 4, spec.genome_len, dtype=np.uint8)
        self._clade = [_substitute(root, spec.clade_div, rng), _substitute(root, spec.clade_div, rng)]
        cuts = np.sort(rng.choice(np.arange(1000, max(1001, spec.genome_len - 1000)),
                                  size=max(0, spec.n_contigs - 1), replace=False)) \
            if spec.n_contigs > 1 and spec.genome_len > 4000 else np.array([], dtype=np.int64)
        self.breaks = np.concatenate([[0], cuts, [spec.genome_len]]).astype(np.int64)
        if spec.skew:
            self._element = rng.integers(0, 4, 5000, dtype=np.uint8)

    @property
    def is_targets(self) -> np.ndarray:
        return np.arange(self.spec.n_genomes) < self.spec.n_targets

    def genome_ascii(self, g: int) -> np.ndarray:
        sp = self.spec
        rng = np.random.default_rng(sp.seed + 1 + g)
        anc = self._clade[0 if g < sp.n_targets else 1]
        codes = _substitute(anc, 1e-4 if sp.skew else sp.mu, rng)
        if sp.skew:
            L = sp.genome_len
            for s in rng.integers(0, max(1, L - 5000), 200 if L >= 1_000_000 else 4):
                codes[s:s + 5000] = self._element[: max(0, min(5000, L - s))]
            n_tract = int(0.02 * L / 200)
            for s in rng.integers(0, max(1, L - 200), n_tract):
                unit = rng.integers(0, 4, rng.integers(1, 3), dtype=np.uint8)
                codes[s:s + 200] = np.resize(unit, 200)[: max(0, min(200, L - s))]
        seq = _ACGT[codes]
        if sp.n_runs_every and g % sp.n_runs_every == sp.n_runs_every - 1:
            L = sp.genome_len
            for s in rng.integers(0, max(1, L - 100), 10):
                seq[s:s + 100] = ord("N")
            n_mask = max(1, int(0.05 * L / 500))
            for s in rng.integers(0, max(1, L - 500), n_mask):
                seq[s:s + 500] |= 0x20          # lower-case soft mask
        return seq

    def records(self, g: int) -> list[tuple[str, np.ndarray]]:
        seq = self.genome_ascii(g)
        return [(f"g{g}_c{c}", seq[self.breaks[c]:self.breaks[c + 1]])
                for c in range(len(self.breaks) - 1)]


def write_fasta(path: Path, records: list[tuple[str, np.ndarray]], width: int = 80) -> None:
    """80-column plain FASTA, built with numpy (no per-line Python loop)."""
    chunks: list[bytes] = []
    for rid, seq in records:
        n = len(seq)
        n_full, rem = divmod(n, width)
        body = np.empty(n + n_full + (1 if rem else 0), dtype=np.uint8)
        if n_full:
            blk = body[: n_full * (width + 1)].reshape(n_full, width + 1)
            blk[:, :width] = seq[: n_full * width].reshape(n_full, width)
            blk[:, width] = 10
        if rem:
            body[n_full * (width + 1): -1] = seq[n_full * width:]
            body[-1] = 10
        chunks.append(b">" + rid.encode() + b"\n")
        chunks.append(body.tobytes())
    Path(path).write_bytes(b"".join(chunks))


def write_set(spec: SynthSpec, out_dir: Path) -> tuple[list[Path], np.ndarray]:
    out_dir = Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    ss = SynthSet(spec)
    paths = []
    for g in range(spec.n_genomes):
        p = out_dir / f"g{g:05d}.fasta"
        write_fasta(p, ss.records(g))
        paths.append(p)
    return paths, ss.is_targets
