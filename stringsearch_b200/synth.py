"""Synthetic inputs of BASELINE.md / SURVEY.md section 8(d) (numpy default_rng, PCG64)."""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def acgt(n: int, seed: int) -> np.ndarray:
    """C1 / C4 / C5 text: uniform ACGT."""
    return _ACGT[np.random.default_rng(seed).integers(0, 4, n, dtype=np.uint8)]


def random_bytes(n: int, seed: int) -> np.ndarray:
    """C2: uniform random bytes."""
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8)


def repetitive(n: int, seed: int, period: int = 1000, mutation_rate: float = 1e-3) -> np.ndarray:
    """C3: `period` random bytes tiled to n, then n*mutation_rate random positions overwritten
    with random bytes (forces ~log2(period / mutation spacing) + more doubling rounds)."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, period, dtype=np.uint8)
    reps = (n + period - 1) // period
    x = np.tile(base, reps)[:n].copy()
    k = int(n * mutation_rate)
    if k:
        pos = rng.integers(0, n, k)
        x[pos] = rng.integers(0, 256, k, dtype=np.uint8)
    return x


def patterns_from_text(text: np.ndarray, q: int, m: int, seed: int, alphabet: np.ndarray = _ACGT):
    """C5 pattern batch: even ids are text[o:o+m] for random o (hits), odd ids are random
    m-mers over `alphabet` (mostly misses).  -> (flat uint8 [q*m], offsets uint64 [q+1])."""
    rng = np.random.default_rng(seed)
    n = text.size
    pats = alphabet[rng.integers(0, alphabet.size, (q, m), dtype=np.uint8)]
    hits = np.arange(0, q, 2)
    o = rng.integers(0, max(1, n - m), hits.size)
    idx = o[:, None] + np.arange(m)[None, :]
    pats[hits] = text[np.minimum(idx, n - 1)]
    off = (np.arange(q + 1, dtype=np.uint64) * np.uint64(m)).astype(np.uint64)
    return np.ascontiguousarray(pats.reshape(-1)), off


WORKLOADS = {
    "acgt_4M": lambda: acgt(4 << 20, 1),
    "rand_256M": lambda: random_bytes(1 << 28, 2),
    "rep_1G": lambda: repetitive(1 << 30, 3),
}
