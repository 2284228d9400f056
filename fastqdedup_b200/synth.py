"""Deterministic synthetic UMI/key workloads.

Not part of the clustering path: this only manufactures the inputs the five
BASELINE.json configs name, for ``bench.py`` and the parity tests.  Everything is
derived from ``numpy.random.default_rng`` seeded with (config seed, chunk index), so any
rank can generate any contiguous range of reads of the same global data set.

Recipe: M molecules with uniformly random ACGT keys; family weights ~ lognormal(0, 1);
N reads drawn with those weights; 0.5 %/base substitutions; 0.1 % of bases replaced by
``N``; optionally 0.1 %/base single-base indels (key re-cut to L from a longer
molecule) and a fraction of reads truncated below L; qualities ``'I'`` (Q40) except the
cfg-3 mix (90 % Q40, 5 % all ``'?'`` = Q30 -- mean 0.0010000000000000002 > 0.001 so
discarded -- and 5 % Q40 with 10 % of the bases at Q10..Q20).
"""
from dataclasses import dataclass, replace

import numpy as np

CHUNK = 1 << 20
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclass(frozen=True)
class SynthConfig:
    name: str
    key_length: int
    n_reads: int
    n_molecules: int
    seed: int
    max_distance: int = 1
    use_edit_distance: bool = False
    method: str = "directional"
    max_average_error_rate: float = 1.0   # 1.0 == -E (filter off)
    sub_rate: float = 0.005
    n_rate: float = 0.001
    indel_rate: float = 0.0
    truncate_frac: float = 0.0
    quality_mix: bool = False

    def scaled(self, n_reads: int) -> "SynthConfig":
        """Same recipe at another size (molecule count scales with it)."""
        ratio = self.n_molecules / self.n_reads
        return replace(self, n_reads=int(n_reads), n_molecules=max(1, int(n_reads * ratio)))


# BASELINE.json configs 1..5
CONFIGS = {
    "cfg1": SynthConfig("cfg1", 12, 1_000_000, 50_000, 1),
    "cfg2": SynthConfig("cfg2", 36, 10_000_000, 2_000_000, 2, method="adjacency"),
    "cfg3": SynthConfig("cfg3", 48, 20_000_000, 4_000_000, 3, max_distance=2,
                        max_average_error_rate=0.001, quality_mix=True),
    "cfg4": SynthConfig("cfg4", 24, 5_000_000, 1_000_000, 4, use_edit_distance=True,
                        indel_rate=0.001),
    "cfg5": SynthConfig("cfg5", 36, 100_000_000, 20_000_000, 5),
}

_MOL_PAD = 8  # molecules are generated a little longer than L so indel reads can be re-cut


class SynthSource:
    """Holds the molecule pool and family weights of one config; generates read ranges."""

    def __init__(self, cfg: SynthConfig):
        self.cfg = cfg
        rng = np.random.default_rng([cfg.seed, 0xC0FFEE])
        self.mol = rng.integers(0, 4, size=(cfg.n_molecules, cfg.key_length + _MOL_PAD),
                                dtype=np.uint8)
        w = rng.lognormal(0.0, 1.0, size=cfg.n_molecules)
        self.cumw = np.cumsum(w / w.sum())
        self.cumw[-1] = 1.0

    def _chunk(self, ci: int, lo: int, hi: int):
        """Reads [ci*CHUNK + lo, ci*CHUNK + hi) -- always generates the whole chunk's random
        stream so a sub-range equals the corresponding slice of the full chunk."""
        cfg = self.cfg
        L = cfg.key_length
        base = ci * CHUNK
        n = min(CHUNK, cfg.n_reads - base)
        rng = np.random.default_rng([cfg.seed, ci])
        ids = np.searchsorted(self.cumw, rng.random(n), side="right")
        np.minimum(ids, cfg.n_molecules - 1, out=ids)
        codes = self.mol[ids, :L].copy()                       # [n, L] in 0..3
        flat = codes.reshape(-1)
        # substitutions
        k = rng.binomial(n * L, cfg.sub_rate)
        pos = rng.integers(0, n * L, size=k)
        shift = rng.integers(1, 4, size=k, dtype=np.uint8)
        flat[pos] = (flat[pos] + shift) & 3
        keys = _ACGT[codes]                                     # ASCII
        lens = None
        # single-base indels: re-cut from the longer molecule
        if cfg.indel_rate > 0.0:
            k = rng.binomial(n * L, cfg.indel_rate)
            ipos = rng.integers(0, n * L, size=k)
            kind = rng.integers(0, 2, size=k)
            newb = rng.integers(0, 4, size=k, dtype=np.uint8)
            for p, kd, nb in zip(ipos.tolist(), kind.tolist(), newb.tolist()):
                r, c = divmod(p, L)
                row = keys[r].copy()
                if kd == 0:      # deletion: shift left, refill the tail from the molecule
                    row[c:L - 1] = keys[r, c + 1:L]
                    row[L - 1] = _ACGT[self.mol[ids[r], L]]
                else:            # insertion: shift right, last base falls off
                    row[c + 1:L] = keys[r, c:L - 1]
                    row[c] = _ACGT[nb]
                keys[r] = row
        # N bases
        k = rng.binomial(n * L, cfg.n_rate)
        npos = rng.integers(0, n * L, size=k)
        keys.reshape(-1)[npos] = ord("N")
        # truncated reads (shorter than the check length => shorter key)
        if cfg.truncate_frac > 0.0:
            lens = np.full(n, L, dtype=np.uint32)
            tr = rng.random(n) < cfg.truncate_frac
            lens[tr] = rng.integers(max(1, L - 6), L, size=int(tr.sum()))
        quals = None
        if cfg.quality_mix:
            quals = np.full((n, L), ord("I"), dtype=np.uint8)
            cls = rng.random(n)
            quals[cls < 0.05] = ord("?")
            mixed = np.nonzero((cls >= 0.05) & (cls < 0.10))[0]
            m = rng.random((len(mixed), L)) < 0.10
            low = rng.integers(33 + 10, 33 + 21, size=(len(mixed), L), dtype=np.uint8)
            sub = quals[mixed]
            sub[m] = low[m]
            quals[mixed] = sub
        sl = slice(lo, hi)
        return (keys[sl], None if lens is None else lens[sl],
                None if quals is None else quals[sl])

    def reads(self, start: int = 0, stop: int = None):
        """-> (keys uint8 [n, L] ASCII, lens uint32[n] | None, quals uint8 [n, L] | None).
        With ``lens`` given, only the first lens[i] bytes of row i are the key."""
        cfg = self.cfg
        stop = cfg.n_reads if stop is None else min(stop, cfg.n_reads)
        ks, ls, qs = [], [], []
        t = start
        while t < stop:
            ci = t // CHUNK
            lo = t - ci * CHUNK
            hi = min(stop - ci * CHUNK, CHUNK)
            k, l, q = self._chunk(ci, lo, hi)
            ks.append(k), ls.append(l), qs.append(q)
            t = ci * CHUNK + hi
        if not ks:
            L = cfg.key_length
            return (np.zeros((0, L), np.uint8), None,
                    np.zeros((0, L), np.uint8) if cfg.quality_mix else None)
        keys = np.concatenate(ks) if len(ks) > 1 else ks[0]
        lens = None if ls[0] is None else np.concatenate(ls)
        quals = None if qs[0] is None else np.concatenate(qs)
        return np.ascontiguousarray(keys), lens, quals


def to_ragged(keys, lens):
    """Fixed-stride rows + lens -> (flat bytes, uint64 offsets) with the padding removed."""
    n, L = keys.shape
    if lens is None:
        return keys.reshape(-1), np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
    mask = np.arange(L, dtype=np.uint32)[None, :] < lens[:, None]
    off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    return keys[mask], off
