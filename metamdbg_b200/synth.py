"""Synthetic metagenome reads from a counter-based generator.

The same arithmetic is implemented on the device by ``mdbg_synth_fill_reads``
(metamdbg_b200/csrc/synth.cu) so that full-size benchmark inputs can be
created directly in HBM while the tests build the identical bytes with numpy.

Model (simplified to substitution errors):
  * G "virtual" genomes; base at virtual position q is ``mix64(q + seed) & 3``
    mapped through "ACGT" -- nothing is stored.
  * read r: genome drawn from a cumulative abundance table, uniform start,
    random strand, length ~ mean + sd * (sum of 4 uniforms - 2) * sqrt(3),
    clipped to [min_len, genome_len].
  * each base is substituted with probability ``err`` by one of the three
    other letters.
"""
from __future__ import annotations

import dataclasses

import numpy as np

M1 = np.uint64(0xBF58476D1CE4E5B9)
M2 = np.uint64(0x94D049BB133111EB)
GOLD = np.uint64(0x9E3779B97F4A7C15)
LETTERS = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP_IDX = np.array([3, 2, 1, 0], dtype=np.uint8)  # A<->T, C<->G in "ACGT" index space


def mix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = (x ^ (x >> np.uint64(30))) * M1
        x = (x ^ (x >> np.uint64(27))) * M2
        x = x ^ (x >> np.uint64(31))
    return x


@dataclasses.dataclass
class ReadSet:
    """Host-side description of a synthetic read set (no bases)."""

    seed: int
    err_q24: int                 # substitution probability * 2^24
    offsets: np.ndarray          # uint64 [n+1] byte offsets of each read in the concatenated buffer
    vstart: np.ndarray           # uint64 [n] virtual genome position of the read's first (forward) base
    strand: np.ndarray           # uint8  [n] 1 = reverse complement
    n_genomes: int
    genome_len: np.ndarray       # uint64 [G]
    index_base: int = 0          # global index of this set's first read (error stream is keyed by global index)

    @property
    def n_reads(self) -> int:
        return int(self.offsets.shape[0] - 1)

    @property
    def n_bases(self) -> int:
        return int(self.offsets[-1])

    def shard(self, rank: int, world: int) -> "ReadSet":
        """Contiguous record range for one rank."""
        n = self.n_reads
        lo, hi = n * rank // world, n * (rank + 1) // world
        return self.subset(lo, hi)

    def subset(self, lo: int, hi: int) -> "ReadSet":
        off = self.offsets[lo:hi + 1] - self.offsets[lo]
        rs = dataclasses.replace(self, offsets=off.copy(), vstart=self.vstart[lo:hi].copy(),
                                 strand=self.strand[lo:hi].copy(), index_base=self.index_base + lo)
        if hasattr(self, "lengths"):
            rs.lengths = self.lengths[lo:hi].copy()  # type: ignore[attr-defined]
        return rs


def make_readset(n_reads: int, mean_len: int, *, seed: int = 1, n_genomes: int = 1,
                 genome_len_range=(2_000_000, 6_000_000), err: float = 0.001,
                 len_sd_frac: float = 0.15, min_len: int = 1000, abundance_sigma: float = 1.0,
                 align: int = 1, sample: int = 0, index_base: int = 0) -> ReadSet:
    """Draw read placements (host, numpy).  ``align`` pads every read start to a
    multiple of ``align`` bytes in the concatenated buffer (padding bytes are
    never part of a read).  ``sample`` > 0 draws another SAMPLE of the same genome
    set (same genomes and bases, its own abundance profile and read placements:
    a co-assembly input, datasetIndex = sample); ``index_base`` is the global
    index of its first read (keys the per-read error stream)."""
    s = np.uint64(seed)
    smp = np.uint64(sample) * np.uint64(0x9E3779B1)
    g = np.arange(n_genomes, dtype=np.uint64)
    with np.errstate(over="ignore"):
        u = mix64(g * GOLD + s + np.uint64(0x1111))
        lo, hi = genome_len_range
        glen = (np.uint64(lo) + u % np.uint64(max(1, hi - lo))).astype(np.uint64)
        gbase = np.concatenate([[np.uint64(0)], np.cumsum(glen)[:-1]]).astype(np.uint64)
        # log-normal abundance via Box-Muller on two hashed uniforms
        u1 = (mix64(g * GOLD + s + np.uint64(0x2222) + smp) >> np.uint64(11)).astype(np.float64) / 2.0**53
        u2 = (mix64(g * GOLD + s + np.uint64(0x3333) + smp) >> np.uint64(11)).astype(np.float64) / 2.0**53
        z = np.sqrt(-2.0 * np.log(np.maximum(u1, 1e-300))) * np.cos(2 * np.pi * u2)
        w = np.exp(abundance_sigma * z) * glen.astype(np.float64)
        cum = np.cumsum(w) / np.sum(w)

        r = np.arange(n_reads, dtype=np.uint64) + np.uint64(index_base)
        ug = (mix64(r * GOLD + s + np.uint64(0x4444)) >> np.uint64(11)).astype(np.float64) / 2.0**53
        gi = np.minimum(np.searchsorted(cum, ug, side="right"), n_genomes - 1)
        ul = mix64(r * GOLD + s + np.uint64(0x5555))
        parts = [((ul >> np.uint64(16 * i)) & np.uint64(0xFFFF)).astype(np.float64) / 65536.0 for i in range(4)]
        zl = (parts[0] + parts[1] + parts[2] + parts[3] - 2.0) * np.sqrt(3.0)
        ln = np.rint(mean_len * (1.0 + len_sd_frac * zl)).astype(np.int64)
        ln = np.clip(ln, min_len, None)
        ln = np.minimum(ln, glen[gi].astype(np.int64))
        us = mix64(r * GOLD + s + np.uint64(0x6666))
        span = (glen[gi].astype(np.int64) - ln + 1).astype(np.uint64)
        st = us % span
        strand = (mix64(r * GOLD + s + np.uint64(0x7777)) & np.uint64(1)).astype(np.uint8)
    padded = (ln + (align - 1)) // align * align
    offsets = np.zeros(n_reads + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(padded).astype(np.uint64)
    # offsets delimit padded slots; true length kept separately
    rs = ReadSet(seed=seed, err_q24=int(round(err * (1 << 24))), offsets=offsets,
                 vstart=(gbase[gi] + st).astype(np.uint64), strand=strand,
                 n_genomes=n_genomes, genome_len=glen, index_base=index_base)
    rs.lengths = ln.astype(np.uint64)  # type: ignore[attr-defined]
    if align == 1:
        assert np.array_equal(np.diff(offsets.astype(np.int64)), ln)
    return rs


def fill_reads(rs: ReadSet, lo: int = 0, hi: int | None = None) -> tuple[np.ndarray, np.ndarray]:
    """Materialise reads [lo, hi) as (ASCII uint8 buffer, uint64 offsets[n+1]) with
    tight (unpadded) offsets."""
    hi = rs.n_reads if hi is None else hi
    lengths = getattr(rs, "lengths", None)
    if lengths is None:
        lengths = np.diff(rs.offsets)
    lens = lengths[lo:hi].astype(np.int64)
    offs = np.zeros(hi - lo + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(lens).astype(np.uint64)
    out = np.empty(int(offs[-1]), dtype=np.uint8)
    s = np.uint64(rs.seed)
    for j, r in enumerate(range(lo, hi)):
        n = int(lens[j])
        rid = np.uint64(rs.index_base + r)
        i = np.arange(n, dtype=np.uint64)
        with np.errstate(over="ignore"):
            if rs.strand[r]:
                q = rs.vstart[r] + np.uint64(n - 1) - i
            else:
                q = rs.vstart[r] + i
            b = (mix64(q + s) & np.uint64(3)).astype(np.uint8)
            if rs.strand[r]:
                b = COMP_IDX[b]
            e = mix64((rid * GOLD) ^ (i + s * np.uint64(0x632BE5AB)))
            sub = (e & np.uint64(0xFFFFFF)) < np.uint64(rs.err_q24)
            shift = (np.uint64(1) + (e >> np.uint64(24)) % np.uint64(3)).astype(np.uint8)
            b = np.where(sub, (b + shift) & np.uint8(3), b).astype(np.uint8)
        out[int(offs[j]):int(offs[j + 1])] = LETTERS[b]
    return out, offs


def pack_2bit(bases: np.ndarray, offsets: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """ASCII reads (letters A, C, G, T only) -> (uint32 words, uint64 word offset per read) in the 2-bit layout of
    include/mdbg_b200.h: 16 bases per word, base j at bits [2j, 2j+1], code (c >> 1) & 3."""
    n = len(offsets) - 1
    lens = np.diff(offsets.astype(np.int64))
    words_per = (lens + 15) // 16
    woff = np.zeros(n + 1, dtype=np.uint64)
    woff[1:] = np.cumsum(words_per).astype(np.uint64)
    out = np.zeros(int(woff[-1]) + 1, dtype=np.uint32)
    codes = ((bases >> np.uint8(1)) & np.uint8(3)).astype(np.uint32)
    for r in range(n):
        c = codes[int(offsets[r]):int(offsets[r + 1])]
        pad = (-len(c)) % 16
        if pad:
            c = np.concatenate([c, np.zeros(pad, np.uint32)])
        c = c.reshape(-1, 16)
        out[int(woff[r]):int(woff[r]) + len(c)] = (c << (2 * np.arange(16, dtype=np.uint32))).sum(axis=1, dtype=np.uint64).astype(np.uint32)
    return out[:int(woff[-1])], woff[:n].copy()
