"""Deterministic synthetic dgCMatrix generators (host side, numpy).

The reference benchmarks on ``Matrix::rsparsematrix`` output (reference README.md:33-38,
vignettes/Documentation.Rmd:377,425).  R is not available, so SURVEY.md section 8(d)
defines integer-exact stand-ins: every quantity is a pure function of
``(seed, column, k)`` through a 64-bit mixing hash, so the CUDA generator in
``csrc/synth.cu`` (used for the >1e8-nnz configs, generated straight into HBM) and this
numpy version produce bit-identical ``i/p/x``.  ``tests/test_synth.py`` checks that.

Recipe (all integer arithmetic is modulo 2**64):
  * column length  L_c = Q[j] + ((Q[j+1]-Q[j]) * f >> 16), (j, f) = 12+16 bits of
    h3(seed, c, 0); Q is a 4097-entry quantile table built on the host (binomial-normal
    for "uniform", Pareto alpha=1.5 for "power-law"); a column is forced empty when
    h3(seed, c, 1) % 1000 < empty_permille.
  * rows: the row range is cut into bands (one band = uniform; geometric bands with equal
    weight = scRNA-like row popularity).  Band j receives c_j = min(size_j, L*w_j // W)
    entries, placed by integer-stratified sampling: the q-th of c_j entries falls in
    [q*s//c_j, (q+1)*s//c_j) + lo_j, offset chosen by h3(seed, c, 2+2k) — strictly
    ascending and duplicate-free by construction (the dgCMatrix invariant).
  * values: sum of four 8-bit hash fields, centred, /100 -> two-decimal approximately
    normal values of mixed sign (mimics rsparsematrix's signif(rnorm, 2)); exact in
    IEEE-754 on both CPU and GPU.
"""
from __future__ import annotations

import dataclasses
from statistics import NormalDist

import numpy as np

TABLE = 4096
_M1 = np.uint64(0x9E3779B97F4A7C15)
_M2 = np.uint64(0xBF58476D1CE4E5B9)
_M3 = np.uint64(0x94D049BB133111EB)
_KA = np.uint64(0xD6E8FEB86659FD93)
_KB = np.uint64(0xA0761D6478BD642F)


def mix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + _M1
        z = (z ^ (z >> np.uint64(30))) * _M2
        z = (z ^ (z >> np.uint64(27))) * _M3
        return z ^ (z >> np.uint64(31))


def h3(seed, a, b):
    with np.errstate(over="ignore"):
        a = np.asarray(a, dtype=np.uint64)
        b = np.asarray(b, dtype=np.uint64)
        return mix64(mix64(np.uint64(seed) + a * _KA) ^ (b * _KB))


def value_from_hash(h):
    h = np.asarray(h, dtype=np.uint64)
    m = np.uint64(0xFF)
    s = (h & m) + ((h >> np.uint64(8)) & m) + ((h >> np.uint64(16)) & m) + ((h >> np.uint64(24)) & m)
    return (s.astype(np.int64) - 510).astype(np.float64) / 100.0


@dataclasses.dataclass
class SynthSpec:
    """Everything both generators need; plain integers and small int64 tables."""

    name: str
    nrow: int
    ncol: int
    seed: int
    len_table: np.ndarray  # int64[TABLE+1], non-decreasing, values in [0, nrow]
    empty_permille: int = 0
    band_lo: np.ndarray = None  # int64[K] ascending, band j = [band_lo[j], band_hi[j])
    band_hi: np.ndarray = None
    band_w: np.ndarray = None  # int64[K] positive weights

    def __post_init__(self):
        if self.band_lo is None:
            self.band_lo = np.array([0], dtype=np.int64)
            self.band_hi = np.array([self.nrow], dtype=np.int64)
            self.band_w = np.array([1], dtype=np.int64)
        self.len_table = np.ascontiguousarray(self.len_table, dtype=np.int64)
        self.band_lo = np.ascontiguousarray(self.band_lo, dtype=np.int64)
        self.band_hi = np.ascontiguousarray(self.band_hi, dtype=np.int64)
        self.band_w = np.ascontiguousarray(self.band_w, dtype=np.int64)
        assert self.len_table.shape == (TABLE + 1,)
        assert np.all(np.diff(self.len_table) >= 0)
        assert self.len_table[0] >= 0 and self.len_table[-1] <= self.nrow

    @property
    def n_bands(self) -> int:
        return int(self.band_lo.shape[0])


# ----------------------------------------------------------------------------------------
# per-column lengths (shared by host generator and by tests that check the device one)
# ----------------------------------------------------------------------------------------
def band_counts(spec: SynthSpec, raw_len: np.ndarray) -> np.ndarray:
    """int64[ncol, K]: entries each band receives for a column of raw length raw_len."""
    size = spec.band_hi - spec.band_lo
    wsum = int(spec.band_w.sum())
    share = (raw_len[:, None] * spec.band_w[None, :]) // wsum
    return np.minimum(share, size[None, :])


def column_lengths(spec: SynthSpec, cols: np.ndarray | None = None) -> np.ndarray:
    """Final stored length of each column (after banding), int64."""
    if cols is None:
        cols = np.arange(spec.ncol, dtype=np.uint64)
    cols = np.asarray(cols, dtype=np.uint64)
    return band_counts(spec, _raw_lengths(spec, cols)).sum(axis=1)


def _raw_lengths(spec: SynthSpec, cols: np.ndarray) -> np.ndarray:
    h = h3(spec.seed, cols, 0)
    j = (h >> np.uint64(52)).astype(np.int64)
    f = ((h >> np.uint64(36)) & np.uint64(0xFFFF)).astype(np.int64)
    q0 = spec.len_table[j]
    q1 = spec.len_table[j + 1]
    raw = q0 + (((q1 - q0) * f) >> 16)
    if spec.empty_permille > 0:
        e = (h3(spec.seed, cols, 1) % np.uint64(1000)).astype(np.int64)
        raw = np.where(e < spec.empty_permille, 0, raw)
    return np.clip(raw, 0, spec.nrow)


def generate_host(spec: SynthSpec, col_begin: int = 0, col_end: int | None = None):
    """Return (i int32[nnz], p int32[ncols+1], x float64[nnz]) for columns [col_begin, col_end).

    p is rebased to start at 0, so a column block is itself a valid dgCMatrix with
    ``nrow`` rows — which is exactly one rank's shard under column sharding.
    """
    if col_end is None:
        col_end = spec.ncol
    cols = np.arange(col_begin, col_end, dtype=np.uint64)
    ncols = cols.shape[0]
    raw = _raw_lengths(spec, cols)
    cnt = band_counts(spec, raw)  # [ncols, K]
    lens = cnt.sum(axis=1)
    p64 = np.zeros(ncols + 1, dtype=np.int64)
    np.cumsum(lens, out=p64[1:])
    nnz = int(p64[-1])
    if nnz >= 2**31:
        raise ValueError("nnz exceeds int32 (dgCMatrix limit, reference RcppSparse.h:30)")
    p = p64.astype(np.int32)
    if nnz == 0:
        return np.zeros(0, np.int32), p, np.zeros(0, np.float64)

    col_of = np.repeat(np.arange(ncols, dtype=np.int64), lens)  # local column of each entry
    k = np.arange(nnz, dtype=np.int64) - p64[col_of]  # position inside its column
    # band of each entry: first j with prefix[j+1] > k
    prefix = np.zeros((ncols, spec.n_bands + 1), dtype=np.int64)
    np.cumsum(cnt, axis=1, out=prefix[:, 1:])
    band = np.zeros(nnz, dtype=np.int64)
    for j in range(1, spec.n_bands):
        band += (k >= prefix[col_of, j]).astype(np.int64)
    q = k - prefix[col_of, band]
    c_j = cnt[col_of, band]
    s_j = (spec.band_hi - spec.band_lo)[band]
    b0 = (q * s_j) // c_j
    b1 = ((q + 1) * s_j) // c_j
    gcol = cols[col_of]
    hr = h3(spec.seed, gcol, (2 + 2 * k).astype(np.uint64))
    off = (hr % (b1 - b0).astype(np.uint64)).astype(np.int64)
    rows = spec.band_lo[band] + b0 + off
    hv = h3(spec.seed, gcol, (3 + 2 * k).astype(np.uint64))
    x = value_from_hash(hv)
    return rows.astype(np.int32), p, np.ascontiguousarray(x)


def dense_vector(seed: int, n: int) -> np.ndarray:
    """The SpMV operand v (seed+1), float64[n], same recipe as the values."""
    return value_from_hash(h3(seed + 1, np.arange(n, dtype=np.uint64), 0))


# ----------------------------------------------------------------------------------------
# quantile tables
# ----------------------------------------------------------------------------------------
def _normal_table(mean: float, sd: float, nrow: int) -> np.ndarray:
    nd = NormalDist()
    qs = [(j + 0.5) / (TABLE + 1) for j in range(TABLE + 1)]
    z = np.array([nd.inv_cdf(q) for q in qs])
    t = np.rint(mean + sd * z).astype(np.int64)
    return np.clip(t, 0, nrow)


def _pareto_table(l_min: float, alpha: float, nrow: int) -> np.ndarray:
    u = 1.0 - np.arange(TABLE + 1, dtype=np.float64) / (TABLE + 1)  # (0, 1]
    t = np.floor(l_min * u ** (-1.0 / alpha))
    return np.clip(t, 0, nrow).astype(np.int64)


def _expected_mean(table: np.ndarray) -> float:
    # E over j uniform, f uniform in [0, 65536): interpolation midpoint
    return float(np.mean(table[:-1] + (table[1:] - table[:-1]) * 0.5))


def geometric_bands(nrow: int, n_levels: int):
    """Rows [nrow>>(l+1), nrow>>l) for l = 0..n_levels-1 plus [0, nrow>>n_levels), ascending, equal weight."""
    edges = [0] + [nrow >> l for l in range(n_levels, -1, -1)]
    edges = sorted(set(e for e in edges if e >= 0))
    lo = np.array(edges[:-1], dtype=np.int64)
    hi = np.array(edges[1:], dtype=np.int64)
    keep = hi > lo
    lo, hi = lo[keep], hi[keep]
    return lo, hi, np.ones_like(lo)


def uniform_spec(nrow: int, ncol: int, density: float, seed: int, name: str = "uniform") -> SynthSpec:
    mean = density * nrow
    sd = (mean * (1.0 - density)) ** 0.5
    return SynthSpec(name, nrow, ncol, seed, _normal_table(mean, sd, nrow))


def powerlaw_spec(nrow: int, ncol: int, mean_len: float, seed: int, alpha: float = 1.5,
                  empty_permille: int = 10, row_levels: int = 0, name: str = "powerlaw") -> SynthSpec:
    """Pareto(alpha) column lengths with the scale solved so the expected mean is mean_len."""
    lo_s, hi_s = 1e-3, float(nrow)
    target = mean_len / (1.0 - empty_permille / 1000.0)
    for _ in range(80):
        mid = 0.5 * (lo_s + hi_s)
        if _expected_mean(_pareto_table(mid, alpha, nrow)) < target:
            lo_s = mid
        else:
            hi_s = mid
    table = _pareto_table(hi_s, alpha, nrow)
    spec = SynthSpec(name, nrow, ncol, seed, table, empty_permille)
    if row_levels > 0:
        lo, hi, w = geometric_bands(nrow, row_levels)
        spec = SynthSpec(name, nrow, ncol, seed, table, empty_permille, lo, hi, w)
    return spec


# BASELINE.json configs / SURVEY.md 8(d) "Concrete configs"
def config(name: str, scale: float = 1.0) -> SynthSpec:
    """C1..C4 at full size (scale=1) or with the column count scaled down for parity tests."""
    name = name.upper()
    if name == "C1":
        return uniform_spec(10_000, max(1, int(10_000 * scale)), 0.01, 1001, "C1 10k x 10k uniform d=0.01")
    if name == "C2":
        return uniform_spec(1_000_000, max(1, int(100_000 * scale)), 1e-3, 1002, "C2 1M x 100k uniform d=1e-3")
    if name == "C3":
        # raw Pareto mean 2200: the small popular-row bands cap their share, which leaves ~1500 stored per column
        return powerlaw_spec(30_000, max(1, int(1_000_000 * scale)), 2200.0, 1003, row_levels=10,
                             name="C3 30k x 1M power-law columns, banded row popularity")
    if name == "C4":
        return powerlaw_spec(1 << 20, max(1, int(2_000_000 * scale)), 1000.0, 1004,
                             name="C4 2^20 x 2M power-law columns")
    raise KeyError(name)
