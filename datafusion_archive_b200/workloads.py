"""Seeded synthetic inputs for the BASELINE.json configs.  Deterministic from the
seed alone, so the CPU oracle, every test and every GPU rank regenerate identical data."""
import numpy as np

from . import _abi as A
from .expr import AggregateFunction, col, lit

GOLDEN = 0x9E3779B97F4A7C15


def mix_keys(k_raw):
    """Bijective 64-bit scramble so GROUP BY keys are not a dense range (a direct-index table would
    be cheating)."""
    with np.errstate(over="ignore"):
        return (k_raw.astype(np.uint64) * np.uint64(GOLDEN) + np.uint64(0x1234567)).view(np.int64)


def c2(n, seed=42, out=None):
    """C2: SELECT a FROM t WHERE a > 0.5, a ~ U[0,1) f64."""
    a = np.random.default_rng(seed).random(n, out=out)
    return [a], (col(0) > lit(0.5)), [col(0)]


def c3(n, seed=42):
    """C3: SELECT a+b, a*b FROM t WHERE b < a over 4 f64 columns (c, d unreferenced)."""
    cols = [np.random.default_rng(seed + i).random(n) for i in range(4)]
    return cols, (col(1) < col(0)), [col(0) + col(1), col(0) * col(1)]


def c4(n, nkeys=100_000, seed=46):
    """C4: SELECT k, SUM(v), COUNT(v) FROM t GROUP BY k; k Int64 (scrambled), v ~ U[0,1) f64."""
    k_raw = np.random.default_rng(seed).integers(0, nkeys, n, dtype=np.int64)
    v = np.random.default_rng(seed + 1).random(n)
    return [mix_keys(k_raw), v], [col(0)], [AggregateFunction("sum", col(1)), AggregateFunction("count", col(1))], k_raw


def c5(n, nkeys=1_000_000, seed=46):
    """C5: SELECT k, MIN(v), MAX(v), SUM(v) FROM t GROUP BY k."""
    k_raw = np.random.default_rng(seed).integers(0, nkeys, n, dtype=np.int64)
    v = np.random.default_rng(seed + 1).random(n)
    aggs = [AggregateFunction("min", col(1)), AggregateFunction("max", col(1)), AggregateFunction("sum", col(1))]
    return [mix_keys(k_raw), v], [col(0)], aggs, k_raw


SCHEMA_F64 = [A.FLOAT64]
