"""Multi-GPU host logic: one process per GPU, contiguous row-range partitioning,
no collective for filter/project (outputs concatenate in rank order), one partial-aggregate merge
for aggregates.  The merge algebra is written once here in numpy so the world_size>1 behaviour can
be tested on CPU over gloo; on GPUs the same algebra runs inside libdfgpu.so (aggregate.cu k_merge,
api.cu agg_exchange_impl over NCCL)."""
import numpy as np

from . import _abi as A


def row_range(rank, world, nrows):
    """GPU g of G owns rows [g*ceil(N/G), min(N, (g+1)*ceil(N/G)))."""
    per = -(-nrows // world)
    lo = min(nrows, rank * per)
    return lo, min(nrows, lo + per)


def shard(arrays, rank, world):
    lo, hi = row_range(rank, world, len(arrays[0]))
    return [a[lo:hi] for a in arrays]


def merge_partials(partials, funcs):
    """partials: list (one per rank) of [key_col, agg_col...] with unique keys per rank;
    funcs: DFGPU_AGG_* per aggregate column.  Returns merged [keys, aggs...] sorted by key."""
    keys = np.concatenate([p[0] for p in partials])
    uk, inv = np.unique(keys, return_inverse=True)
    out = [uk]
    for j, f in enumerate(funcs):
        vals = np.concatenate([p[1 + j] for p in partials])
        if f in (A.AGG_SUM, A.AGG_COUNT):
            acc = np.zeros(len(uk), dtype=vals.dtype)
            np.add.at(acc, inv, vals)
        elif f == A.AGG_MIN:
            acc = np.full(len(uk), np.inf if vals.dtype.kind == "f" else np.iinfo(vals.dtype).max, dtype=vals.dtype)
            np.minimum.at(acc, inv, vals)
        else:
            acc = np.full(len(uk), -np.inf if vals.dtype.kind == "f" else np.iinfo(vals.dtype).min, dtype=vals.dtype)
            np.maximum.at(acc, inv, vals)
        out.append(acc)
    return out


def concat_in_rank_order(parts):
    """filter/project: rank-ordered concatenation preserves global row order."""
    return [np.concatenate([p[i] for p in parts]) for i in range(len(parts[0]))]
