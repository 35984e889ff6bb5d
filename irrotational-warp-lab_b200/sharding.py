"""Slab decomposition of the 3D volume along axis 0 and the shard-count-independent reduction.

The reference has no multi-device path ("Parallelism strategies: none"); this is
the B200-side design for BASELINE config 4 (n = 1024^3 on 1/2/4/8 GPUs):

* the index space [0, n0) of axis 0 (the slowest axis of the C-order [i, j, k] volume, so every
  rank's rho slab is one contiguous block) is cut into contiguous slabs whose boundaries are
  multiples of ``FLUSH_PLANES``;
* each rank evaluates its slab with the fused kernel; the 2-plane phi halo is *recomputed*
  analytically from the global coordinates (no halo exchange);
* the only collective is one all-gather of the per-flush-block partial table
  (``ceil(n0/16) x 30`` doubles, 15 KB at n = 1024), after which every rank adds the rows in the
  same fixed order -- so the energies are bit-identical for 1, 2, 4 or 8 ranks.

The preferred decomposition for 2, 4, 8 or 16 ranks is by *interleaved tile shards* instead: rank r evaluates
every P-th tile of the (j, k) plane along the whole of axis 0 (no halo along axis 0, and every rank gets the same
mix of face, wall and shell tiles).  The tile sums of a flush block are 16 chains combined by a fixed binary tree;
a rank owns whole chains, so all-gathering the chain sums (245 KB at n = 1024) and applying the tree on every rank
again gives bit-identical energies for 1, 2, 4, 8 or 16 ranks (``tile_shard``, ``all_gather_chains``,
``combine_chains_host``).

Pure host logic; imports neither CUDA nor the oracle, and is exercised on CPU with gloo.
"""
from __future__ import annotations

import math

import numpy as np

from ._capi import FLUSH_PLANES, MAX_SHELLS, ROW_WIDTH, TILE_CHAINS

# row layout (include/iwb200.h): e_pos e_neg cnt_pos cnt_neg cnt_zero cnt_nan | shell_pos[8] shell_neg[8] shell_cnt[8]
INT_SLOTS = np.array([2, 3, 4, 5] + list(range(6 + 2 * MAX_SHELLS, ROW_WIDTH)))
FLOAT_SLOTS = np.array([q for q in range(ROW_WIDTH) if q not in set(INT_SLOTS.tolist())])


def num_flush_blocks(n0: int) -> int:
    return (n0 + FLUSH_PLANES - 1) // FLUSH_PLANES


def slab_bounds(n0: int, world_size: int) -> list[tuple[int, int]]:
    """Contiguous [i_begin, i_end) per rank; boundaries are multiples of FLUSH_PLANES.
    Ranks beyond the number of flush blocks get an empty slab (i_begin == i_end)."""
    nfb = num_flush_blocks(n0)
    out = []
    for r in range(world_size):
        b0 = (nfb * r) // world_size
        b1 = (nfb * (r + 1)) // world_size
        out.append((min(b0 * FLUSH_PLANES, n0), min(b1 * FLUSH_PLANES, n0)))
    return out


def rows_of_slab(i_begin: int, i_end: int) -> tuple[int, int]:
    """Flush-block rows [r0, r1) of the table that a slab writes."""
    if i_end <= i_begin:
        return (i_begin // FLUSH_PLANES, i_begin // FLUSH_PLANES)
    return (i_begin // FLUSH_PLANES, (i_end + FLUSH_PLANES - 1) // FLUSH_PLANES)


def reduce_table_host(table: np.ndarray, nshell: int, n_points: int | None = None) -> dict:
    """Fixed-order (row 0, 1, 2, ...) sum of a [nrows, ROW_WIDTH] float64 table whose integer slots
    hold int64 bit patterns -- the host twin of ``k_reduce_rows`` (csrc/energy3d.cuh)."""
    table = np.ascontiguousarray(table, dtype=np.float64).reshape(-1, ROW_WIDTH)
    acc = np.zeros(ROW_WIDTH, dtype=np.float64)
    iacc = np.zeros(ROW_WIDTH, dtype=np.int64)
    ints = table.view(np.int64)
    for r in range(table.shape[0]):          # sequential on purpose: same order as the device kernel
        acc[FLOAT_SLOTS] += table[r, FLOAT_SLOTS]
        iacc[INT_SLOTS] += ints[r, INT_SLOTS]
    out = {
        "e_pos": float(acc[0]), "e_neg": float(acc[1]),
        "cnt_pos": int(iacc[2]), "cnt_neg": int(iacc[3]), "cnt_zero": int(iacc[4]),
        "shell_pos": [float(acc[6 + s]) for s in range(nshell)],
        "shell_neg": [float(acc[6 + MAX_SHELLS + s]) for s in range(nshell)],
        "shell_cnt": [int(iacc[6 + 2 * MAX_SHELLS + s]) for s in range(nshell)],
    }
    if n_points is not None:
        out["cnt_nan"] = int(n_points) - out["cnt_pos"] - out["cnt_neg"] - out["cnt_zero"]
        apply_reference_nonfinite(out)
    return out


def apply_reference_nonfinite(res: dict) -> dict:
    """Host twin of ``apply_reference_nonfinite`` (csrc/host_common.h): the reference multiplies by its masks
    (``e * (e > 0)``, reproduce_rodal_exact.py:57-61, 220-224), so NaN * 0 and inf * 0 poison every sum; the
    kernels select instead, and this puts the reference's NaNs back.  In place; needs ``cnt_nan``."""
    nan = float("nan")
    any_nan = res["cnt_nan"] > 0
    pinf, ninf = math.isinf(res["e_pos"]), math.isinf(res["e_neg"])
    if not (any_nan or pinf or ninf):
        return res
    if any_nan or ninf:
        res["e_pos"] = nan
    if any_nan or pinf:
        res["e_neg"] = nan
    for s in range(len(res["shell_pos"])):
        if any_nan or ninf or (pinf and not math.isinf(res["shell_pos"][s])):
            res["shell_pos"][s] = nan
        if any_nan or pinf or (ninf and not math.isinf(res["shell_neg"][s])):
            res["shell_neg"][s] = nan
    return res


def pack_row(e_pos, e_neg, cnt_pos, cnt_neg, cnt_zero, shell_pos=(), shell_neg=(), shell_cnt=()) -> np.ndarray:
    """Build one table row from plain numbers (used by tests that emulate ranks on CPU)."""
    row = np.zeros(ROW_WIDTH, dtype=np.float64)
    irow = row.view(np.int64)
    row[0], row[1] = e_pos, e_neg
    irow[2], irow[3], irow[4] = cnt_pos, cnt_neg, cnt_zero
    for s, v in enumerate(shell_pos):
        row[6 + s] = v
    for s, v in enumerate(shell_neg):
        row[6 + MAX_SHELLS + s] = v
    for s, v in enumerate(shell_cnt):
        irow[6 + 2 * MAX_SHELLS + s] = v
    return row


def all_gather_table(local_rows: "torch.Tensor", n0: int, world_size: int, group=None) -> "torch.Tensor":
    """All-gather of the per-rank row blocks into the full [nfb, ROW_WIDTH] table.

    ``local_rows`` is this rank's [r1-r0, ROW_WIDTH] float64 tensor (CUDA with NCCL, CPU with gloo).
    Slabs may own different numbers of rows, so blocks are padded to the largest and trimmed after.
    """
    import torch
    import torch.distributed as dist

    bounds = slab_bounds(n0, world_size)
    counts = [rows_of_slab(a, b)[1] - rows_of_slab(a, b)[0] for a, b in bounds]
    if min(counts) == max(counts) and counts[0] > 0:
        # equal slabs (the usual case): the gathered buffer IS the table -- one collective, no extra kernels
        full = torch.empty((world_size * counts[0], ROW_WIDTH), dtype=torch.float64, device=local_rows.device)
        dist.all_gather_into_tensor(full.view(-1), local_rows.contiguous().view(-1), group=group)
        return full
    width = max(max(counts), 1)
    padded = torch.zeros((width, ROW_WIDTH), dtype=torch.float64, device=local_rows.device)
    if local_rows.numel():
        padded[: local_rows.shape[0]] = local_rows
    gathered = torch.empty((world_size, width, ROW_WIDTH), dtype=torch.float64, device=local_rows.device)
    dist.all_gather_into_tensor(gathered.view(-1), padded.view(-1), group=group)
    return torch.cat([gathered[r, : counts[r]] for r in range(world_size)], dim=0)


# ---- interleaved tile shards ---------------------------------------------------------------------------------

def tile_shard(world_size: int, rank: int) -> tuple[int, int] | None:
    """(tile_stride, tile_first) of ``rank``, or None when ``world_size`` does not divide the number of chains
    (then the slab decomposition is used)."""
    if world_size >= 1 and TILE_CHAINS % world_size == 0:
        return (world_size, rank)
    return None


def all_gather_chains(local: "torch.Tensor", world_size: int, group=None) -> "torch.Tensor":
    """All-gather of the per-rank chain sums ``[nfb, TILE_CHAINS // P, ROW_WIDTH]`` into the rank-major
    ``[P, nfb, TILE_CHAINS // P, ROW_WIDTH]`` buffer that ``iwb_reduce_chains`` / ``combine_chains_host`` read."""
    import torch
    import torch.distributed as dist

    full = torch.empty((world_size,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(full.view(-1), local.contiguous().view(-1), group=group)
    return full


def combine_chains_host(gathered: np.ndarray) -> np.ndarray:
    """Host twin of ``k_combine_chains`` (csrc/energy3d.cuh): ``[P, nfb, TILE_CHAINS // P, ROW_WIDTH]`` chain sums ->
    the ``[nfb, ROW_WIDTH]`` table.  Chain c is entry ``c // P`` of rank ``c % P``; the chains are combined by the
    binary tree s[c] += s[c + h], h = 8, 4, 2, 1 -- float slots as float64 additions, integer slots as int64."""
    g = np.ascontiguousarray(gathered, dtype=np.float64)
    P, nfb, nlc, width = g.shape
    assert width == ROW_WIDTH and P * nlc == TILE_CHAINS
    s = np.empty((TILE_CHAINS, nfb, ROW_WIDTH), dtype=np.float64)
    for c in range(TILE_CHAINS):
        s[c] = g[c % P, :, c // P, :]
    si = s.view(np.int64)
    h = TILE_CHAINS // 2
    while h > 0:
        for c in range(h):
            s[c][:, FLOAT_SLOTS] = s[c][:, FLOAT_SLOTS] + s[c + h][:, FLOAT_SLOTS]
            si[c][:, INT_SLOTS] = si[c][:, INT_SLOTS] + si[c + h][:, INT_SLOTS]
        h //= 2
    return s[0].copy()
