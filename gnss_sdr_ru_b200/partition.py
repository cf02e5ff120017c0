"""Sharding of the two hot paths across ranks (one process per GPU).

* tracking: IF streams are independent -> stream s lives on rank s % world; no
  data-path collective.
* acquisition: the (sv, Doppler bin) rows are independent given the (small) record on every rank
  -> row r = sv_index*n_bins + bin is computed by rank r % world (csrc/acq.cu applies the same rule
  from gnssb200_acq_cfg.part_index/part_count); the only collective is one all-gather of the row
  tables (16 B per row), after which every rank runs the tiny peak / second-peak / threshold logic.
"""
from __future__ import annotations

import numpy as np

from . import abi


def streams_of_rank(n_streams: int, rank: int, world: int) -> list:
    return [s for s in range(n_streams) if s % world == rank]


def row_owner(n_rows: int, world: int) -> np.ndarray:
    return np.arange(n_rows) % world


def merge_row_tables(gathered: np.ndarray) -> np.ndarray:
    """gathered: (world, n_rows) structured array of gnssb200_acq_row as returned by an all-gather
    of every rank's table (rows a rank does not own carry peak = -1).  Returns the merged (n_rows,) table."""
    world, n_rows = gathered.shape
    merged = gathered[row_owner(n_rows, world), np.arange(n_rows)]
    if (merged["peak"] < 0).any():
        raise ValueError("row table incomplete after merge: some owner did not write its rows")
    return merged


def prns_of_rank(prn_list, rank: int, world: int) -> list:
    """Serial-search cell map: the sweep of one PRN over its Doppler bins is one sequential run
    of one correlator channel (the NCO phases carry over from bin to bin, quirk Q6), so the PRN is the unit that
    shards: entry i of the list is searched by rank i % world."""
    return [p for i, p in enumerate(prn_list) if i % world == rank]


def merge_cell_maps(gathered_cells: np.ndarray, gathered_counts: np.ndarray, prn_list, world: int) -> dict:
    """gathered_cells: (world, per_rank, cells_cap) gnssb200_serial_cell, gathered_counts: (world, per_rank); rank r's slot
    j holds entry r + j*world of prn_list.  Returns {prn: cells} in the order of prn_list."""
    out = {}
    for i, p in enumerate(prn_list):
        r, j = i % world, i // world
        out[int(p)] = gathered_cells[r, j, : gathered_counts[r, j]].copy()
    return out


def all_gather_rows(rows_tensor, world: int):
    """torch.distributed all-gather of a device (NCCL) or host (gloo) uint8 row table."""
    import torch
    import torch.distributed as dist

    out = torch.empty(world * rows_tensor.numel(), dtype=rows_tensor.dtype, device=rows_tensor.device)
    dist.all_gather_into_tensor(out, rows_tensor)
    return out.cpu().numpy().view(abi.ACQ_ROW_DTYPE).reshape(world, -1)
