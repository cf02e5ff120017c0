"""world_size-2 gloo test of the N>1 host logic: stream sharding and the acquisition row-table
all-gather + merge + finalize (the rows themselves come from the NumPy oracle here, no GPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gnss_sdr_ru_b200 import abi
    from gnss_sdr_ru_b200.partition import all_gather_rows, merge_cell_maps, merge_row_tables, prns_of_rank, row_owner, streams_of_rank

    g = np.load(os.path.join(ROOT, "tests", "golden", "acq_golden.npz"))
    rows_full = g["rows"]  # (sv, bin, [peak, arg, blk]) from the float64 oracle
    n_sv, nb, _ = rows_full.shape
    table = np.zeros(n_sv * nb, dtype=abi.ACQ_ROW_DTYPE)
    table["peak"] = -1.0
    mine = row_owner(n_sv * nb, world) == rank
    flat = rows_full.reshape(-1, 3)
    table["peak"][mine] = flat[mine, 0]
    table["code_phase"][mine] = flat[mine, 1].astype(np.int32)
    table["block"][mine] = flat[mine, 2].astype(np.int32)
    table["second"][mine] = 1.0
    t = torch.from_numpy(table.view(np.uint8).copy())
    gathered = all_gather_rows(t, world)
    merged = merge_row_tables(gathered)
    ok = np.allclose(merged["peak"], flat[:, 0].astype(np.float32)) and np.array_equal(merged["code_phase"], flat[:, 1].astype(np.int32))
    # every rank derives the same per-sv decision from the merged table
    peak_bin = [int(np.argmax(merged["peak"].reshape(n_sv, nb)[s])) + 1 for s in range(n_sv)]
    ok = ok and peak_bin == [int(b) for b in g["bin"]]
    # stream sharding covers every stream exactly once
    owned = streams_of_rank(64, rank, world)
    cnt = torch.zeros(64, dtype=torch.int32)
    cnt[owned] = 1
    dist.all_reduce(cnt)
    ok = ok and bool((cnt == 1).all())
    # serial-search cell map: PRN i of the list on rank i % world, gathered tables merged back in list order
    prns = [27, 9, 32, 1, 5]
    mine = prns_of_rank(prns, rank, world)
    per_rank, cap = (len(prns) + world - 1) // world, 7
    cells = np.zeros((per_rank, cap), dtype=abi.SERIAL_CELL_DTYPE)
    cnt_c = np.zeros(per_rank, dtype=np.int32)
    for j, p in enumerate(mine):
        cnt_c[j] = 1 + p % 5
        cells["prn"][j, : cnt_c[j]] = p
        cells["codes"][j, : cnt_c[j]] = np.arange(cnt_c[j])
    tc = torch.from_numpy(cells.view(np.uint8).reshape(-1).copy())
    gc = torch.empty(world * tc.numel(), dtype=torch.uint8)
    dist.all_gather_into_tensor(gc, tc)
    tn = torch.from_numpy(cnt_c)
    gn = torch.empty(world * per_rank, dtype=torch.int32)
    dist.all_gather_into_tensor(gn, tn)
    cm = merge_cell_maps(gc.numpy().view(abi.SERIAL_CELL_DTYPE).reshape(world, per_rank, cap), gn.numpy().reshape(world, per_rank), prns, world)
    ok = ok and list(cm) == prns and all(len(cm[p]) == 1 + p % 5 and (cm[p]["prn"] == p).all() for p in prns)
    # max-over-ranks timing reduction used by bench.py
    tmax = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ok = ok and float(tmax) == float(world)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
