"""Two ranks on two GPUs (NCCL): the sharded acquisition (rows r % world, one all-gather of the row tables) and the
sharded serial-search cell map (PRNs i % world) give the single-GPU tables.  Skipped on a one-GPU box; the host
logic (merge rules, ownership) is covered on CPU by tests/test_sharding_gloo.py."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    try:
        _worker_body(rank, world, port, q)
    except Exception:  # report instead of leaving the other rank in a collective and the parent waiting
        import traceback

        q.put((rank, False, traceback.format_exc()))


def _worker_body(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from gnss_sdr_ru_b200 import abi
    from gnss_sdr_ru_b200.acquisition import AcquisitionEngine, Settings
    from gnss_sdr_ru_b200.receiver import TrackingEngine, acq_serial, acq_serial_distributed
    from gnss_sdr_ru_b200.scenarios import gps_acq_scenario
    from gnss_sdr_ru_b200.synth import make_record

    ok = True
    # ---- FFT acquisition: 5 PRNs x 29 bins, every rank holds the record
    rec = make_record(gps_acq_scenario(11, prns=(3, 9)), 16000 * 3, seed=11)
    st = Settings.gps(acqSearchBand=14.0, acqCohIntegration=1, acqSatelliteList=[3, 5, 9, 22, 31])
    d = torch.from_numpy(rec.view(np.uint8).copy()).cuda()
    ae = AcquisitionEngine(device=rank)
    res = ae.acquisition_distributed(d.data_ptr(), rec.size // 2, st, return_rows=True)
    one = ae.acquisition(rec, st, return_rows=True)  # the whole grid on this GPU
    for k in ("bin", "codePhaseRaw", "codePhase", "freqChannel", "carrFreq", "peakMetric", "peak", "second"):
        ok = ok and np.array_equal(res[k], one[k])
    ok = ok and np.array_equal(res["rows"], one["rows"])
    # ---- serial search cell map: 5 PRNs, two bins either side
    n = 8192 * 700
    trec = make_record(gps_acq_scenario(12, prns=(27,)), n, seed=12)
    td = torch.from_numpy(trec.view(np.uint8).copy()).cuda()
    eng = TrackingEngine(n_streams=1, device=rank)
    prns = [27, 9, 32, 1, 5]
    got = acq_serial_distributed(eng.h, td.data_ptr(), abi.FMT_INT8_IQ, n, prns, search_max_f=2, max_prn_delay=40, cells_cap=400)
    want = acq_serial(eng.h, td.data_ptr(), abi.FMT_INT8_IQ, n, prns, search_max_f=2, max_prn_delay=40, cells_cap=400)
    ok = ok and list(got) == prns and all(np.array_equal(got[p], want[p]) and len(got[p]) > 300 for p in prns)
    q.put((rank, bool(ok), ""))
    dist.barrier()
    dist.destroy_process_group()


def _run(world):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=240) for _ in range(world)]
    finally:
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.kill()
    assert sorted(r[:2] for r in res) == [(r, True) for r in range(world)], [r[2] for r in res]


def test_distributed_api_single_rank():
    """The same calls in a one-rank NCCL group (any GPU box): partition 0 of 1, all-gather of one table, merge."""
    _run(1)


def test_two_gpus_sharded_acquisition_and_serial_search():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run(2)
