"""Streaming ingest: a producer thread feeds the circular buffer in USB-read-sized pieces
while the consumer pumps blocks into the GPU channel loop; the result must be identical to the oracle's
single pass over the complete record (bit exact dump records and receiver state)."""
import threading
import time

import numpy as np
import pytest

from gnss_sdr_ru_b200 import abi

pytestmark = pytest.mark.gpu

NS = 8192
PRNS = [27, 0, 0, 31, 0, 0, 0, 0, 9, 0, 32, 5]
WARM = ((0, 1), (8, -1), (10, 1))


def _oracle(oracle_lib):
    o = oracle_lib.Oracle()
    o.cold_allocate(PRNS)
    for ch, nf in WARM:
        k = o.rx.chan[ch]
        k.n_freq, k.del_freq, k.codes = nf, (-2 * nf if nf > 0 else 1 - 2 * nf), 0
        k.carrier_freq = o.cfg.gps_carrier_ref + o.cfg.d_freq * nf
        o.ch_carrier(ch, k.carrier_freq)
    return o


@pytest.mark.parametrize("fmt", [abi.FMT_INT8_IQ, abi.FMT_PACKED2])
def test_ring_fed_tracking_equals_single_pass(oracle_lib, track_record, fmt):
    from gnss_sdr_ru_b200.ingest import StreamingIngest
    from gnss_sdr_ru_b200.receiver import TrackingEngine
    from gnss_sdr_ru_b200.synth import pack2

    rec, _ = track_record
    nblk = 900
    rec = rec[: 2 * NS * nblk]
    wire = pack2(rec) if fmt == abi.FMT_PACKED2 else rec.view(np.uint8)
    eng = TrackingEngine(n_streams=2)
    for s in range(2):
        eng.simple_cold_allocate(s, PRNS)
        for ch, nf in WARM:
            eng.warm_start(s, ch, nf)
    eng.upload()
    ing = StreamingIngest(eng, stream=1, fmt=fmt, nsamp=NS, ring_blocks=96, dump_cap=1200)
    overflowed = []

    def producer():
        piece = 16384 * 3 + 4096  # not a multiple of the block size: blocks complete across writes, the ring wraps mid-write
        pos = 0
        while pos < wire.size:
            chunk = wire[pos : pos + piece]
            if ing.write(chunk) == 0:  # ring full: the reference's collector would stop here; a file reader just waits
                overflowed.append(pos)
                time.sleep(0.0005)
                continue
            pos += chunk.size
        ing.SetFinishedLoadingData()

    t = threading.Thread(target=producer)
    t.start()
    done = 0
    t0 = time.time()
    while done < nblk and time.time() - t0 < 120:
        n = ing.pump()
        done += n
        if n == 0:
            time.sleep(0.0002)
    t.join()
    dumps, cnt = ing.sync()
    st = ing.status()
    assert done == nblk and st.blocks_done == nblk and st.bytes_in_buffer == 0 and st.finished == 1
    assert st.bytes_loaded == wire.size == st.bytes_output
    eng.download()
    ing.close()
    o = _oracle(oracle_lib)
    n, odumps, ocnt = o.run(rec, NS, nblk, dump_cap=1200)
    assert np.array_equal(cnt, ocnt)
    for ch in range(12):
        assert np.array_equal(dumps[ch, : cnt[ch]], odumps[ch, : ocnt[ch]]), f"channel {ch}"
    assert bytes(memoryview(eng.rx[1]).cast("B")) == bytes(memoryview(o.rx).cast("B"))
    assert eng.rx[0].blocks_done == 0  # the other stream of the handle was not touched
    assert overflowed, "the small ring was expected to fill up at least once"


def test_overflow_flag_and_refusal():
    from gnss_sdr_ru_b200.ingest import StreamingIngest
    from gnss_sdr_ru_b200.receiver import TrackingEngine

    eng = TrackingEngine(n_streams=1)
    eng.upload()
    ing = StreamingIngest(eng, stream=0, fmt=abi.FMT_PACKED2, nsamp=NS, ring_blocks=4)
    blk = np.zeros(NS // 2, dtype=np.uint8)
    for _ in range(4):
        assert ing.write(blk) == blk.size
    assert not ing.CheckCircularBufferOverflow()
    assert ing.write(blk[:16]) == 0 and ing.CheckCircularBufferOverflow()
    assert ing.DataLeftInBuffer() == 4 * blk.size
    ing.close()
