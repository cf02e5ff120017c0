"""Host-side mirror of the front-end sampler's circular buffer (FE/PC_SIDE_SOFTWARE/WIN/GPS1A_SAMPLER/src/
CircularBuffer.h:9-193), with the GPU channel loop as the consumer instead of the file writer
(csrc/ingest.cu).  Method names follow the reference class where a counterpart exists."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from .lib import GnssB200Error, check, lib


class StreamingIngest:
    def __init__(self, engine, stream: int, fmt: int, nsamp: int = 8192, ring_blocks: int = 4096, dump_cap: int = 0):
        self.L = lib()
        self.dump_cap = dump_cap
        self.g = self.L.gnssb200_ingest_open(engine.h, stream, fmt, nsamp, ring_blocks, dump_cap)
        if not self.g:
            msg = self.L.gnssb200_last_error_string()
            raise GnssB200Error(f"gnssb200_ingest_open failed: {msg.decode() if msg else '?'}")

    def close(self):
        if self.g:
            self.L.gnssb200_ingest_close(self.g)
            self.g = None

    # ---- producer (CollectFromUSB side) ----
    def write(self, data) -> int:
        a = np.ascontiguousarray(data).view(np.uint8).ravel()
        return int(self.L.gnssb200_ingest_write(self.g, a.ctypes.data, a.size))

    def SetFinishedLoadingData(self):
        self.L.gnssb200_ingest_finish(self.g)

    # ---- consumer (WriteBufferToFile side, here: the tracking kernel) ----
    def pump(self, max_blocks: int = 0) -> int:
        n = int(self.L.gnssb200_ingest_pump(self.g, max_blocks))
        if n < 0:
            check(-1, "gnssb200_ingest_pump")
        return n

    def status(self) -> abi.IngestStat:
        st = abi.IngestStat()
        check(self.L.gnssb200_ingest_status(self.g, C.byref(st)), "gnssb200_ingest_status")
        return st

    def DataLeftInBuffer(self) -> int:
        return int(self.status().bytes_in_buffer)

    def FinishedLoadingData(self) -> bool:
        return bool(self.status().finished)

    def CheckCircularBufferOverflow(self) -> bool:
        return bool(self.status().overflow)

    def sync(self):
        """wait for the issued kernels; returns (dump records [12][dump_cap], counts [12])"""
        if self.dump_cap > 0:
            d = np.zeros((12, self.dump_cap), dtype=abi.DUMP_DTYPE)
            c = np.zeros(12, dtype=np.int32)
            check(self.L.gnssb200_ingest_sync(self.g, d.ctypes.data, c.ctypes.data), "gnssb200_ingest_sync")
            return d, c
        check(self.L.gnssb200_ingest_sync(self.g, None, None), "gnssb200_ingest_sync")
        return None, None
