"""Host-side mirror of the Scilab receivers' navigation-message search, computed on the GPU from the
device buffers the tracking kernels wrote (csrc/navbits.cu):

  [firstSubFrame, activeChnList] = findPreambles(trkRslt_status, trkRslt_I_P, set_numberOfChannels)
                                   SCI/GPS/L1/findPreambles.sci:30-169
  [firstString, activeChnList]   = findTimeMarks(trkRslt_status, trkRslt_I_P, set_numberOfChnls)
                                   SCI/GLONASS/L1/findTimeMarks.sci:25-66

Same names and argument meaning as the reference; indices are its 1-based millisecond counts."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from .lib import GnssB200Error, check, lib

NAV_F64, NAV_I32 = 0, 1


class NavBitsEngine:
    def __init__(self, handle=None, device: int = 0):
        self.L = lib()
        self._own = handle is None
        self.h = handle if handle is not None else self.L.gnssb200_open(device, None)
        if not self.h:
            raise GnssB200Error("gnssb200_open failed (no CUDA device? there is no CPU path)")

    def close(self):
        if self._own and self.h:
            self.L.gnssb200_close(self.h)
            self.h = None

    def _run(self, fn, d_ptr, dtype, ch_stride, ms_stride, n_ch, n_ms, status, stream=None):
        active = np.array([0 if s == "-" else 1 for s in status], dtype=np.int32) if status is not None else np.ones(n_ch, np.int32)
        if len(active) != n_ch:
            raise ValueError("status must have one entry per channel")
        first = np.zeros(n_ch, dtype=np.int32)
        keep = np.zeros(n_ch, dtype=np.int32)
        check(fn(self.h, d_ptr, dtype, ch_stride, ms_stride, n_ch, n_ms, active.ctypes.data, first.ctypes.data, keep.ctypes.data, stream),
              fn.__name__)
        return first.astype(np.int64), [k + 1 for k in range(n_ch) if keep[k]]

    # ---- device-resident input (what the tracking kernels left in HBM) ----
    def findPreambles_device(self, d_ptr, dtype, ch_stride_bytes, ms_stride_bytes, n_ch, n_ms, status=None, stream=None):
        return self._run(self.L.gnssb200_find_preambles, d_ptr, dtype, ch_stride_bytes, ms_stride_bytes, n_ch, n_ms, status, stream)

    def findTimeMarks_device(self, d_ptr, dtype, ch_stride_bytes, ms_stride_bytes, n_ch, n_ms, status=None, stream=None):
        return self._run(self.L.gnssb200_find_time_marks, d_ptr, dtype, ch_stride_bytes, ms_stride_bytes, n_ch, n_ms, status, stream)

    # ---- the reference's call shape: host arrays in, (first, activeChnList) out ----
    def _host(self, dev_fn, trkRslt_status, trkRslt_I_P):
        import torch

        a = np.ascontiguousarray(trkRslt_I_P)
        if a.dtype.kind in "iu":
            a = a.astype(np.int32)
            dtype = NAV_I32
        else:
            a = a.astype(np.float64)
            dtype = NAV_F64
        n_ch, n_ms = a.shape
        d = torch.from_numpy(a).cuda()
        return dev_fn(d.data_ptr(), dtype, d.stride(0) * d.element_size(), d.element_size(), n_ch, n_ms, trkRslt_status)

    def findPreambles(self, trkRslt_status, trkRslt_I_P, set_numberOfChannels=None):
        return self._host(self.findPreambles_device, trkRslt_status, trkRslt_I_P)

    def findTimeMarks(self, trkRslt_status, trkRslt_I_P, set_numberOfChnls=None):
        return self._host(self.findTimeMarks_device, trkRslt_status, trkRslt_I_P)
