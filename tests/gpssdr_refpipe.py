"""doPrepIF + doAcqStrong / doAcqMedium / doAcqWeak (RT/objects/acquisition.cpp:182-570) composed, step by step, from a set of
PRIMITIVES -- the reference's own compiled ones (oracle/_ref/libgpssdr_ref.so, prefix gsr_) or the restatement's
(prefix gso_).  With the reference's primitives this is the strongest pin of oracle/gpssdr_oracle.c's search
loops this container can produce: only the loop structure is restated, every arithmetic step is the reference's."""
import ctypes as C

import numpy as np

NS, ROWLEN = 2048, 2048 + 201
R1 = np.zeros(16, np.int32)
R2 = np.array([0, 0, 0, 0, 0, 0, 0, 1, 0, 1, 0, 1, 1, 1, 1, 1], np.int32)
PREP_MS = {0: 1, 1: 10, 2: 310}


class Prims:
    def __init__(self, lib, prefix):
        for n in ("fft", "cmulsc", "cacc", "cmag", "max", "sine_gen", "wipeoff_gen"):
            setattr(self, n, getattr(lib, prefix + n))
        self.cmuls = getattr(lib, prefix + "cmuls", None)  # in-place variant, used for the 0 Hz offset (:197)


class RefAcquisition:
    def __init__(self, prims: Prims, fif=38400.0, n_rows=1240):
        self.P = prims
        self.rows = np.zeros((n_rows, ROWLEN, 2), np.int16)  # baseband_rows, zeroed like a new oracle object
        self.wipe = []
        for k in range(4):
            w = np.zeros((10 * NS, 2), np.int16)
            prims.sine_gen(w.ctypes.data, -fif - 250.0 * k, 2048000.0, 10 * NS)
            self.wipe.append(w)
        self.dft = np.zeros((10, 10, 4), np.int16)
        for r in range(10):
            prims.wipeoff_gen(self.dft[r].ctypes.data, float(np.float32(r) * 25.0 - 112.5), 1000.0, 10)

    def prep(self, _type, buff, max_rows=None, only_offsets=None):
        """doPrepIF; max_rows / only_offsets limit the work to the rows a later search reads"""
        ms = PREP_MS[_type]
        b = np.ascontiguousarray(buff, dtype=np.int16).reshape(-1, 2)[: ms * NS]
        for off in range(4):
            for m in range(ms):
                row = off * ms + m
                if (max_rows is not None and row >= max_rows) or (only_offsets is not None and off not in only_offsets):
                    continue
                a = np.ascontiguousarray(b[m * NS:(m + 1) * NS])
                w = np.ascontiguousarray(self.wipe[off][(m % 10) * NS:(m % 10 + 1) * NS])  # the tables repeat every 10 ms (:112-123)
                if off == 0 and self.P.cmuls is not None:
                    x = a.copy()
                    self.P.cmuls(x.ctypes.data, w.ctypes.data, NS, 14)
                else:
                    x = np.zeros_like(a)
                    self.P.cmulsc(a.ctypes.data, w.ctypes.data, x.ctypes.data, NS, 14)
                self.P.fft(x.ctypes.data, NS, R1.ctypes.data, 0, 1)
                self.rows[row, :100] = x[NS - 100:]
                self.rows[row, 100:100 + NS] = x
                self.rows[row, 100 + NS:100 + NS + 100] = x[:100]

    def strong_cells(self, code, doppmin, doppmax):
        """doAcqStrong :244-302, per (lcv, lcv2): (magt, indext)"""
        code = np.ascontiguousarray(code, dtype=np.int16)
        out = []
        for l in range(int(doppmin / 1000), int(doppmax / 1000)):
            for l2 in range(4):
                src = np.ascontiguousarray(self.rows[l2, 100 + l:100 + l + NS])
                x = np.zeros((NS, 2), np.int16)
                self.P.cmulsc(src.ctypes.data, code.ctypes.data, x.ctypes.data, NS, 10)
                self.P.fft(x.ctypes.data, NS, R2.ctypes.data, 1, 1)
                self.P.cmag(x.ctypes.data, NS)
                idx, mag = C.c_int32(), C.c_int32()
                self.P.max(x.ctypes.data, C.byref(idx), C.byref(mag), NS)
                out.append((l, l2, mag.value, idx.value))
        return out

    @staticmethod
    def pick_strong(cells):
        mag, res = 0, dict(code_phase=0, doppler=0, magnitude=0)
        for l, l2, magt, indext in cells:
            if magt > mag:
                mag = magt
                res = dict(code_phase=2048 - indext, doppler=int((l * 1000) + float(np.float32(l2) * 250)), magnitude=mag)
        return res

    def weak_cells(self, code, doppmin, doppmax, offsets=(0, 1, 2, 3), alignments=(0, 1)):
        """doAcqWeak :433-570, per (lcv, lcv2, k): (magt, indext) of the accumulated 10 x 2048 power matrix"""
        code = np.ascontiguousarray(code, dtype=np.int16)
        out = []
        ia, qa = C.c_int32(), C.c_int32()
        for l in range(int(doppmin / 1000), int(doppmax / 1000)):
            for l2 in offsets:
                for k in alignments:
                    power = np.zeros((10, NS), np.int32)
                    for i in range(15):
                        coh = np.zeros((10, NS, 2), np.int16)
                        for l3 in range(10):
                            src = np.ascontiguousarray(self.rows[l2 * 310 + l3 + i * 20 + k * 10, 100 + l:100 + l + NS])
                            self.P.cmulsc(src.ctypes.data, code.ctypes.data, coh[l3].ctypes.data, NS, 9)
                            self.P.fft(coh[l3].ctypes.data, NS, R2.ctypes.data, 1, 1)
                        doppler = float(l * 1000) + float(np.float32(l2 * 250))
                        shift = int(np.floor(float(i) * .02 * 2048000.0 * doppler / 1.57542e9))
                        data = np.ascontiguousarray(coh.transpose(1, 0, 2))
                        temp = np.zeros((NS, 10, 2), np.int16)
                        for d in range(NS):
                            p = data[d].ctypes.data
                            for r in range(10):
                                self.P.cacc(p, self.dft[r].ctypes.data, 10, C.byref(ia), C.byref(qa))
                                temp[d, r, 0] = ia.value >> 16
                                temp[d, r, 1] = qa.value >> 16
                        self.P.cmag(temp.ctypes.data, 10 * NS)  # x86_cmag(temp, 10) per delay; elementwise, so one call
                        pw = temp.view(np.int32).reshape(NS, 10)
                        cols = (np.arange(NS) + shift + NS) % NS
                        power[:, cols] += pw.T
                    idx, mag = C.c_int32(), C.c_int32()
                    power = np.ascontiguousarray(power)
                    self.P.max(power.ctypes.data, C.byref(idx), C.byref(mag), 10 * NS)
                    out.append((l, l2, k, mag.value, idx.value))
        return out

    @staticmethod
    def pick_weak(cells):
        mag, res = 0, dict(code_phase=0, doppler=0, magnitude=0)
        for l, l2, k, magt, indext in cells:
            if magt > mag:
                mag = magt
                res = dict(code_phase=indext % NS, doppler=int((l * 1000) + (l2 * 250) + (indext // NS) * 25.0), magnitude=mag)
        return res

    def medium_cells(self, code, doppmin, doppmax):
        """per (lcv, lcv2): (magt, indext) of x86_max over the 10 x 2048 power matrix"""
        code = np.ascontiguousarray(code, dtype=np.int16)
        out = []
        for l in range(int(doppmin / 1000), int(doppmax / 1000) + 1):
            for l2 in range(4):
                coh = np.zeros((10, NS, 2), np.int16)
                for l3 in range(10):
                    src = np.ascontiguousarray(self.rows[l2 * 20 + l3, 100 + l:100 + l + NS])
                    self.P.cmulsc(src.ctypes.data, code.ctypes.data, coh[l3].ctypes.data, NS, 10)
                    self.P.fft(coh[l3].ctypes.data, NS, R2.ctypes.data, 1, 1)
                power = np.zeros((10, NS, 2), np.int16)
                ia, qa = C.c_int32(), C.c_int32()
                data = np.ascontiguousarray(coh.transpose(1, 0, 2))  # [delay][10]
                for d in range(NS):
                    p = data[d].ctypes.data
                    for r in range(10):
                        self.P.cacc(p, self.dft[r].ctypes.data, 10, C.byref(ia), C.byref(qa))
                        power[r, d, 0] = ia.value >> 16  # an int32 >> 16 always fits the int16 field
                        power[r, d, 1] = qa.value >> 16
                self.P.cmag(power.ctypes.data, 10 * NS)
                idx, mag = C.c_int32(), C.c_int32()
                self.P.max(power.ctypes.data, C.byref(idx), C.byref(mag), 10 * NS)
                out.append((l, l2, mag.value, idx.value))
        return out

    @staticmethod
    def pick(cells):
        """the reference's choice over its loop order: a later cell wins only when strictly larger (:399-407)"""
        mag, res = 0, dict(code_phase=0, doppler=0, magnitude=0)
        for l, l2, magt, indext in cells:
            if magt > mag:
                mag = magt
                res = dict(code_phase=indext % NS, doppler=int((l * 1000) + (l2 * 250) + (indext // NS) * 25.0), magnitude=mag)
        return res


_LO16 = (8, 7, 6, 3, 0, -3, -6, -7, -8, -7, -6, -3, 0, 3, 6, 7)  # round(8 cos(2 pi k / 16)): integer carrier, 16 phases


def int_record(seed, ms, sats, noise=12):
    """complex int16 record at 2.048 Msps made with integer arithmetic only (reproducible bit for bit anywhere):
    uniform noise in [-noise, noise] + per satellite (sv 0-based, integer amplitude, Doppler Hz, code offset in samples)
    amp/8 * C/A chip * 16-phase carrier at 38400 Hz + Doppler from a 32-bit phase accumulator"""
    from gnss_sdr_ru_b200 import gpssdr_codes

    chips = gpssdr_codes.prn_gen().astype(np.int64)
    n = ms * NS
    rng = np.random.default_rng(seed)
    x = rng.integers(-noise, noise + 1, size=(n, 2)).astype(np.int64)
    k = np.arange(n, dtype=np.uint64)
    lo = np.array(_LO16, np.int64)
    for sv, amp, dopp, off in sats:
        inc = np.uint64(round((38400 + dopp) * 4294967296 / 2048000))
        ph = ((inc * k) & np.uint64(0xFFFFFFFF)) >> np.uint64(28)
        ci = ((np.arange(n, dtype=np.int64) + off) * 1023 // 2048) % 1023
        c = chips[ci, sv] * amp
        x[:, 0] += (c * lo[ph.astype(np.int64)]) >> 3
        x[:, 1] += (c * lo[(ph.astype(np.int64) + 12) & 15]) >> 3  # sin = cos shifted by three quarters of a turn
    return x.astype(np.int16)
