"""TEST INFRASTRUCTURE (oracle) -- NumPy restatement of the Scilab receivers' navigation-message search.

Follows, line by line:
  SCI/GPS/L1/findPreambles.sci:30-169        [firstSubFrame, activeChnList] = findPreambles(status, I_P, nCh)
  SCI/GPS/L1/include/navPartyChk.sci:57-99   status = navPartyChk(ndat)
  SCI/GLONASS/L1/findTimeMarks.sci:25-66     [firstString, activeChnList] = findTimeMarks(status, I_P, nCh)
(SCI = trunk/GNSS_SOFTWARE_RECEIVERS/POSTPROCESSING_SCILAB_RECEIVERS).  Indices are the reference's 1-based
values.  Scilab's convol() is written out as numpy.convolve on integers (Scilab computes it through an FFT;
its result equals these integers up to rounding noise).  Parity unpinned by any reference test vector; the
parity routine is pinned against IS-GPS-200 table 20-XIV by encoding words and checking them
(tests/test_navbits.py).  Only tests/, smoke() and bench.py's CPU leg may import this module.
"""
from __future__ import annotations

import numpy as np


def navPartyChk(ndat):
    """navPartyChk.sci:57-99.  ndat: 32 values in {-1, 0, +1}; returns +1 / -1 (parity ok) or 0."""
    d = np.concatenate([[0], np.asarray(ndat, dtype=np.int64)])  # 1-based
    if d[2] != 1:
        d[3:27] = -d[3:27]
    P = lambda *idx: int(np.prod(d[list(idx)]))
    parity = [
        P(1, 3, 4, 5, 7, 8, 12, 13, 14, 15, 16, 19, 20, 22, 25),
        P(2, 4, 5, 6, 8, 9, 13, 14, 15, 16, 17, 20, 21, 23, 26),
        P(1, 3, 5, 6, 7, 9, 10, 14, 15, 16, 17, 18, 21, 22, 24),
        P(2, 4, 6, 7, 8, 10, 11, 15, 16, 17, 18, 19, 22, 23, 25),
        P(2, 3, 5, 7, 8, 9, 11, 12, 16, 17, 18, 19, 20, 23, 24, 26),
        P(1, 5, 7, 8, 10, 11, 12, 13, 15, 17, 21, 24, 25, 26),
    ]
    if sum(int(p == x) for p, x in zip(parity, d[27:33])) == 6:
        return int(-1 * d[2])
    return 0


def findPreambles(trkRslt_status, trkRslt_I_P, set_numberOfChannels=None):
    """findPreambles.sci:30-169.  trkRslt_status: sequence of chars ('-' = not tracking);
    trkRslt_I_P: array [channel][ms].  Returns (firstSubFrame[nCh] 1-based ms or 0, activeChnList 1-based)."""
    I_P = np.asarray(trkRslt_I_P)
    n_ch = I_P.shape[0] if set_numberOfChannels is None else set_numberOfChannels
    searchStartOffset = 5000
    firstSubFrame = np.zeros(n_ch, dtype=np.int64)
    preamble_bits = np.array([1, 1, -1, 1, -1, -1, -1, 1])
    preamble_ms = np.kron(preamble_bits, np.ones(20, dtype=np.int64))
    activeChnList = [k + 1 for k in range(len(trkRslt_status)) if trkRslt_status[k] != "-"]
    kept = list(activeChnList)
    for channelNr in activeChnList:
        row = I_P[channelNr - 1]
        bits = np.sign(row[searchStartOffset:]).astype(np.int64)
        if bits.size == 0:
            kept.remove(channelNr)
            continue
        tlm = np.convolve(preamble_ms, bits)
        tlm = tlm[159:]  # tlmXcorrResult(160:length(...))
        index = np.nonzero(np.abs(tlm) > 153)[0] + 1 + searchStartOffset
        for i in range(len(index)):
            index2 = index - index[i]
            if np.any(index2 == 6000):
                lo, hi = index[i] - 40, index[i] + 20 * 60 - 1  # 1-based inclusive
                b = np.asarray(row[lo - 1 : hi], dtype=np.float64)
                b = b.reshape(-1, 20).sum(axis=1)  # matrix(bits, 20, n); sum(bits, 'r')
                b = np.sign(b).astype(np.int64)
                if navPartyChk(b[0:32]) != 0 and navPartyChk(b[30:62]) != 0:
                    firstSubFrame[channelNr - 1] = index[i]
                    break
        if firstSubFrame[channelNr - 1] == 0:
            kept.remove(channelNr)  # setdiff(activeChnList, channelNr)
    return firstSubFrame, kept


def findTimeMarks(trkRslt_status, trkRslt_I_P, set_numberOfChnls=None):
    """findTimeMarks.sci:25-66.  Returns (firstString[nCh] 1-based ms or 0, list of channels with a mark).
    The reference removes a channel without a mark with ``activeChnList(channelNr) = []`` -- deletion by
    POSITION, which drops the wrong entry (or raises) once an earlier channel was inactive; the list
    returned here is the evident intent (channels that have a time mark)."""
    I_P = np.asarray(trkRslt_I_P)
    n_ch = I_P.shape[0] if set_numberOfChnls is None else set_numberOfChnls
    searchStartOffset = 0
    firstString = np.zeros(n_ch, dtype=np.int64)
    activeChnList = [k + 1 for k in range(len(trkRslt_status)) if trkRslt_status[k] != "-"]
    tm_bits = np.array([-1, 1, 1, -1, 1, -1, -1, 1, -1, -1, -1, -1, 1, -1, 1, -1, 1, 1, 1, -1, 1, 1, -1, -1, -1, 1, 1, 1, 1, 1])
    tm_long = np.kron(-tm_bits, np.ones(10, dtype=np.int64))
    kept = []
    for channelNr in activeChnList:
        nav_bits = np.sign(I_P[channelNr - 1][searchStartOffset:]).astype(np.int64)
        if nav_bits.size == 0:
            continue
        r = np.convolve(tm_long, nav_bits)
        r = r[299:]  # tm_corr_rslt(300:length(...))
        index = np.nonzero(np.abs(r) > 290)[0] + 1
        if index.size == 0:
            continue
        firstString[channelNr - 1] = index[0]
        kept.append(channelNr)
    return firstString, kept


# ---- signal-side helpers for the tests (IS-GPS-200 20.3.5.2: the transmit-side parity equations) ----
def gps_encode_word(d24, D29s, D30s):
    """24 source bits (0/1) + the last two bits of the previous word -> 30 transmitted bits (0/1)."""
    d = [0] + [int(x) for x in d24]  # 1-based source bits
    D = [0] * 31
    for i in range(1, 25):
        D[i] = d[i] ^ D30s
    x = lambda *idx: sum(d[i] for i in idx) & 1
    D[25] = D29s ^ x(1, 2, 3, 5, 6, 10, 11, 12, 13, 14, 17, 18, 20, 23)
    D[26] = D30s ^ x(2, 3, 4, 6, 7, 11, 12, 13, 14, 15, 18, 19, 21, 24)
    D[27] = D29s ^ x(1, 3, 4, 5, 7, 8, 12, 13, 14, 15, 16, 19, 20, 22)
    D[28] = D30s ^ x(2, 4, 5, 6, 8, 9, 13, 14, 15, 16, 17, 20, 21, 23)
    D[29] = D30s ^ x(1, 3, 5, 6, 7, 9, 10, 14, 15, 16, 17, 18, 21, 22, 24)
    D[30] = D29s ^ x(3, 5, 6, 8, 9, 10, 11, 13, 15, 19, 22, 23, 24)
    return D[1:]


def gps_nav_bits(n_subframes, rng, preamble=(1, 0, 0, 0, 1, 0, 1, 1)):
    """n_subframes x 300 bits (0/1) of parity-correct words; word 1 of every subframe starts with the preamble."""
    out = []
    D29s = D30s = 0
    for _ in range(n_subframes):
        for w in range(10):
            d24 = rng.integers(0, 2, size=24)
            if w == 0:
                d24[:8] = preamble
            if w in (1, 9):
                # HOW and word 10: bits 23, 24 are solved so that D29 = D30 = 0 (20.3.5.2); try the four choices
                for t in range(4):
                    d24[22], d24[23] = t >> 1, t & 1
                    D = gps_encode_word(d24, D29s, D30s)
                    if D[28] == 0 and D[29] == 0:
                        break
            D = gps_encode_word(d24, D29s, D30s)
            out.extend(D)
            D29s, D30s = D[28], D[29]
    return np.array(out, dtype=np.int64)
