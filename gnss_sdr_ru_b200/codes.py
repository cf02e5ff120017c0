"""Spreading-code generators used on the host side (record synthesis, code spectra upload).

GPS L1 C/A  : G1/G2 Gold codes.  Same ±1 chips as the reference C correlator
              (OSG/correlator/correlator.c:63-91, chip 0 forced to +1) and the Scilab generator
              (SCI/GPS/L1/include/generateCAcode.sci:42-87); PRN 1 starts 1100100000.
GLONASS L1OF: 511-chip m-sequence (ST code), 9-stage register, output tap 7, feedback 5 xor 9,
              all-ones start (SCI/GLONASS/L1/include/generateSTcode.sci:35-42,
              NAM/rtl/code_gen.v:121-139).

These are plain restatements of the published ICD generators; chips are returned as int8 ±1 where
logic '1' maps to +1 for C/A (reference convention 2*bit-1).
"""
from __future__ import annotations

import functools

import numpy as np

# G2 output phase-select taps per PRN (IS-GPS-200 table 3-I), 1-based register stages.
_G2_TAPS = [
    (2, 6), (3, 7), (4, 8), (5, 9), (1, 9), (2, 10), (1, 8), (2, 9), (3, 10), (2, 3), (3, 4), (5, 6),
    (6, 7), (7, 8), (8, 9), (9, 10), (1, 4), (2, 5), (3, 6), (4, 7), (5, 8), (6, 9), (1, 3), (4, 6),
    (5, 7), (6, 8), (7, 9), (8, 10), (1, 6), (2, 7), (3, 8), (4, 9),
]


@functools.lru_cache(maxsize=None)
def _ca_bits(prn: int) -> bytes:
    g1 = [1] * 10
    g2 = [1] * 10
    t1, t2 = _G2_TAPS[prn - 1]
    out = bytearray(1023)
    for i in range(1023):
        out[i] = g1[9] ^ g2[t1 - 1] ^ g2[t2 - 1]
        f1 = g1[2] ^ g1[9]
        f2 = g2[1] ^ g2[2] ^ g2[5] ^ g2[7] ^ g2[8] ^ g2[9]
        g1 = [f1] + g1[:9]
        g2 = [f2] + g2[:9]
    return bytes(out)


def ca_code(prn: int) -> np.ndarray:
    """C/A chips for PRN 1..32 as int8 in {-1,+1}; logic 1 -> +1 (reference: 2*prn_code-1)."""
    if not 1 <= prn <= 32:
        raise ValueError("GPS PRN must be 1..32")
    bits = np.frombuffer(_ca_bits(prn), dtype=np.uint8).astype(np.int8)
    return (2 * bits - 1).astype(np.int8)


@functools.lru_cache(maxsize=None)
def _st_bits() -> bytes:
    reg = [1] * 9
    out = bytearray(511)
    for i in range(511):
        out[i] = reg[6]
        fb = reg[4] ^ reg[8]
        reg = [fb] + reg[:8]
    return bytes(out)


def st_code() -> np.ndarray:
    """GLONASS ST code, 511 chips, int8 ±1 with the Scilab generator's polarity.

    generateSTcode.sci works in ±1 arithmetic: register filled with -1, output reg(7), feedback
    reg(5)*reg(9), and the output is negated at the end.  Mapping -1 <-> logic 1 turns the products
    into XORs; the final negation makes logic 1 -> +1.
    """
    bits = np.frombuffer(_st_bits(), dtype=np.uint8).astype(np.int8)
    return (2 * bits - 1).astype(np.int8)
