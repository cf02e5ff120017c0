"""TEST INFRASTRUCTURE ONLY -- reader for the binary files Scilab 5's save() writes (the format of the reference's
SCI/GLONASS/L1/trackingResults.dat and SCI/GLONASS/L2/trackingResults.dat, postProcessing.sce:143).

Layout, as found in those files: per variable 24 one-byte Scilab character codes (name, blank padded, stored as six
int32), then the object: int32 type followed by
   1  real / complex matrix   m, n, it, then m*n*(it+1) doubles, column major
   4  boolean matrix          m, n, then m*n int32
   8  integer matrix          m, n, it (1/2/4 signed, 11/12/14 unsigned), then the data padded to 4 bytes
  10  string matrix           m, n, 0, m*n+1 int32 pointers, then one int32 character code per character
  15/16/17  list/tlist/mlist  n, n+1 int32 pointers (stack sizes, in doubles), then the n elements back to back
Everything is little endian and packed without alignment.  A struct is an mlist whose first element is the string
row ["st", "dims", field names...]; a 1 x N struct array stores every field as a list of N values."""
import struct

import numpy as np

_CODE = {i: str(i) for i in range(10)}
for _i in range(26):
    _CODE[10 + _i] = chr(ord("a") + _i)
    _CODE[-(10 + _i)] = chr(ord("A") + _i)
_CODE.update({36: "_", 37: "#", 38: "!", 39: "$", 40: " ", 41: "(", 42: ")", 43: ";", 44: ":", 45: "+", 46: "-", 47: "*",
              48: "/", 49: "\\", 50: "=", 51: ".", 52: ",", 53: "'", 54: "[", 55: "]", 56: "%", 57: "|", 58: "&", 59: "<",
              60: ">", 61: "~", 62: "^"})


def _text(codes):
    return "".join(_CODE.get(c, "?") for c in codes)


def _obj(d, off):
    (t,) = struct.unpack_from("<i", d, off)
    if t == 1:
        m, n, it = struct.unpack_from("<3i", d, off + 4)
        cnt = m * n
        a = np.frombuffer(d, "<f8", cnt * (it + 1), off + 16)
        v = a[:cnt] if not it else a[:cnt] + 1j * a[cnt:]
        return v.reshape(n, m).T.copy(), off + 16 + 8 * cnt * (it + 1)
    if t == 4:
        m, n = struct.unpack_from("<2i", d, off + 4)
        a = np.array(struct.unpack_from("<%di" % (m * n), d, off + 12), dtype=bool).reshape(n, m).T
        return a, off + 12 + 4 * m * n
    if t == 8:
        m, n, it = struct.unpack_from("<3i", d, off + 4)
        dt = {1: "i1", 2: "<i2", 4: "<i4", 11: "u1", 12: "<u2", 14: "<u4"}[it]
        a = np.frombuffer(d, dt, m * n, off + 16).reshape(n, m).T.copy()
        return a, off + 16 + ((a.nbytes + 3) & ~3)
    if t == 10:
        m, n, _ = struct.unpack_from("<3i", d, off + 4)
        cnt = m * n
        ptr = struct.unpack_from("<%di" % (cnt + 1), d, off + 16)
        base = off + 16 + 4 * (cnt + 1)
        strs = [_text(struct.unpack_from("<%di" % (ptr[k + 1] - ptr[k]), d, base + 4 * (ptr[k] - 1))) for k in range(cnt)]
        return np.array(strs, dtype=object).reshape(n, m).T, base + 4 * (ptr[cnt] - 1)
    if t in (15, 16, 17):
        (n,) = struct.unpack_from("<i", d, off + 4)
        ptr = struct.unpack_from("<%di" % (n + 1), d, off + 8)
        pos = off + 8 + 4 * (n + 1)
        items = []
        for k in range(n):
            if ptr[k + 1] == ptr[k]:
                items.append(None)
                continue
            v, pos = _obj(d, pos)
            items.append(v)
        if t != 15 and items and isinstance(items[0], np.ndarray) and items[0].dtype == object:
            names = [str(x) for x in items[0].ravel()]
            return {"__type__": names[0], **dict(zip(names[1:], items[1:]))}, pos
        return items, pos
    raise ValueError("Scilab object type %d at offset %d is not handled" % (t, off))


def load(path):
    """{variable name: value}: matrices as 2-D numpy arrays, strings as object arrays, lists as Python lists, tlists /
    mlists (structs) as dicts of their fields plus "__type__"."""
    d = open(path, "rb").read()
    out, off = {}, 0
    while off + 28 <= len(d):
        name = _text(struct.unpack_from("<24b", d, off)).strip()
        out[name], off = _obj(d, off + 24)
    return out
