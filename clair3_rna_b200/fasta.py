"""Indexed FASTA access (replaces `samtools faidx`, /root/reference/shared/utils.py:168-194)."""
import numpy as np


def read_fai(path: str) -> dict:
    fai = {}
    with open(path + ".fai") as fp:
        for row in fp:
            c = row.rstrip("\n").split("\t")
            fai[c[0]] = tuple(int(x) for x in c[1:5])       # length, offset, linebases, linewidth
    return fai


def fetch_into(path: str, fai: dict, contig: str, start1: int, end1: int, out: np.ndarray) -> np.ndarray:
    """upper-cased bases of the 1-based inclusive region, clipped to the contig, written into the uint8 buffer `out`
    (e.g. pinned host memory, engine.PinnedPool) by the native reader (csrc/fasta_io.cpp); returns the filled part."""
    import ctypes as C
    from . import lib as L
    length, offset, linebases, linewidth = fai[contig]
    n = C.c_int64(0)
    rc = L.load().c3r_fasta_fetch(path.encode(), length, offset, linebases, linewidth, int(start1), int(end1),
                                  out.ctypes.data, int(out.size), C.byref(n))
    if rc != 0:
        raise RuntimeError("c3r_fasta_fetch(%s, %s:%d-%d) failed with %d" % (path, contig, start1, end1, rc))
    return out[:n.value]


def fetch(path: str, fai: dict, contig: str, start1: int, end1: int) -> np.ndarray:
    """upper-cased bases of the 1-based inclusive region, clipped to the contig (numpy form, no library needed)."""
    length, offset, linebases, linewidth = fai[contig]
    start1, end1 = max(1, start1), min(length, end1)
    if end1 < start1:
        return np.zeros(0, np.uint8)
    first_line, last_line = (start1 - 1) // linebases, (end1 - 1) // linebases
    with open(path, "rb") as fp:
        fp.seek(offset + first_line * linewidth)
        raw = fp.read((last_line - first_line + 1) * linewidth)
    a = np.frombuffer(raw, np.uint8)
    n_lines = last_line - first_line + 1
    if a.size == n_lines * linewidth:
        # every line complete: drop the line terminators with one strided copy
        a = a.reshape(n_lines, linewidth)[:, :linebases].reshape(-1)
    else:
        # the last line of the contig may be short (and may lack a terminator)
        full = (a.size // linewidth) * linewidth
        tail = a[full:]
        a = np.concatenate([a[:full].reshape(-1, linewidth)[:, :linebases].reshape(-1), tail[(tail != 10) & (tail != 13)]])
    lo = (start1 - 1) - first_line * linebases
    a = a[lo: lo + (end1 - start1 + 1)].copy()
    if a.size and a.max() >= 97:
        lower = (a >= 97) & (a <= 122)
        a[lower] -= 32
    return a
