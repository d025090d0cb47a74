"""bgzip + tabix for the merged VCF without the external tools: the reference's merge stage ends with
`bgzip -f out.vcf; tabix -f -p vcf out.vcf.gz` (src/sort_vcf.py:70-76).  The writer is native
(c3r_vcf_write_bgzf, csrc/bam_io.cpp: same BGZF block writer and binning code as the BAM/BAI writer); the small
reader below (BGZF virtual offsets + .tbi) is what the tests use to check region queries against it."""
from __future__ import annotations

import gzip
import struct
import zlib

from . import lib as L


def write_vcf_gz(path: str, header: str, rows, level: int = 6) -> None:
    """`path` (BGZF) + `path`.tbi from a header string and data lines sorted by position within each contig"""
    text = header.rstrip("\n") + "\n" + "".join(r + "\n" for r in rows)
    raw = text.encode("ascii")
    rc = L.load().c3r_vcf_write_bgzf(path.encode(), raw, len(raw), level)
    if rc != 0:
        raise ValueError("c3r_vcf_write_bgzf failed (rc=%d): rows must be sorted by position within contig blocks" % rc)


# ----------------------------------------------------------------------------- reader (tests, spot checks)
def _reg2bins(beg: int, end: int):
    end -= 1
    bins = [0]
    for shift, off in ((26, 1), (23, 9), (20, 73), (17, 585), (14, 4681)):
        bins.extend(range(off + (beg >> shift), off + (end >> shift) + 1))
    return bins


def read_tbi(path: str) -> dict:
    with gzip.open(path, "rb") as fp:
        b = fp.read()
    if b[:4] != b"TBI\x01":
        raise ValueError("not a tabix index")
    n_ref, fmt, col_seq, col_beg, col_end, meta, skip, l_nm = struct.unpack_from("<8i", b, 4)
    o = 36
    names = b[o:o + l_nm].split(b"\0")[:-1]
    o += l_nm
    refs = {}
    for name in names:
        (n_bin,) = struct.unpack_from("<i", b, o)
        o += 4
        bins = {}
        for _ in range(n_bin):
            bin_id, n_chunk = struct.unpack_from("<Ii", b, o)
            o += 8
            bins[bin_id] = [struct.unpack_from("<QQ", b, o + 16 * k) for k in range(n_chunk)]
            o += 16 * n_chunk
        (n_intv,) = struct.unpack_from("<i", b, o)
        o += 4
        linear = list(struct.unpack_from("<%dQ" % n_intv, b, o))
        o += 8 * n_intv
        refs[name.decode()] = dict(bins=bins, linear=linear)
    return dict(format=fmt, col_seq=col_seq, col_beg=col_beg, col_end=col_end, meta=chr(meta), skip=skip, refs=refs)


def _read_from(fp, voff_beg: int, voff_end: int) -> bytes:
    """uncompressed bytes between two BGZF virtual offsets"""
    out = bytearray()
    coff, uoff = voff_beg >> 16, voff_beg & 0xffff
    while True:
        fp.seek(coff)
        hdr = fp.read(18)
        if len(hdr) < 18:
            break
        bsize = struct.unpack_from("<H", hdr, 16)[0] + 1
        data = zlib.decompress(fp.read(bsize - 18)[:-8], -15)
        stop = (voff_end & 0xffff) if coff == (voff_end >> 16) else len(data)
        out += data[uoff:stop]
        if coff >= (voff_end >> 16):
            break
        coff += bsize
        uoff = 0
    return bytes(out)


def tabix_query(path: str, contig: str, start1: int, end1: int) -> list:
    """data lines of `path` (BGZF + .tbi) whose [POS, POS + len(REF)) overlaps the 1-based inclusive region"""
    idx = read_tbi(path + ".tbi")
    ref = idx["refs"].get(contig)
    if ref is None:
        return []
    beg, end = start1 - 1, end1
    min_off = ref["linear"][beg >> 14] if (beg >> 14) < len(ref["linear"]) else (ref["linear"][-1] if ref["linear"] else 0)
    chunks = sorted(c for b in _reg2bins(beg, end) for c in ref["bins"].get(b, []) if c[1] > min_off)
    out = []
    with open(path, "rb") as fp:
        for cb, ce in chunks:
            for line in _read_from(fp, max(cb, min_off) if cb < min_off else cb, ce).decode("ascii").splitlines():
                c = line.split("\t", 4)
                if len(c) < 5 or c[0] != contig:
                    continue
                p0 = int(c[1]) - 1
                if p0 < end and p0 + max(1, len(c[3])) > beg:
                    out.append(line)
    return out
