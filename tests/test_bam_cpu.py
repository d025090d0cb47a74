"""BAM / BGZF / BAI reader and writer (csrc/bam_io.cpp) against an independent pure-Python
implementation of the formats (SAMv1 §4, §5.2) kept in this file: files written by the library
are parsed here with gzip + struct, files written here are read by the library."""
import gzip
import os
import struct
import zlib

import numpy as np
import pytest

from tests.golden import cases as golden_cases


@pytest.fixture(scope="module")
def bam_mod():
    from clair3_rna_b200 import build
    build.build()
    from clair3_rna_b200 import bam
    return bam


# ------------------------------------------------------------------ independent python side
def py_reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14: return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17: return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20: return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23: return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26: return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def py_parse_bam(path):
    """-> (refs [(name, len)], records [(tid, pos, flag, mapq, cigar list, seq bytes, l_seq, aux bytes)])"""
    raw = gzip.decompress(open(path, "rb").read())        # BGZF is a multi-member gzip file
    assert raw[:4] == b"BAM\1"
    l_text, = struct.unpack_from("<i", raw, 4)
    o = 8 + l_text
    n_ref, = struct.unpack_from("<i", raw, o)
    o += 4
    refs = []
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", raw, o)
        name = raw[o + 4:o + 4 + l_name - 1].decode()
        l_ref, = struct.unpack_from("<i", raw, o + 4 + l_name)
        refs.append((name, l_ref))
        o += 8 + l_name
    recs = []
    while o < len(raw):
        bs, = struct.unpack_from("<i", raw, o)
        tid, pos, l_name, mapq, _bin, n_cig, flag, l_seq = struct.unpack_from("<iiBBHHHi", raw, o + 4)
        p = o + 4 + 32 + l_name
        cig = list(struct.unpack_from("<%dI" % n_cig, raw, p))
        p += 4 * n_cig
        seq = raw[p:p + (l_seq + 1) // 2]
        p += (l_seq + 1) // 2 + l_seq
        aux = raw[p:o + 4 + bs]
        recs.append((tid, pos, flag, mapq, cig, seq, l_seq, aux))
        o += 4 + bs
    return refs, recs


def py_write_bam(path, refs, recs, payload=700):
    """independent writer: small BGZF blocks (records cross block boundaries), one index chunk per record.
    recs: (tid, pos, flag, mapq, cigar list, seq bytes, l_seq, aux bytes) sorted by (tid, pos)."""
    text = b"@HD\tVN:1.6\tSO:coordinate\n"
    head = b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs))
    for name, ln in refs:
        head += struct.pack("<i", len(name) + 1) + name.encode() + b"\0" + struct.pack("<i", ln)
    stream = bytearray(head)
    spans = []                                         # uncompressed (start, end) of each record
    for i, (tid, pos, flag, mapq, cig, seq, l_seq, aux) in enumerate(recs):
        rl = sum(c >> 4 for c in cig if (c & 15) in (0, 2, 3, 7, 8))
        name = b"q%d\0" % i
        body = struct.pack("<iiBBHHHiiii", tid, pos, len(name), mapq, py_reg2bin(pos, pos + max(rl, 1)), len(cig), flag,
                           l_seq, -1, -1, 0) + name + struct.pack("<%dI" % len(cig), *cig) + seq + b"\xff" * l_seq + aux
        spans.append((len(stream), len(stream) + 4 + len(body)))
        stream += struct.pack("<i", len(body)) + body
    # cut into blocks
    blocks, coffs, out = [], [], bytearray()
    for s in range(0, len(stream), payload):
        chunk = bytes(stream[s:s + payload])
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        data = co.compress(chunk) + co.flush()
        total = len(data) + 26
        coffs.append(len(out))
        out += struct.pack("<BBBBIBBHBBHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, 66, 67, 2, total - 1) + data
        out += struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk))
    eof_off = len(out)
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    data = co.compress(b"") + co.flush()
    out += struct.pack("<BBBBIBBHBBHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, 66, 67, 2, len(data) + 25) + data + struct.pack("<II", 0, 0)
    open(path, "wb").write(out)

    def voff(u):
        b = u // payload
        if b >= len(coffs):
            return eof_off << 16
        return (coffs[b] << 16) | (u % payload)

    ix = bytearray(b"BAI\1" + struct.pack("<i", len(refs)))
    for tid in range(len(refs)):
        bins, linear = {}, {}
        for (t, pos, flag, mapq, cig, seq, l_seq, aux), (us, ue) in zip(recs, spans):
            if t != tid:
                continue
            rl = sum(c >> 4 for c in cig if (c & 15) in (0, 2, 3, 7, 8))
            end = pos + max(rl, 1)
            bins.setdefault(py_reg2bin(pos, end), []).append((voff(us), voff(ue)))
            for w in range(pos >> 14, ((end - 1) >> 14) + 1):
                linear.setdefault(w, voff(us))
        ix += struct.pack("<i", len(bins))
        for b in sorted(bins):
            ix += struct.pack("<Ii", b, len(bins[b]))
            for cb, ce in bins[b]:
                ix += struct.pack("<QQ", cb, ce)
        n_intv = max(linear) + 1 if linear else 0
        ix += struct.pack("<i", n_intv)
        last = 0
        for w in range(n_intv):
            last = linear.get(w, last)
            ix += struct.pack("<Q", last)
    open(path + ".bai", "wb").write(ix)


def batch_records(batch, tid=0):
    recs = []
    for i in range(batch.n_reads):
        cig = [int(c) for c in batch.cigar[batch.cigar_off[i]:batch.cigar_off[i + 1]]]
        ql = sum(c >> 4 for c in cig if (c & 15) in (0, 1, 4, 7, 8))
        a = int(batch.seq_off[i]) // 2
        seq = bytes(batch.seq[a:a + (ql + 1) // 2])
        aux = b"HPC" + bytes([int(batch.hp[i])]) if batch.hp[i] else b""
        recs.append((tid, int(batch.pos[i]), int(batch.flag[i]), int(batch.mapq[i]), cig, seq, ql, aux))
    return recs


def assert_same(a, b):
    for k in ("pos", "flag", "mapq", "hp", "cigar_off", "cigar", "seq_off", "seq"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k


# ------------------------------------------------------------------ tests
@pytest.mark.parametrize("name", ["cfg1_ont_drna", "cfg4_hifi_phased"])
def test_writer_output_parses_with_an_independent_reader(bam_mod, tmp_path, name):
    batch, ref_bytes, contig = golden_cases.build(name)
    path = str(tmp_path / "w.bam")
    bam_mod.write_bam(path, [(contig, len(ref_bytes))], {contig: batch})
    refs, recs = py_parse_bam(path)
    assert refs == [(contig, len(ref_bytes))]
    assert len(recs) == batch.n_reads
    want = batch_records(batch)
    for got, exp in zip(recs, want):
        assert got[:5] == exp[:5]
        assert got[6] == exp[6] and got[5] == exp[5] and got[7] == exp[7]
    # the last BGZF block is the 28-byte EOF marker
    assert open(path, "rb").read()[-28:] == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


@pytest.mark.parametrize("name", ["cfg1_ont_drna", "phased_noisy"])
def test_fetch_roundtrip_and_regions(bam_mod, tmp_path, name):
    batch, ref_bytes, contig = golden_cases.build(name)
    n = len(ref_bytes)
    path = str(tmp_path / "r.bam")
    bam_mod.write_bam(path, [("chr0", 5000), (contig, n), ("chrZ", 100)], {contig: batch}, level=1)
    with bam_mod.BamFile(path, threads=3) as bf:
        assert bf.references == ["chr0", contig, "chrZ"]
        assert bf.lengths == [5000, n, 100]
        assert_same(bf.fetch(contig, 1, n), batch)
        assert bf.fetch("chr0", 1, 5000).n_reads == 0 and bf.fetch("chrZ", 1, 100).n_reads == 0
        rng = np.random.default_rng(5)
        for _ in range(25):
            s = int(rng.integers(1, n))
            e = min(n, s + int(rng.integers(1, 40000)))
            assert_same(bf.fetch(contig, s, e), batch.fetch(s, e))
        stats = bf.idxstats()
        assert stats[1][2] + stats[1][3] == batch.n_reads and stats[0][2:] == (0, 0)
        with pytest.raises(KeyError):
            bf.fetch("nope", 1, 10)


def test_reader_on_independently_written_file(bam_mod, tmp_path):
    """file and index produced by the python writer above: 700-byte BGZF blocks, so almost every record
    crosses a block boundary and every record is its own index chunk"""
    batch, ref_bytes, contig = golden_cases.build("cfg2_ont_cdna")
    n = len(ref_bytes)
    path = str(tmp_path / "p.bam")
    py_write_bam(path, [(contig, n)], batch_records(batch))
    with bam_mod.BamFile(path) as bf:
        assert_same(bf.fetch(contig, 1, n), batch)
        rng = np.random.default_rng(9)
        for _ in range(15):
            s = int(rng.integers(1, n))
            e = min(n, s + int(rng.integers(1, 30000)))
            assert_same(bf.fetch(contig, s, e), batch.fetch(s, e))


def test_special_records(bam_mod, tmp_path):
    """HP tag in every integer width among other aux types, CG:B,I long CIGAR, SEQ '*', odd lengths"""
    M, I, D, N, S = 0, 1, 2, 3, 4
    def c(l, op): return (l << 4) | op
    seq7 = bytes([0x12, 0x48, 0x84, 0x10])                 # A C G T T G A (7 bases)
    long_ops = []
    for k in range(40000):
        long_ops += [c(1, M), c(1, D)] if k % 2 == 0 else [c(1, M), c(1, I)]
    ql = sum(x >> 4 for x in long_ops if (x & 15) in (0, 1, 4))
    rl = sum(x >> 4 for x in long_ops if (x & 15) in (0, 2, 3))
    long_seq = bytes([0x12] * ((ql + 1) // 2))
    cg_aux = b"CGBI" + struct.pack("<i", len(long_ops)) + struct.pack("<%dI" % len(long_ops), *long_ops)
    recs = [
        (0, 100, 0, 60, [c(7, M)], seq7, 7, b"NMC\x01" + b"HPc\x01" + b"XZZabc\0"),
        (0, 120, 16, 60, [c(3, M), c(50, N), c(4, M)], seq7, 7, b"XBBs\x02\x00\x00\x00\x01\x00\x02\x00" + b"HPS\x02\x00"),
        (0, 130, 0, 30, [c(7, M)], seq7, 7, b"HPi\x03\x00\x00\x00" + b"XFf\x00\x00\x80\x3f"),
        (0, 140, 0, 30, [c(7, M)], b"", 0, b"XAAq"),                                  # SEQ '*'
        (0, 150, 0, 60, [c(ql, S), c(rl, N)], long_seq, ql, cg_aux + b"HPC\x02"),      # long CIGAR in CG
        (0, 160, 4, 0, [], seq7, 7, b""),                                              # placed unmapped
    ]
    path = str(tmp_path / "s.bam")
    py_write_bam(path, [("c", 200000)], recs, payload=5000)
    with bam_mod.BamFile(path) as bf:
        b = bf.fetch("c", 1, 200000)
    assert b.n_reads == 6
    assert b.hp.tolist() == [1, 2, 3, 0, 2, 0]
    assert b.read_cigar(1) == [(3, M), (50, N), (4, M)]
    assert b.read_seq(0) == "ACGTTGA="                      # odd length: padded with a zero nibble
    assert b.read_seq(3) == "NNNNNNN="
    assert b.cigar_off[5] - b.cigar_off[4] == len(long_ops)
    assert b.cigar[b.cigar_off[4]:b.cigar_off[5]].tolist() == long_ops
    assert int(b.seq_off[5] - b.seq_off[4]) == (ql + 1) // 2 * 2
    assert b.flag.tolist() == [0, 16, 0, 0, 0, 4]
    # regions: the spliced read (120..177) overlaps 170 through its N op, the others do not
    with bam_mod.BamFile(path) as bf:
        assert bf.fetch("c", 171, 171).pos.tolist() == [120, 150]
        assert bf.fetch("c", 100000, 100010).n_reads == 0


def test_long_cigar_written_by_the_library(bam_mod, tmp_path):
    from clair3_rna_b200.reads import ReadBatch
    ops = [(1, 0), (1, 2)] * 35000 + [(1, 0)]
    codes = np.array([1, 2, 4, 8] * 8751, np.uint8)[:35001]
    batch = ReadBatch.from_records("c", [(10, 0, 60, 1, ops, codes), (20, 16, 60, 0, [(5, 0)], codes[:5])])
    path = str(tmp_path / "l.bam")
    bam_mod.write_bam(path, [("c", 100000)], {"c": batch})
    refs, recs = py_parse_bam(path)
    assert len(recs[0][4]) == 2 and (recs[0][4][0] & 15) == 4 and (recs[0][4][1] & 15) == 3     # kSmN placeholder
    with bam_mod.BamFile(path) as bf:
        assert_same(bf.fetch("c", 1, 100000), batch)


def test_errors(bam_mod, tmp_path):
    p = str(tmp_path / "x.bam")
    open(p, "wb").write(b"not a bam file at all, just text that is long enough to look at")
    with pytest.raises(RuntimeError):
        bam_mod.BamFile(p)
    with pytest.raises(RuntimeError):
        bam_mod.BamFile(str(tmp_path / "missing.bam"))
    batch, ref_bytes, contig = golden_cases.build("cfg2_ont_cdna")
    good = str(tmp_path / "g.bam")
    bam_mod.write_bam(good, [(contig, len(ref_bytes))], {contig: batch})
    os.remove(good + ".bai")
    with pytest.raises(RuntimeError):
        bam_mod.BamFile(good)
    # corrupt a byte in the middle of a compressed block: CRC / inflate must catch it
    bam_mod.write_bam(good, [(contig, len(ref_bytes))], {contig: batch})
    raw = bytearray(open(good, "rb").read())
    raw[len(raw) // 2] ^= 0x5a
    open(good, "wb").write(raw)
    with pytest.raises(RuntimeError) as e:            # caught at open (header pass) or at the fetch
        with bam_mod.BamFile(good) as bf:
            bf.fetch(contig, 1, len(ref_bytes))
    assert "BGZF" in str(e.value) or "malformed" in str(e.value) or "corrupt" in str(e.value)
