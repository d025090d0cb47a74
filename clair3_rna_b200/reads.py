"""Flat alignment records: the host-side layout the kernels consume.

One `ReadBatch` holds the alignment records of one contig (or of one region
of it) in BAM record order (coordinate sorted), as struct-of-arrays:

    pos        int32  [R]     0-based leftmost reference position
    flag       uint16 [R]     SAM flag
    mapq       uint8  [R]
    hp         uint8  [R]     HP:i tag value, 0 = tag absent (255 = any value > 254)
    cigar_off  int32  [R+1]   CSR offsets into `cigar`
    cigar      uint32 [n_ops] BAM encoding  (len << 4) | op,  op in MIDNSHP=X
    seq_off    int64  [R+1]   base offsets into the nibble pool, each start even
    seq        uint8  [..]    4-bit nt16 codes, high nibble first (BAM encoding)

These are the fields `samtools mpileup` reads for the option set the reference
uses (/root/reference/src/create_tensor_pileup.py:436-451); base qualities are
not kept because the reference passes --min-BQ 0 and ignores the quality column
(:487-490).
"""
from __future__ import annotations

from dataclasses import dataclass
import numpy as np

from . import params as P

NT16 = "=ACMGRSVTWYHKDBN"
_NT16_OF = np.full(256, 15, dtype=np.uint8)
for _i, _c in enumerate(NT16):
    _NT16_OF[ord(_c)] = _i
    _NT16_OF[ord(_c.lower())] = _i

_REF_CONSUME = np.array([1, 0, 1, 1, 0, 0, 0, 1, 1], dtype=np.int64)   # M I D N S H P = X
_QRY_CONSUME = np.array([1, 1, 0, 0, 1, 0, 0, 1, 1], dtype=np.int64)


def encode_seq(seq: str) -> np.ndarray:
    """ASCII bases -> nt16 codes (one per base, not yet packed)."""
    return _NT16_OF[np.frombuffer(seq.encode("ascii"), dtype=np.uint8)]


def pack_nibbles(codes: np.ndarray) -> np.ndarray:
    """nt16 codes (even count) -> bytes, high nibble first."""
    assert codes.size % 2 == 0
    c = codes.astype(np.uint8)
    return (c[0::2] << 4) | c[1::2]


def parse_cigar(text: str) -> list:
    """'10M2I5N3M' -> [(10, 0), (2, 1), (5, 3), (3, 0)]"""
    out, num = [], 0
    for ch in text:
        if ch.isdigit():
            num = num * 10 + ord(ch) - 48
        else:
            out.append((num, P.CIGAR_CHARS.index(ch)))
            num = 0
    return out


@dataclass
class ReadBatch:
    contig: str
    pos: np.ndarray
    flag: np.ndarray
    mapq: np.ndarray
    hp: np.ndarray
    cigar_off: np.ndarray
    cigar: np.ndarray
    seq_off: np.ndarray
    seq: np.ndarray

    # ------------------------------------------------------------------ build
    @staticmethod
    def from_records(contig: str, records) -> "ReadBatch":
        """records: iterable of (pos0, flag, mapq, hp, [(len, op), ...], seq_codes uint8[l_qseq])
        in coordinate order."""
        records = list(records)
        n = len(records)
        pos = np.empty(n, np.int32)
        flag = np.empty(n, np.uint16)
        mapq = np.empty(n, np.uint8)
        hp = np.empty(n, np.uint8)
        cigar_off = np.zeros(n + 1, np.int32)
        seq_off = np.zeros(n + 1, np.int64)
        cig_parts, seq_parts = [], []
        for i, (p, f, q, h, cig, codes) in enumerate(records):
            pos[i], flag[i], mapq[i], hp[i] = p, f, q, min(int(h), 255)
            cig_parts.append(np.array([(l << 4) | op for l, op in cig], dtype=np.uint32))
            cigar_off[i + 1] = cigar_off[i] + len(cig)
            codes = np.asarray(codes, dtype=np.uint8)
            if codes.size & 1:
                codes = np.concatenate([codes, np.zeros(1, np.uint8)])
            seq_parts.append(codes)
            seq_off[i + 1] = seq_off[i] + codes.size
        cigar = np.concatenate(cig_parts) if cig_parts else np.zeros(0, np.uint32)
        codes = np.concatenate(seq_parts) if seq_parts else np.zeros(0, np.uint8)
        return ReadBatch(contig, pos, flag, mapq, hp, cigar_off, cigar, seq_off, pack_nibbles(codes))

    # ------------------------------------------------------------- properties
    @property
    def n_reads(self) -> int:
        return int(self.pos.size)

    @property
    def n_ops(self) -> int:
        return int(self.cigar.size)

    def ref_span(self) -> np.ndarray:
        """reference length consumed by each read (int64 [R])."""
        ops = (self.cigar & 15).astype(np.int64)
        lens = (self.cigar >> 4).astype(np.int64) * _REF_CONSUME[ops]
        cs = np.concatenate([[0], np.cumsum(lens)])
        return cs[self.cigar_off[1:]] - cs[self.cigar_off[:-1]]

    def end(self) -> np.ndarray:
        """0-based exclusive end (bam_endpos)."""
        return self.pos.astype(np.int64) + self.ref_span()

    def n_aligned_bases(self) -> int:
        ops = self.cigar & 15
        m = (ops == P.CIG_M) | (ops == P.CIG_EQ) | (ops == P.CIG_X)
        return int((self.cigar[m] >> 4).sum())

    # ---------------------------------------------------------------- slicing
    def select(self, idx: np.ndarray) -> "ReadBatch":
        """sub-batch of the reads `idx` (ascending), arrays re-based."""
        idx = np.asarray(idx, dtype=np.int64)
        n = idx.size
        nops = (self.cigar_off[idx + 1] - self.cigar_off[idx]).astype(np.int64)
        cigar_off = np.zeros(n + 1, np.int32)
        np.cumsum(nops, out=cigar_off[1:])
        nb = self.seq_off[idx + 1] - self.seq_off[idx]
        seq_off = np.zeros(n + 1, np.int64)
        np.cumsum(nb, out=seq_off[1:])

        def gather(starts, lens, total):
            if n == 0:
                return np.zeros(0, np.int64)
            rep = np.repeat(starts - np.concatenate([[0], np.cumsum(lens)[:-1]]), lens)
            return rep + np.arange(total, dtype=np.int64)

        cg = gather(self.cigar_off[idx].astype(np.int64), nops, int(cigar_off[-1]))
        sb = gather(self.seq_off[idx] // 2, nb // 2, int(seq_off[-1] // 2))
        return ReadBatch(self.contig, self.pos[idx], self.flag[idx], self.mapq[idx], self.hp[idx],
                         cigar_off, self.cigar[cg], seq_off, self.seq[sb])

    def fetch(self, start1: int, end1: int) -> "ReadBatch":
        """records overlapping the 1-based inclusive region, as an index fetch of
        `samtools mpileup -r ctg:start-end` would deliver them."""
        s0, e0 = start1 - 1, end1            # half open, 0-based
        end = self.end()
        keep = np.nonzero((self.pos < e0) & (end > s0))[0]
        return self.select(keep)

    # --------------------------------------------------------------------- io
    def save(self, path: str) -> None:
        np.savez(path, contig=np.array(self.contig), pos=self.pos, flag=self.flag, mapq=self.mapq,
                 hp=self.hp, cigar_off=self.cigar_off, cigar=self.cigar, seq_off=self.seq_off, seq=self.seq)

    @staticmethod
    def load(path: str) -> "ReadBatch":
        z = np.load(path, allow_pickle=False)
        return ReadBatch(str(z["contig"]), z["pos"], z["flag"], z["mapq"], z["hp"],
                         z["cigar_off"], z["cigar"], z["seq_off"], z["seq"])

    # ------------------------------------------------------------ inspection
    def read_seq(self, i: int) -> str:
        a, b = int(self.seq_off[i]), int(self.seq_off[i + 1])
        by = self.seq[a // 2: b // 2]
        codes = np.empty(by.size * 2, np.uint8)
        codes[0::2] = by >> 4
        codes[1::2] = by & 15
        return "".join(NT16[c] for c in codes)

    def read_cigar(self, i: int) -> list:
        c = self.cigar[self.cigar_off[i]: self.cigar_off[i + 1]]
        return [(int(x >> 4), int(x & 15)) for x in c]
