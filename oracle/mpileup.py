"""ORACLE (test infrastructure, never on the product path).

CPU restatement of the text `samtools mpileup` prints for the option set the
reference uses at /root/reference/src/create_tensor_pileup.py:436-451

    samtools mpileup BAM -r ctg:s-e --reverse-del --min-MQ 5 --min-BQ 0
                     --excl-flags 2316 [--output-extra HP]

samtools/htslib are a third-party dependency of the reference (bioconda
`clair3` environment, version unpinned in the repo, README requires >= 1.10) and
their source is not under /root/reference, so this file restates the published
behaviour of htslib `sam.c: bam_plp_push / resolve_cigar2` and samtools
`bam_plcmd.c: mplp_func / pileup_seq` (>= 1.11 for ins+del at one anchor), as
listed in SURVEY.md §8(a) "A0 notes":

  * a read is admitted iff mapped, !(flag & excl_flags), MAPQ >= min_mq and not
    (PAIRED and not PROPER_PAIR); htslib additionally always drops
    flag & (UNMAP|SECONDARY|QCFAIL|DUP);
  * a column is printed for each 1-based position inside the region that is
    covered by >= 1 admitted read, including positions under D (`*` fwd, `#` rev
    with --reverse-del) and N (`>` fwd, `<` rev);
  * per read, in BAM order: `^`+chr(min(mapq,93)+33) at its first column, the
    base letter (upper fwd, lower rev; no -f so never `.`/`,`), then at the last
    column before an indel `+<len><seq>` and/or `-<len><N..>` (case by strand),
    `$` at its last column;
  * the 7th column (--output-extra HP) lists one value per read of the column
    in the same order, `*` when the tag is absent.

Parity unpinned by any reference-owned test (the reference has none); the pin
is the reference's own parser run verbatim on this text (tests/golden).
"""
from __future__ import annotations

import numpy as np

NT16 = "=ACMGRSVTWYHKDBN"
M, I, D, N, S, H, PAD, EQ, X = range(9)
_ALWAYS_DROP = 0x4 | 0x100 | 0x200 | 0x400


def admitted(flag: int, mapq: int, excl_flags: int, min_mq: int) -> bool:
    if flag & _ALWAYS_DROP or flag & excl_flags:
        return False
    if mapq < min_mq:
        return False
    if (flag & 0x1) and not (flag & 0x2):
        return False
    return True


class _Cursor:
    """Per-read column program: blocks of reference offsets with their tokens."""
    __slots__ = ("pos", "end", "rev", "hp", "blocks", "bi")

    def __init__(self, pos, rev, hp, mapq, cigar, codes):
        self.pos, self.rev, self.hp, self.bi = pos, rev, hp, 0
        letters = [NT16[c] for c in codes]
        if rev:
            letters = [c.lower() for c in letters]
        star = '#' if rev else '*'
        skip = '<' if rev else '>'
        nchar = 'n' if rev else 'N'
        blocks = []                      # (off_start, off_end, tokens | skip char)
        x = y = 0
        ops = [(l, op) for l, op in cigar]
        k, n = 0, len(ops)
        # leading query-only ops
        while k < n and ops[k][1] in (I, S, H, PAD):
            if ops[k][1] in (I, S):
                y += ops[k][0]
            k += 1
        while k < n:
            l, op = ops[k]
            if op in (M, EQ, X):
                toks = letters[y: y + l]
                y += l
            elif op == D:
                toks = [star] * l
            elif op == N:
                toks = None
            else:                        # I / S / H / P after the start: query only
                if op in (I, S):
                    y += l
                k += 1
                continue
            if toks is not None:
                # peek the next op at the last column of this op (htslib resolve_cigar2)
                suffix = ""
                j = k + 1
                if j < n and ops[j][1] == I:
                    ilen, yy = 0, y
                    while j < n and ops[j][1] in (I, PAD):
                        if ops[j][1] == I:
                            ilen += ops[j][0]
                        j += 1
                    suffix += "+%d%s" % (ilen, "".join(letters[yy: yy + ilen]))
                if j < n and ops[j][1] == D and op != D:
                    dlen = 0
                    while j < n and ops[j][1] == D:
                        dlen += ops[j][0]
                        j += 1
                    suffix += "-%d%s" % (dlen, nchar * dlen)
                if suffix:
                    toks = list(toks)
                    toks[-1] = toks[-1] + suffix
                blocks.append((x, x + l, toks))
            else:
                blocks.append((x, x + l, skip))
            x += l
            k += 1
        self.end = pos + x
        self.blocks = blocks
        # head / tail marks
        if blocks:
            head = '^' + chr(min(mapq, 93) + 33)
            s, e, t = blocks[0]
            if isinstance(t, list):
                t = list(t)
                t[0] = head + t[0]
                blocks[0] = (s, e, t)
            else:
                blocks[0] = (s, e, ("HEAD", head, t))
            s, e, t = blocks[-1]
            if isinstance(t, list):
                t = list(t)
                t[-1] = t[-1] + '$'
                blocks[-1] = (s, e, t)

    def token(self, p):
        off = p - self.pos
        while off >= self.blocks[self.bi][1]:
            self.bi += 1
        s, e, t = self.blocks[self.bi]
        if isinstance(t, list):
            return t[off - s]
        if isinstance(t, tuple):          # skip block carrying the head mark
            return (t[1] + t[2]) if off == 0 else t[2]
        last = self.bi == len(self.blocks) - 1 and off == e - 1
        return t + ('$' if last else '')


def mpileup_rows(batch, start1: int, end1: int, excl_flags: int = 2316, min_mq: int = 5, bed=None):
    """yield (pos1, depth, bases, hp_csv) for every printed column of ctg:start1-end1.

    `batch` is a clair3_rna_b200.reads.ReadBatch in coordinate order.  `bed` = [(start0, end0)] restates
    `samtools mpileup -l file.bed`: only columns whose 0-based position lies in an interval are printed
    (bam_plcmd.c: `if (conf->bed && !bed_overlap(conf->bed, name, pos, pos + 1)) continue`)."""
    import bisect
    if bed is not None:
        bed = sorted((int(a), int(b)) for a, b in bed if b > a)
        merged = []
        for a, b in bed:
            if merged and a <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], b)
            else:
                merged.append([a, b])
        bed_starts = [a for a, _ in merged]
        bed_ends = [b for _, b in merged]

    def in_bed(q):
        if bed is None:
            return True
        k = bisect.bisect_right(bed_starts, q) - 1
        return k >= 0 and q < bed_ends[k]
    n = batch.n_reads
    pos = batch.pos
    order_ok = np.all(pos[1:] >= pos[:-1]) if n > 1 else True
    assert order_ok, "records must be coordinate sorted"
    s0, e0 = start1 - 1, end1            # 0-based half open
    active = []
    nxt = 0
    p = s0
    while p < e0:
        # admit reads starting at or before p
        while nxt < n and pos[nxt] <= p:
            i = nxt
            nxt += 1
            if not admitted(int(batch.flag[i]), int(batch.mapq[i]), excl_flags, min_mq):
                continue
            cig = batch.read_cigar(i)
            a, b = int(batch.seq_off[i]), int(batch.seq_off[i + 1])
            by = batch.seq[a // 2: b // 2]
            codes = np.empty(by.size * 2, np.uint8)
            codes[0::2] = by >> 4
            codes[1::2] = by & 15
            cur = _Cursor(int(pos[i]), bool(batch.flag[i] & 0x10), int(batch.hp[i]), int(batch.mapq[i]),
                          cig, codes.tolist())
            if cur.end > p and cur.blocks:
                active.append(cur)
        if not active:
            if nxt >= n:
                break
            p = max(p + 1, int(pos[nxt]))           # jump over the uncovered gap
            continue
        active = [c for c in active if c.end > p]
        if active and in_bed(p):
            toks = [c.token(p) for c in active]
            hps = ",".join(str(c.hp) if c.hp else '*' for c in active)
            yield p + 1, len(active), "".join(toks), hps
        p += 1


def mpileup_text(batch, contig: str, start1: int, end1: int, excl_flags=2316, min_mq=5, with_hp=False, bed=None):
    """iterator over the text lines samtools would print."""
    for pos1, depth, bases, hps in mpileup_rows(batch, start1, end1, excl_flags, min_mq, bed=bed):
        line = "%s\t%d\tN\t%d\t%s\t%s" % (contig, pos1, depth, bases, "I" * depth)
        if with_hp:
            line += "\t" + hps
        yield line
