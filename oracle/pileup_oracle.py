"""ORACLE (test infrastructure, never on the product path).

CPU restatement of the integer half of the reference's pileup calling path:

  column_vector()   <- generate_tensor            src/create_tensor_pileup.py:85-302
  candidate_rows()  <- CreateTensorPileup loop    src/create_tensor_pileup.py:463-611
  batch_tensor()    <- tensor_generator_from      clair3_rna/utils.py:64-138

(all paths under /root/reference).  It consumes the mpileup text restated in
oracle/mpileup.py and produces, per candidate, the 33 x C integer window, the
alt_info string and the flanking reference, i.e. exactly what the reference's
producer prints and its consumer parses.

Pinned by tests/golden/*.npz, which were produced by running the reference's
own create_tensor_pileup / tensor_generator_from verbatim in the build
container (tests/golden/make_golden.py); tests/test_oracle_golden.py checks
this file against them.
"""
from __future__ import annotations

import re
import numpy as np

CHANNEL = ('A', 'C', 'G', 'T', 'I', 'I1', 'D', 'D1', '*', 'a', 'c', 'g', 't', 'i', 'i1', 'd', 'd1', '#')
CH = {c: i for i, c in enumerate(CHANNEL)}
PH = {c: 18 + i for i, c in enumerate(('AP', 'CP', 'GP', 'TP', 'IP', 'DP', 'AM', 'CM', 'GM', 'TM', 'IM', 'DM'))}
FLANK = 16
WIN = 33
MAX_DEPTH = 144
SKIP_PROP = 0.2

_TOKEN = re.compile(r"\^.|[ACGTNacgtn#*]|[+-](\d+)|[<>$]|.")


def tokenize(bases: str):
    """bases string -> (entries, n_head, n_tail, n_skip_rev '<', n_skip_fwd '>', slots)

    entries: list of str tokens that enter the reference's base_list, in order;
    slots:   for each entry, the index of the HP value it is paired with, or -1
             for indel tokens (create_tensor_pileup.py:113-145)."""
    entries, slots = [], []
    head = tail = lt = gt = 0
    slot = 0
    i, n = 0, len(bases)
    while i < n:
        c = bases[i]
        if c == '^':
            head += 1
            i += 2
            continue
        if c in "ACGTNacgtn#*":
            entries.append(c)
            slots.append(slot)
            slot += 1
        elif c in "+-":
            j = i + 1
            num = 0
            while bases[j].isdigit():
                num = num * 10 + ord(bases[j]) - 48
                j += 1
            entries.append(c + bases[j: j + num])
            slots.append(-1)
            i = j + num
            continue
        elif c == '<':
            lt += 1
            slot += 1
        elif c == '>':
            gt += 1
            slot += 1
        elif c == '$':
            tail += 1
        i += 1
    return entries, head, tail, lt, gt, slots


def effective_ref_base(b: str) -> str:
    """evc_base_from (create_tensor_pileup.py:64-74) for an upper-cased reference base."""
    return b if b in "ACGT" else 'A'


def column_vector(pos1, bases, ref_seq, ref_start1, hp_list, snp_min_af, indel_min_af):
    """One mpileup column -> (vector, alt_items, depth, pass_af, max_skip).

    alt_items is the ordered list [(key, count)] the reference keeps in alt_dict."""
    ref_base = effective_ref_base(ref_seq[pos1 - ref_start1].upper())
    entries, head, tail, lt, gt, slots = tokenize(bases)
    max_skip = max(tail, head, lt, gt)
    phased = hp_list is not None
    vec = [0] * (30 if phased else 18)

    # distinct tokens in first-occurrence order with their multiplicities
    order, mult = [], {}
    for t in entries:
        if t in mult:
            mult[t] += 1
        else:
            mult[t] = 1
            order.append(t)

    if phased:
        hp_of = [hp_list[s] if s >= 0 else '0' for s in slots]
        for k, t in enumerate(entries):
            if t[0] in "+-":
                if k == 0:
                    continue
                prev = hp_of[k - 1]
                name = ('I' if t[0] == '+' else 'D')
                if prev == '1':
                    vec[PH[name + 'P']] += 1
                elif prev == '2':
                    vec[PH[name + 'M']] += 1
            else:
                u = t.upper()
                if u in ('A', 'C', 'G', 'T'):
                    if hp_of[k] == '1':
                        vec[PH[u + 'P']] += 1
                    elif hp_of[k] == '2':
                        vec[PH[u + 'M']] += 1

    alt, cls = {}, {}
    depth = alt_cnt = ins_cnt = del_cnt = 0
    best = {'I': 0, 'i': 0, 'D': 0, 'd': 0}
    for t in order:
        c = mult[t]
        if t[0] == '+':
            k = 'I' + ref_base + t[1:].upper()
            alt[k] = alt.get(k, 0) + c
            cls['I'] = cls.get('I', 0) + c
            ins_cnt += c
            ch = 'I' if t[1] in "ACGTN*" else 'i'
            vec[CH[ch]] += c
            best[ch] = max(best[ch], c)
        elif t[0] == '-':
            ln = len(t) - 1
            off = pos1 - ref_start1
            k = 'D' + ref_seq[off + 1: off + ln + 1]
            alt[k] = alt.get(k, 0) + c
            cls['D'] = cls.get('D', 0) + c
            del_cnt += c
            ch = 'D' if t[1] in "N*ACGT" else 'd'
            vec[CH[ch]] += c
            best[ch] = max(best[ch], c)
        else:
            u = t.upper()
            if u in ('A', 'C', 'G', 'T'):
                cls[u] = cls.get(u, 0) + c
                depth += c
                if u != ref_base:
                    alt['X' + u] = alt.get('X' + u, 0) + c
                    alt_cnt += c
                vec[CH[t]] += c
            elif t in ('*', '#'):
                del_cnt += c
                depth += c
                vec[CH[t]] += c
    ref_cnt = max(0, depth - del_cnt - ins_cnt - alt_cnt)
    if ref_cnt > 0:
        k = 'R' + ref_base
        alt[k] = alt.get(k, 0) + ref_cnt
    vec[CH['I1']], vec[CH['i1']] = best['I'], best['i']
    vec[CH['D1']], vec[CH['d1']] = best['D'], best['d']

    denom = depth if depth > 0 else 1
    ranked = sorted(cls.items(), key=lambda kv: kv[1], reverse=True)      # stable
    pass_top = bool(ranked) and ranked[0][0] != ref_base
    pass_snp = pass_indel = False
    for k, c in ranked:
        if k == ref_base:
            continue
        if k in ('I', 'D'):
            pass_indel = pass_indel or (float(c) / denom >= indel_min_af)
        else:
            pass_snp = pass_snp or (float(c) / denom >= snp_min_af)
    fwd = vec[0] + vec[1] + vec[2] + vec[3]
    rev = vec[9] + vec[10] + vec[11] + vec[12]
    vec[CH[ref_base]] = -fwd
    vec[CH[ref_base.lower()]] = -rev
    pass_af = pass_top or pass_snp or pass_indel
    return vec, list(alt.items()), depth, pass_af, max_skip


def max_del_length(pos1, bases, ref_seq, ref_start1):
    """longest deleted reference string among the column's deletion tokens (create_tensor_pileup.py:220,234-237):
    the slice of the loaded reference is what is measured, so it is clipped where the reference ends."""
    entries = tokenize(bases)[0]
    off = pos1 - ref_start1
    best = 0
    for t in entries:
        if t[0] == '-':
            best = max(best, len(ref_seq[off + 1: off + len(t)]))
    return best


def confident_tree(bed_rows, extend_start1, extend_end1):
    """bed_tree_from(bed_file_path, contig_name, bed_ctg_start, bed_ctg_end) of shared/interval_tree.py:8-77 for
    one contig: rows (start0, end0) in file order -> the intervals the tree holds."""
    out = []
    for a, b in bed_rows:
        if extend_start1 and extend_end1:
            if b < extend_start1 or a > extend_end1:
                continue
        if a == b:
            b += 1
        out.append((a, b))
    return out


def region_in(intervals, begin, end):
    """is_region_in(tree, ctg, begin, end) (shared/interval_tree.py:80-89): any interval with iv.begin < end and
    iv.end > begin."""
    return any(a < end and b > begin for a, b in intervals)


def flank_seq(ref_seq, center1, ref_start1):
    lo = center1 - FLANK - ref_start1
    hi = center1 + FLANK + 1 - ref_start1
    if lo >= 0 and hi <= len(ref_seq):
        return ref_seq[lo:hi]
    out = 'A' * max(0, -lo) + ref_seq[max(0, lo):hi]
    if hi > len(ref_seq):
        out += 'A' * (hi - len(ref_seq))
    return out


def candidate_rows(columns, ref_seq, ref_start1, snp_min_af=0.08, indel_min_af=0.15, min_coverage=4,
                   padding=False, phased=False, confident=None, known=None, head_tail=False):
    """columns: iterable of (pos1, depth_col, bases, hp_csv).  Yields per emitted
    candidate (pos1, ref33, window[list of 33 lists], alt_info_str, depth).

    Mirrors the ring buffer, the >=33-contiguous-rows rule, the in-place
    splice-junction padding on shared rows and the `del depth_dict[center]`
    side effect of create_tensor_pileup.py:463-611.  `confident` = intervals of the --bed_fn tree (see
    confident_tree) or None; `known` = set of 1-based --vcf_fn positions of this chunk or None (:551-556).
    head_tail = --enable_variant_calling_at_sequence_head_and_tail (:467,508-511,613-637): the ring starts (and
    restarts at every gap) as 33 references to ONE all-zero row, so no window is ever incomplete, a padding write
    into a not-yet-overwritten slot lands in that shared row, and the candidates still pending when the stream
    ends are flushed over fresh zero rows without the padding step."""
    n_ch = 30 if phased else 18

    def fresh_ring():
        return [[0] * n_ch] * WIN if head_tail else [None] * WIN
    ring = fresh_ring()
    slot = 0
    prev = -1
    pending = []
    alt_of, depth_of, skip_of = {}, {}, {}
    for pos1, _d, bases, hps in columns:
        hp_list = hps.split(',') if phased else None
        rb = ref_seq[pos1 - ref_start1].upper()
        if prev + 1 != pos1:
            ring = fresh_ring()
            slot = 0
            pending = []
        prev = pos1
        vec, alt_items, depth, pass_af, max_skip = column_vector(
            pos1, bases, ref_seq, ref_start1, hp_list, snp_min_af, indel_min_af)
        if padding:
            skip_of[pos1] = max_skip
            depth_of[pos1] = depth
        if depth > 0 and (snp_min_af == 0.0 or indel_min_af == 0.0):
            pass_af = True
        pass_bed = confident is None or region_in(confident, pos1 - 1,
                                                  pos1 + max_del_length(pos1, bases, ref_seq, ref_start1) + 1)
        if (pass_bed and rb in "ACGT" and pass_af and depth >= min_coverage and known is None) or (
                known is not None and pos1 in known):
            pending.append(pos1)
            alt_of[pos1] = alt_items
            depth_of[pos1] = depth
        ring[slot] = vec
        slot = (slot + 1) % WIN
        if pending and pos1 - pending[0] == FLANK:
            center = pending.pop(0)
            if all(r is not None for r in ring):
                cdepth = depth_of[center]
                window = ring[slot:] + ring[:slot]          # shares the row objects
                if padding:
                    span = range(center - FLANK, center + FLANK + 1)
                    mx_depth = max(depth_of[p] for p in span if p in depth_of)
                    mx_skip = max(skip_of[p] for p in span if p in skip_of)
                    if mx_skip / float(mx_depth) > SKIP_PROP:
                        cref = ref_seq[center - ref_start1]
                        sf = abs(window[FLANK][CH[cref.upper()]])
                        sr = abs(window[FLANK][CH[cref.lower()]])
                        pf = sf / float(sf + sr) if sf + sr > 0 else 0
                        pr = 1 - pf
                        for k in range(WIN):
                            p = center - FLANK + k
                            cur = depth_of[p] if p in depth_of else 0
                            if cur < cdepth * SKIP_PROP and k != FLANK:
                                r = ref_seq[p - ref_start1].upper()
                                window[k][CH[r]] = -1 * int(cdepth * pf)
                                window[k][CH[r.lower()]] = -1 * int(cdepth * pr)
                alt_info = str(cdepth) + '-' + ' '.join('%s %d' % kv for kv in alt_of[center])
                yield (center, flank_seq(ref_seq, center, ref_start1),
                       [list(r) for r in window], alt_info, cdepth)
                del alt_of[center], depth_of[center]
    if head_tail:
        for pos1 in range(prev + 1, prev + FLANK + 1):
            ring[slot] = [0] * n_ch
            slot = (slot + 1) % WIN
            center = pos1 - FLANK
            if center in pending:
                window = ring[slot:] + ring[:slot]
                alt_info = str(depth_of[center]) + '-' + ' '.join('%s %d' % kv for kv in alt_of[center])
                yield (center, flank_seq(ref_seq, center, ref_start1),
                       [list(r) for r in window], alt_info, depth_of[center])


def batch_tensor(window, depth):
    """tensor_generator_from's per-row transform (clair3_rna/utils.py:85-92,120):
    float64 divide by depth/144 when depth > 216, then truncation into int32."""
    t = np.array(window, dtype=np.int32).reshape(-1)
    if depth > 0 and depth > MAX_DEPTH * 1.5:
        t = t / (depth / MAX_DEPTH)
    out = np.empty(t.shape, np.int32)
    out[:] = t
    return out.reshape(WIN, -1)


def run_region(batch, ref_seq, ref_start1, start1, end1, *, snp_min_af=0.08, indel_min_af=0.15,
               min_coverage=4, min_mq=5, excl_flags=2316, padding=False, phased=False,
               pileup_bed=None, confident_bed=None, known=None, head_tail=False):
    """flat reads -> dict of arrays for every emitted candidate of the region.

    pileup_bed: rows (start0, end0) of --extend_bed (mpileup -l); confident_bed: rows of --bed_fn in file order;
    known: iterable of 1-based --vcf_fn positions of this chunk."""
    from .mpileup import mpileup_rows
    pos, ref33, tens, alt, depth = [], [], [], [], []
    cols = mpileup_rows(batch, start1, end1, excl_flags, min_mq, bed=pileup_bed)
    confident = None if confident_bed is None else confident_tree(confident_bed, start1, end1)
    for c, r33, win, ai, d in candidate_rows(cols, ref_seq, ref_start1, snp_min_af, indel_min_af,
                                             min_coverage, padding, phased, confident=confident,
                                             known=None if known is None else set(known), head_tail=head_tail):
        pos.append(c)
        ref33.append(r33)
        tens.append(batch_tensor(win, d))
        alt.append(ai)
        depth.append(d)
    C = 30 if phased else 18
    return dict(pos=np.array(pos, np.int64), ref33=ref33,
                tensor=np.stack(tens) if tens else np.zeros((0, WIN, C), np.int32),
                alt_info=alt, depth=np.array(depth, np.int64))
