"""Region modes of the per-chunk producer: `--bed_fn` (confident regions), `--extend_bed` (the BED handed to
`samtools mpileup -l`), `--vcf_fn` (genotyping at known sites) and `--ctgStart/--ctgEnd`, restated from
/root/reference/src/create_tensor_pileup.py:373-418, 446-451, 480-481, 551-556, shared/interval_tree.py:8-89,
shared/utils.py:196-216 and the BED/VCF splitters of run_clair3_rna:231-289.

The host decides the geometry of a chunk (which reads, which reference slice) and hands three optional site
filters to the device through `c3r_submit_chunk_filtered` (include/c3r_b200.h):

  pileup_bed   rows exist only at positions inside these intervals          (mpileup -l)
  confident    a candidate needs [pos-1, pos+max_del_length+1) to overlap   (is_region_in on the --bed_fn tree)
  known        candidates are exactly these positions, whatever their AF    (pos in known_variants_set)

Intervals are 0-based half-open and travel as sorted, disjoint, non-touching int32 pairs.
"""
from __future__ import annotations

import dataclasses
import gzip
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import params as P

Rows = List[Tuple[int, int]]


def _open_text(path):
    with open(path, "rb") as fp:
        magic = fp.read(2)
    return gzip.open(path, "rt") if magic == b"\x1f\x8b" else open(path, "rt")


def read_bed_rows(path: str, contig: Optional[str] = None):
    """BED rows in file order.  contig given -> [(start0, end0)], else {contig: [(start0, end0)]}.
    Same row checks as bed_tree_from (interval_tree.py:44-58): '#' rows skipped, whitespace separated."""
    out = [] if contig is not None else {}
    with _open_text(path) as fp:
        for row_id, row in enumerate(fp):
            if not row.strip() or row[0] == '#':
                continue
            c = row.split()
            a, b = int(c[1]), int(c[2])
            if b < a or a < 0 or b < 0:
                raise ValueError("[ERROR] Invalid bed input in %d-th row %s %d %d" % (row_id + 1, c[0], a, b))
            if contig is not None:
                if c[0] == contig:
                    out.append((a, b))
            else:
                out.setdefault(c[0], []).append((a, b))
    return out


def read_known_positions(path: str, contig: Optional[str] = None):
    """vcf_candidates_from (shared/utils.py:196-216): sorted distinct 1-based POS of the contig
    (contig None -> {contig: [pos]})."""
    out = set() if contig is not None else {}
    with _open_text(path) as fp:
        for row in fp:
            if not row.strip() or row[0] == '#':
                continue
            c = row.split(maxsplit=3)
            if contig is not None:
                if c[0] == contig:
                    out.add(int(c[1]))
            else:
                out.setdefault(c[0], set()).add(int(c[1]))
    if contig is not None:
        return sorted(out)
    return {k: sorted(v) for k, v in out.items()}


def extend_bed_rows(rows: Rows, by: int = P.NO_OF_POSITIONS) -> Rows:
    """split_extend_bed (run_clair3_rna:268-289): every row widened by 33 bp on both sides."""
    return [(max(0, a - by), max(0, b + by)) for a, b in rows]


def extend_known_rows(positions: Sequence[int], by: int = P.NO_OF_POSITIONS) -> Rows:
    """split_extend_vcf (run_clair3_rna:231-265): [pos-1-33, pos+33) per site; sites closer than 33 bp to the
    contig start get no row (and hence no pileup column)."""
    return [(p - 1 - by, p + by) for p in positions if p - 1 - by >= 0]


def merge_intervals(rows: Rows) -> np.ndarray:
    """sorted, disjoint, non-touching int32 pairs; empty rows dropped (they contain no position)."""
    iv = sorted((int(a), int(b)) for a, b in rows if b > a)
    out = []
    for a, b in iv:
        if out and a <= out[-1][1]:
            out[-1][1] = max(out[-1][1], b)
        else:
            out.append([a, b])
    return np.asarray(out, np.int32).reshape(-1, 2)


def confident_intervals(rows: Rows, extend_start1: int, extend_end1: int) -> np.ndarray:
    """the --bed_fn tree of one chunk (bed_tree_from with bed_ctg_start/end, interval_tree.py:60-70): rows
    outside the read region are left out, empty rows become 1 bp."""
    keep = []
    for a, b in rows:
        if extend_start1 and extend_end1 and (b < extend_start1 or a > extend_end1):
            continue
        keep.append((a, b + 1 if a == b else b))
    return merge_intervals(keep)


@dataclasses.dataclass
class ChunkPlan:
    start1: int                      # reads / mpileup region, 1-based inclusive
    end1: int
    ref_start1: int                  # reference slice
    ref_end1: int
    pileup_bed: Optional[np.ndarray] = None
    confident: Optional[np.ndarray] = None
    known: Optional[np.ndarray] = None       # sorted 1-based positions

    def site_filter(self):
        if self.pileup_bed is None and self.confident is None and self.known is None:
            return None
        return dict(pileup_bed=self.pileup_bed, confident=self.confident, known=self.known)


def plan_chunk(contig_len: int, *, chunk_id: Optional[int] = None, chunk_num: Optional[int] = None,
               ctg_start: Optional[int] = None, ctg_end: Optional[int] = None,
               extend_rows: Optional[Rows] = None, confident_rows: Optional[Rows] = None,
               known_positions: Optional[Sequence[int]] = None) -> Optional[ChunkPlan]:
    """Geometry and filters of one producer call, or None when the reference would return without output.

    chunk_id is 1-based.  extend_rows / confident_rows are the contig's rows of --extend_bed / --bed_fn,
    known_positions the contig's sorted --vcf_fn sites (create_tensor_pileup.py:373-418)."""
    cid = chunk_id - 1 if chunk_id else None
    known = None
    if confident_rows is None and cid is not None:
        size = contig_len // chunk_num + 1 if contig_len % chunk_num else contig_len // chunk_num
        ctg_start = size * cid
        ctg_end = ctg_start + size
    if confident_rows is not None and cid is not None:
        if not extend_rows:
            raise ValueError("--bed_fn with --chunk_id needs the contig in --extend_bed (its span sets the chunk geometry)")
        bed_start = min(a for a, _ in extend_rows)
        bed_end = max(b for _, b in extend_rows)
        span = bed_end - bed_start
        size = span // chunk_num + 1 if span % chunk_num else span // chunk_num
        ctg_start = bed_start + 1 + size * cid
        ctg_end = ctg_start + size
    if known_positions is not None:
        known = []
        if cid is not None:                          # the site list is only loaded for chunked calls (:397-405)
            total = len(known_positions)
            per = total // chunk_num if total % chunk_num == 0 else total // chunk_num + 1
            known = list(known_positions[cid * per: cid * per + per])
            if not known:
                return None
            ctg_start, ctg_end = min(known), max(known)
    if ctg_start is not None and ctg_end is not None:
        start1 = max(1, ctg_start - P.NO_OF_POSITIONS)
        end1 = ctg_end + P.NO_OF_POSITIONS
        ref_start1 = max(1, ctg_start - P.EXPAND_REFERENCE_REGION)
        ref_end1 = ctg_end + P.EXPAND_REFERENCE_REGION
    else:
        if confident_rows is not None:
            raise ValueError("--bed_fn needs --chunk_id/--chunk_num or --ctgStart/--ctgEnd")
        start1, end1 = 1, contig_len + P.NO_OF_POSITIONS
        ref_start1, ref_end1 = 1, contig_len
    plan = ChunkPlan(start1, end1, ref_start1, ref_end1)
    if extend_rows is not None:
        plan.pileup_bed = merge_intervals(extend_rows)
    if confident_rows is not None:
        plan.confident = confident_intervals(confident_rows, start1, end1)
    if known is not None:
        plan.known = np.asarray(sorted(set(known)), np.int32)
    return plan
