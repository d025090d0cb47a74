"""Deterministic synthetic spliced long reads for the five BASELINE.json configs.

There is no BAM/FASTA fixture in the reference and no network, so every
parity test and bench line runs on reads made here.  The generator follows the
restrictions SURVEY.md §8(a) "A0 notes" lists so that every htslib >= 1.10
behaves identically on them: ops in {M,I,D,N,S}, first/last non-clip op M, no
two of {I,D,N} adjacent, SEQ always present, flags in {0,16} plus filtered
decoys (256 / 2048 / low MAPQ), never PAIRED/QCFAIL/DUP.

Reference bases come from a counter-based RNG keyed by (seed, contig, block) so
a 3 Gb genome never has to exist in memory.
"""
from __future__ import annotations

from dataclasses import dataclass, field
import math
import numpy as np

from . import params as P
from .reads import ReadBatch

BLOCK = 1 << 20
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_NT16_OF_IDX = np.array([1, 2, 4, 8], dtype=np.uint8)          # A C G T -> nt16

GRCH38 = [("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555),
          ("chr5", 181538259), ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636),
          ("chr9", 138394717), ("chr10", 133797422), ("chr11", 135086622), ("chr12", 133275309),
          ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189), ("chr16", 90338345),
          ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
          ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415)]


@dataclass
class SynthConfig:
    name: str
    platform: str                    # full platform flag, e.g. ont_dorado_drna004
    contigs: list                    # [(name, length)]
    seed: int
    genes_per_mb: float = 40.0
    depth: float = 20.0
    sense_only: bool = True          # dRNA: every read on its gene's strand
    sub: float = 0.03
    ins: float = 0.02
    dele: float = 0.03
    phased: bool = False             # HP tags + C=30
    padding: bool = False            # --enable_padding_in_splice_junction_regions
    hi_depth_genes: int = 0          # first N genes of a contig get `hi_depth`
    hi_depth: float = 500.0
    min_exon: int = 30
    near_junction_variants: bool = False
    read_len_median: float = 900.0


def config(idx: int, scale: float = 1.0) -> SynthConfig:
    """The five configs of BASELINE.json / SURVEY.md §8(d).  `scale` < 1 shrinks
    contig lengths (tests); 1.0 is the full size."""
    def L(n):
        return max(20000, int(n * scale))
    if idx == 1:
        return SynthConfig("cfg1_ont_drna004_1mb", "ont_dorado_drna004", [("chr1", L(1_000_000))], 20260001,
                           genes_per_mb=40, depth=20, sense_only=True)
    if idx == 2:
        return SynthConfig("cfg2_ont_r10_cdna_chr20", "ont_r10_dorado_cdna", [("chr20", L(64_444_167))], 20260002,
                           genes_per_mb=8.6, depth=30, sense_only=False, sub=0.01, ins=0.005, dele=0.01)
    if idx == 3:
        return SynthConfig("cfg3_hifi_mas_pad", "hifi_mas_minimap2", [("chr1", L(5_000_000))], 20260003,
                           genes_per_mb=12, depth=30, sense_only=False, sub=0.001, ins=0.0005, dele=0.0005,
                           padding=True, hi_depth_genes=max(1, int(20 * min(1.0, scale * 4))), hi_depth=500,
                           min_exon=20, near_junction_variants=True, read_len_median=1500)
    if idx == 4:
        return SynthConfig("cfg4_hifi_sequel2_phased", "hifi_sequel2_minimap2", [("chr1", L(5_000_000))], 20260004,
                           genes_per_mb=12, depth=40, sense_only=False, sub=0.001, ins=0.0005, dele=0.0005,
                           phased=True, read_len_median=1500)
    if idx == 5:
        return SynthConfig("cfg5_ont_drna004_wgs", "ont_dorado_drna004", [(n, L(l)) for n, l in GRCH38], 20260005,
                           genes_per_mb=6.5, depth=20, sense_only=True)
    raise ValueError(idx)


# ----------------------------------------------------------------- reference
def ref_block(seed: int, contig_idx: int, block: int) -> np.ndarray:
    g = np.random.Generator(np.random.Philox(key=[seed, (contig_idx << 20) | block]))
    return _ACGT[g.integers(0, 4, BLOCK, dtype=np.uint8)]


class Reference:
    """Lazily generated reference of one config (ASCII upper-case bytes)."""

    def __init__(self, cfg: SynthConfig):
        self.cfg = cfg
        self.names = [n for n, _ in cfg.contigs]
        self.lengths = dict(cfg.contigs)
        self._cache = {}

    def fetch(self, contig: str, start0: int, end0: int) -> np.ndarray:
        """bases of [start0, end0) clipped to the contig."""
        ci = self.names.index(contig)
        start0, end0 = max(0, start0), min(self.lengths[contig], end0)
        if end0 <= start0:
            return np.zeros(0, np.uint8)
        parts = []
        for b in range(start0 // BLOCK, (end0 - 1) // BLOCK + 1):
            key = (ci, b)
            if key not in self._cache:
                if len(self._cache) > 96:
                    self._cache.clear()
                self._cache[key] = ref_block(self.cfg.seed, ci, b)
            blk = self._cache[key]
            lo, hi = max(start0, b * BLOCK) - b * BLOCK, min(end0, (b + 1) * BLOCK) - b * BLOCK
            parts.append(blk[lo:hi])
        return np.concatenate(parts)

    def fetch_str(self, contig: str, start0: int, end0: int) -> str:
        return self.fetch(contig, start0, end0).tobytes().decode("ascii")


# --------------------------------------------------------------------- genes
@dataclass
class Gene:
    strand: int                      # 0 '+', 1 '-'
    exons: list                      # [(start0, end0)] ascending
    depth: float
    # variants in genome coordinates
    snps: dict = field(default_factory=dict)       # pos0 -> (alt_idx, hap_mask(1|2|3), af)   af<0: by haplotype
    indels: dict = field(default_factory=dict)     # pos0 (anchor base) -> (kind 'I'|'D', length, hap_mask, ins_codes)
    skip_exon: int = -1              # isoform B skips this internal exon


def make_genes(cfg: SynthConfig, contig_idx: int, ref: Reference) -> list:
    name, length = cfg.contigs[contig_idx]
    rng = np.random.default_rng([cfg.seed, contig_idx, 7])
    n_genes = max(1, int(round(cfg.genes_per_mb * length / 1e6)))
    slot = length / n_genes
    genes = []
    for gi in range(n_genes):
        n_ex = int(rng.integers(3, 13))
        ex_len = np.clip(np.exp(rng.normal(math.log(150), 0.8, n_ex)), cfg.min_exon, 1500).astype(np.int64)
        introns = rng.integers(500, 20001, n_ex - 1)
        span = int(ex_len.sum() + introns.sum())
        room = int(slot * 0.9) - 200
        if span > room:
            over = span - room
            cut = np.minimum(introns - 80, (introns * (over / max(1, introns.sum()) + 0.02)).astype(np.int64) + 1)
            introns = introns - np.maximum(cut, 0)
            span = int(ex_len.sum() + introns.sum())
            while span > room and n_ex > 2:
                n_ex -= 1
                ex_len, introns = ex_len[:n_ex], introns[:n_ex - 1]
                span = int(ex_len.sum() + introns.sum())
        if span > room:
            continue
        g0 = int(gi * slot) + 100 + int(rng.integers(0, max(1, room - span)))
        exons, cur = [], g0
        for k in range(n_ex):
            exons.append((cur, cur + int(ex_len[k])))
            cur += int(ex_len[k]) + (int(introns[k]) if k < n_ex - 1 else 0)
        depth = cfg.hi_depth if gi < cfg.hi_depth_genes else cfg.depth * float(np.exp(rng.normal(0, 0.4)))
        g = Gene(int(rng.integers(0, 2)), exons, depth)
        if n_ex >= 4 and rng.random() < 0.5:
            g.skip_exon = int(rng.integers(1, n_ex - 1))
        _plant_variants(cfg, g, rng, ref, name)
        genes.append(g)
    return genes


def _plant_variants(cfg, g: Gene, rng, ref: Reference, contig: str) -> None:
    taken = set()

    def free(p, rad):
        return all((p + d) not in taken for d in range(-rad, rad + 1))

    for (s, e) in g.exons:
        n = e - s
        seq = ref.fetch(contig, s, e)
        # SNPs: het 1/kb, hom 1/3kb, editing-like low AF 1/2kb
        for kind, rate in (("het", 1 / 1000), ("hom", 1 / 3000), ("edit", 1 / 2000)):
            k = rng.poisson(rate * n * (3 if cfg.near_junction_variants else 1))
            for _ in range(k):
                if cfg.near_junction_variants and rng.random() < 0.6:
                    off = int(rng.integers(0, min(12, n)))
                    p = s + off if rng.random() < 0.5 else e - 1 - off
                else:
                    p = s + int(rng.integers(0, n))
                if not free(p, 0):
                    continue
                refi = int(np.searchsorted(_ACGT, seq[p - s]))
                if kind == "edit":
                    # A>G on the sense strand (T>C seen on '+' for '-' genes)
                    want, alt = (0, 2) if g.strand == 0 else (3, 1)
                    if refi != want:
                        continue
                    g.snps[p] = (alt, 3, float(rng.uniform(0.1, 0.4)))
                else:
                    alt = (refi + int(rng.integers(1, 4))) % 4
                    g.snps[p] = (alt, 3 if kind == "hom" else int(rng.integers(1, 3)), -1.0)
                taken.add(p)
        # het indels 1/5kb, anchored >= 3 bases inside the exon
        k = rng.poisson(n / 5000 * (4 if cfg.near_junction_variants else 1))
        for _ in range(k):
            ln = int(min(rng.geometric(0.7), 8))
            if n < 2 * 3 + ln + 2:
                continue
            if cfg.near_junction_variants and rng.random() < 0.6:
                p = s + 2 + int(rng.integers(0, 6)) if rng.random() < 0.5 else e - 4 - ln - int(rng.integers(0, 6))
                p = min(max(p, s + 2), e - 4 - ln)
            else:
                p = s + 2 + int(rng.integers(0, n - 5 - ln))
            if not free(p, ln + 3):
                continue
            if rng.random() < 0.5:
                g.indels[p] = ("I", ln, int(rng.integers(1, 3)), _NT16_OF_IDX[rng.integers(0, 4, ln)])
                taken.update(range(p - 2, p + 3))
            else:
                g.indels[p] = ("D", ln, int(rng.integers(1, 3)), None)
                taken.update(range(p - 2, p + ln + 3))


# --------------------------------------------------------------------- reads
def _read_from_gene(cfg, g: Gene, rng, ref: Reference, contig: str, tseq_cache: dict):
    """one alignment record (pos0, flag, mapq, hp, cigar, codes) or None."""
    iso_b = g.skip_exon >= 0 and rng.random() < 0.3
    exons = [x for i, x in enumerate(g.exons) if not (iso_b and i == g.skip_exon)]
    key = (id(g), iso_b)
    if key not in tseq_cache:
        tpos = np.concatenate([np.arange(s, e, dtype=np.int64) for s, e in exons])
        tbase = np.concatenate([ref.fetch(contig, s, e) for s, e in exons])
        tidx = np.searchsorted(_ACGT, tbase).astype(np.uint8)
        bound = np.zeros(tpos.size, bool)        # True where the next transcript base is across a junction
        bound[:-1] = (tpos[1:] - tpos[:-1]) != 1
        tseq_cache[key] = (tpos, tidx, bound)
    tpos, tidx, bound = tseq_cache[key]
    T = tpos.size
    rl = int(min(max(50, math.exp(rng.normal(math.log(cfg.read_len_median), 0.5))), T))
    a = int(rng.integers(0, T - rl + 1))
    b = a + rl
    if rl < 12:
        return None
    hap = int(rng.integers(0, 2))                # haplotype 0/1 -> HP 1/2
    pos_g = tpos[a:b]
    base = tidx[a:b].copy()
    jn = bound[a:b].copy()
    jn[-1] = False
    n = rl
    # planted SNPs
    if g.snps:
        for i in np.nonzero(np.isin(pos_g, np.fromiter(g.snps.keys(), np.int64, len(g.snps))))[0]:
            alt, mask, af = g.snps[int(pos_g[i])]
            if (af < 0 and (mask >> hap) & 1) or (af >= 0 and rng.random() < af):
                base[i] = alt
    # sequencing substitutions / N
    u = rng.random(n)
    sub = u < cfg.sub
    base[sub] = (base[sub] + rng.integers(1, 4, int(sub.sum()))) % 4
    codes = _NT16_OF_IDX[base]
    codes[(u > 1 - 0.001)] = 15                  # 'N' base calls
    # indel events: index i = anchor base (event follows base i)
    ev = {}
    blocked = np.zeros(n + 1, bool)
    blocked[:2] = True
    blocked[n - 3:] = True
    jidx = np.nonzero(jn)[0]
    for j in jidx:                               # keep {I,D} two bases away from N
        blocked[max(0, j - 2): j + 3] = True
    if g.indels:
        for i in np.nonzero(np.isin(pos_g, np.fromiter(g.indels.keys(), np.int64, len(g.indels))))[0]:
            kind, ln, mask, ins_codes = g.indels[int(pos_g[i])]
            if not (mask >> hap) & 1 or rng.random() < 0.05:
                continue
            span = ln if kind == "D" else 0
            if i + span + 2 >= n or blocked[i: i + span + 3].any() or jn[i: i + span + 1].any():
                continue
            ev[int(i)] = (kind, ln, ins_codes)
            blocked[max(0, i - 2): i + span + 3] = True
    ue = rng.random(n)
    for i in np.nonzero(ue < cfg.ins + cfg.dele)[0]:
        i = int(i)
        ln = int(min(rng.geometric(0.7), 6))
        if ue[i] < cfg.ins:
            if blocked[i: i + 3].any():
                continue
            ev[i] = ("I", ln, _NT16_OF_IDX[rng.integers(0, 4, ln)])
            blocked[max(0, i - 2): i + 3] = True
        else:
            if i + ln + 2 >= n or blocked[i: i + ln + 3].any() or jn[i: i + ln + 1].any():
                continue
            ev[i] = ("D", ln, None)
            blocked[max(0, i - 2): i + ln + 3] = True
    # walk -> CIGAR + SEQ
    cigar, seq_parts = [], []
    if rng.random() < 0.3:
        sl = int(rng.integers(1, 31))
        cigar.append((sl, P.CIG_S))
        seq_parts.append(_NT16_OF_IDX[rng.integers(0, 4, sl)])
    m_run, m_start, i = 0, 0, 0
    while i < n:
        m_run += 1
        e = ev.get(i)
        if e is not None:
            kind, ln, ins_codes = e
            cigar.append((m_run, P.CIG_M))
            seq_parts.append(codes[m_start: i + 1])
            m_run = 0
            if kind == "I":
                cigar.append((ln, P.CIG_I))
                seq_parts.append(ins_codes)
                m_start = i + 1
            else:
                cigar.append((ln, P.CIG_D))
                i += ln
                m_start = i + 1
        elif jn[i]:
            cigar.append((m_run, P.CIG_M))
            seq_parts.append(codes[m_start: i + 1])
            m_run = 0
            cigar.append((int(pos_g[i + 1] - pos_g[i] - 1), P.CIG_N))
            m_start = i + 1
        i += 1
    if m_run:
        cigar.append((m_run, P.CIG_M))
        seq_parts.append(codes[m_start: n])
    if rng.random() < 0.3:
        sl = int(rng.integers(1, 31))
        cigar.append((sl, P.CIG_S))
        seq_parts.append(_NT16_OF_IDX[rng.integers(0, 4, sl)])
    strand = g.strand if cfg.sense_only else int(rng.integers(0, 2))
    flag = P.FLAG_REVERSE if strand else 0
    r = rng.random()
    if r < 0.02:
        flag |= P.FLAG_SECONDARY
    elif r < 0.04:
        flag |= P.FLAG_SUPPLEMENTARY
    mapq = 60 if rng.random() < 0.9 else int(rng.integers(0, 21))
    hp = 0
    if cfg.phased and rng.random() < 0.9:
        hp = hap + 1
    return (int(pos_g[0]), flag, mapq, hp, cigar, np.concatenate(seq_parts))


def make_contig_reads(cfg: SynthConfig, contig_idx: int, ref: Reference | None = None,
                      max_genes: int | None = None) -> ReadBatch:
    """All alignment records of one contig, coordinate sorted."""
    ref = ref or Reference(cfg)
    name, _ = cfg.contigs[contig_idx]
    genes = make_genes(cfg, contig_idx, ref)
    if max_genes is not None:
        genes = genes[:max_genes]
    rng = np.random.default_rng([cfg.seed, contig_idx, 11])
    recs = []
    cache = {}
    for g in genes:
        exonic = sum(e - s for s, e in g.exons)
        mean_len = min(exonic, cfg.read_len_median * 1.13)
        n_reads = int(rng.poisson(g.depth * exonic / mean_len))
        for _ in range(n_reads):
            r = _read_from_gene(cfg, g, rng, ref, name, cache)
            if r is not None:
                recs.append(r)
        cache.clear()
    recs.sort(key=lambda r: r[0])
    return ReadBatch.from_records(name, recs)


def chunk_list(cfg: SynthConfig, chunk_size: int = P.CHUNK_SIZE) -> list:
    """[(contig, chunk_id1, chunk_num)] as run_clair3_rna builds it
    (/root/reference/run_clair3_rna:378-381,441-449)."""
    out = []
    for name, length in cfg.contigs:
        num = max(1, -(-length // chunk_size))
        out.extend((name, i + 1, num) for i in range(num))
    return out


def chunk_geometry(contig_len: int, chunk_id1: int, chunk_num: int):
    """(ctg_start, ctg_end, read_start1, read_end1, ref_start1, ref_end1) exactly as
    /root/reference/src/create_tensor_pileup.py:390-418 computes them."""
    cid = chunk_id1 - 1
    chunk_size = contig_len // chunk_num + 1 if contig_len % chunk_num else contig_len // chunk_num
    ctg_start = chunk_size * cid
    ctg_end = ctg_start + chunk_size
    ext_s = max(1, ctg_start - P.NO_OF_POSITIONS)
    ext_e = ctg_end + P.NO_OF_POSITIONS
    ref_s = max(1, ctg_start - P.EXPAND_REFERENCE_REGION)
    ref_e = ctg_end + P.EXPAND_REFERENCE_REGION
    return ctg_start, ctg_end, ext_s, ext_e, ref_s, ref_e
