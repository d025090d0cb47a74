"""Hand-derived micro-cases for the `samtools mpileup` restatement (oracle/mpileup.py), one per rule of the A0
contract (SURVEY.md 8(a) "A0 notes").

samtools / htslib are absent from this image and from /root/reference, so the expected strings below were written
out BY HAND from the published description of the pileup format, not produced by any code in this repository:

  samtools-mpileup(1), "Pileup Format":
    [M1] "each line consists of chromosome, 1-based coordinate, reference base, the number of reads covering the
          site, read bases, base qualities"
    [M2] "a dot stands for a match to the reference base on the forward strand, a comma for a match on the reverse
          strand ... 'ACGTN' for a mismatch on the forward strand and 'acgtn' for a mismatch on the reverse strand"
         - without -f (the reference's command line has none, create_tensor_pileup.py:436-451) there is no reference
         base to match, the reference column is N and every base is printed as its letter, upper / lower by strand
    [M3] "a '>' or '<' for a reference skip" (forward / reverse strand)
    [M4] "A pattern '\\+[0-9]+[ACGTNacgtn*#]+' indicates there is an insertion between this reference position and
          the next reference position. The length of the insertion is given by the integer in the pattern,
          followed by the inserted sequence."
    [M5] "a pattern '-[0-9]+[ACGTNacgtn]+' represents a deletion from the reference. The deleted bases will be
          presented as '*' in the following lines."  --reverse-del: "Use '#' character for deletions on the
          reverse strand" (without -f the deleted bases print as N / n)
    [M6] "a symbol '^' marks the start of a read. The ASCII of the character following '^' minus 33 gives the
          mapping quality. A symbol '$' marks the end of a read segment."
  samtools-mpileup(1), options:
    [O1] "--ff, --excl-flags STR|INT  Filter flags: skip reads with any of the mask bits set
          [UNMAP,SECONDARY,QCFAIL,DUP]"  (the reference passes 2316 = UNMAP|MUNMAP|SECONDARY|SUPPLEMENTARY)
    [O2] "-q, --min-MQ INT  Minimum mapping quality for an alignment to be used"
    [O3] "-A, --count-orphans  Do not skip anomalous read pairs in variant calling" - without it a paired read
          that is not in a proper pair is skipped
    [O4] "-l, --positions FILE  BED or position list file containing a list of regions or sites where pileup or
          BCF should be generated"
    [O5] "--output-extra STR  Comma-separated list of extra fields ... or SAM tags": one value per read of the
          column, in read order, '*' for a read without the tag
    [O6] "-r, --region STR  Only generate pileup in region" (1-based, inclusive)
  htslib sam.h, bam_pileup1_t: "is_del: 1 iff the base on the padded read is a deletion; is_refskip: 1 iff the
    base on the padded read is part of CIGAR N op; indel: indel length, 0 for no indel, positive for ins and
    negative for del" - set at the LAST position before the indel; BAM_DEF_MASK (UNMAP|SECONDARY|QCFAIL|DUP) is
    dropped by the pileup engine whatever --excl-flags says.

Reads are written as (pos0, flag, mapq, hp, cigar, seq); columns as (pos1, depth, bases[, hp])."""
import numpy as np

from clair3_rna_b200.reads import ReadBatch, encode_seq, parse_cigar
from oracle.mpileup import mpileup_rows, mpileup_text


def batch(*reads):
    recs = [(p, f, q, h, parse_cigar(c), encode_seq(s)) for p, f, q, h, c, s in reads]
    return ReadBatch.from_records("ctg", recs)


def cols(b, s1, e1, **kw):
    return [(p, d, t) for p, d, t, _ in mpileup_rows(b, s1, e1, **kw)]


def test_head_tail_marks_and_strand_case():
    # [M2][M6]: one forward read ACGT at 1-based 11..14 with MAPQ 60 -> '^' + chr(60+33) = '^]' before the first base,
    # '$' after the last; a reverse read prints lower case
    b = batch((10, 0, 60, 0, "4M", "ACGT"), (11, 16, 7, 0, "3M", "TTG"))
    assert cols(b, 1, 100) == [
        (11, 1, "^]A"),
        (12, 2, "C^(t"),          # MAPQ 7 -> chr(40) = '('
        (13, 2, "Gt"),
        (14, 2, "T$g$"),
    ]


def test_mapq_character_is_capped_at_tilde():
    # [M6]: qualities above 93 print as '~' (ASCII 126 is the last printable character)
    b = batch((0, 0, 200, 0, "2M", "AC"))
    assert cols(b, 1, 10) == [(1, 1, "^~A"), (2, 1, "C$")]


def test_deletion_tokens_by_strand():
    # [M5] + --reverse-del: forward 2M2D2M, reverse 2M2D2M at the same place.  The '-2NN' / '-2nn' pattern sits on the
    # last base before the deletion, the deleted positions print '*' (forward) and '#' (reverse) and count in the depth
    b = batch((0, 0, 60, 0, "2M2D2M", "ACGT"), (0, 16, 60, 0, "2M2D2M", "ACGT"))
    assert cols(b, 1, 10) == [
        (1, 2, "^]A^]a"),
        (2, 2, "C-2NNc-2nn"),
        (3, 2, "*#"),
        (4, 2, "*#"),
        (5, 2, "Gg"),
        (6, 2, "T$t$"),
    ]


def test_insertion_tokens_by_strand():
    # [M4]: 2M2I2M, inserted bases AG: '+2AG' after the base BEFORE the insertion, case by strand; the inserted bases
    # occupy no column
    b = batch((4, 0, 60, 0, "2M2I2M", "CCAGTT"), (4, 16, 60, 0, "2M2I2M", "CCAGTT"))
    assert cols(b, 1, 20) == [
        (5, 2, "^]C^]c"),
        (6, 2, "C+2AGc+2ag"),
        (7, 2, "Tt"),
        (8, 2, "T$t$"),
    ]


def test_insertion_then_deletion_at_one_anchor():
    # samtools >= 1.11 prints both patterns on the anchor base: 2M1I1D2M -> 'C+1A-1N'
    b = batch((0, 0, 60, 0, "2M1I1D2M", "CCAGG"))
    assert cols(b, 1, 10) == [(1, 1, "^]C"), (2, 1, "C+1A-1N"), (3, 1, "*"), (4, 1, "G"), (5, 1, "G$")]


def test_reference_skip_columns_are_printed_and_counted():
    # [M3]: 2M3N2M forward and reverse.  No indel pattern before an N; a column under N only ('>' / '<') is still a
    # column with depth = reads spanning it (this keeps Clair's 33-row runs alive across an intron)
    b = batch((0, 0, 60, 0, "2M3N2M", "ACGT"), (1, 16, 60, 0, "1M3N1M", "CG"))
    assert cols(b, 1, 10) == [
        (1, 1, "^]A"),
        (2, 2, "C^]c"),
        (3, 2, "><"),
        (4, 2, "><"),
        (5, 2, "><"),
        (6, 2, "Gg$"),
        (7, 1, "T$"),
    ]


def test_soft_clips_and_leading_insertion_take_no_column():
    # S consumes the query only: 2S2M1S at pos0 5 starts at column 6 with the 3rd base of SEQ
    b = batch((5, 0, 60, 0, "2S2M1S", "TTACG"))
    assert cols(b, 1, 20) == [(6, 1, "^]A"), (7, 1, "C$")]


def test_flag_filters():
    # [O1]: 2316 masks UNMAP(4) MUNMAP(8) SECONDARY(256) SUPPLEMENTARY(2048); QCFAIL(512) and DUP(1024) are dropped by
    # the pileup engine's BAM_DEF_MASK even though 2316 does not name them
    keep = (0, 0, 60, 0, "2M", "AC")
    for flag in (4, 8, 256, 2048, 512, 1024, 256 | 16):
        b = batch(keep, (0, flag, 60, 0, "2M", "GG"))
        assert cols(b, 1, 5) == [(1, 1, "^]A"), (2, 1, "C$")], flag


def test_orphan_rule():
    # [O3]: PAIRED (1) without PROPER_PAIR (2) is skipped; PAIRED|PROPER_PAIR is used
    b = batch((0, 1, 60, 0, "2M", "GG"), (0, 1 | 2, 60, 0, "2M", "AC"), (0, 1 | 2 | 16, 60, 0, "2M", "TT"))
    assert cols(b, 1, 5) == [(1, 2, "^]A^]t"), (2, 2, "C$t$")]


def test_min_mapq():
    # [O2]: --min-MQ 5 keeps MAPQ 5, drops MAPQ 4
    b = batch((0, 0, 4, 0, "2M", "GG"), (0, 0, 5, 0, "2M", "AC"))
    assert cols(b, 1, 5) == [(1, 1, "^&A"), (2, 1, "C$")]       # chr(5+33) = '&'


def test_column_order_is_record_order_and_depth_counts_everything():
    # three reads starting at the same position keep their BAM order in every column they share
    b = batch((0, 0, 60, 0, "3M", "AAA"), (0, 16, 60, 0, "1M1D1M", "CC"), (0, 0, 60, 0, "1M2N1M", "GG"))
    assert cols(b, 1, 5) == [
        (1, 3, "^]A^]c-1n^]G"),
        (2, 3, "A#>"),
        (3, 3, "A$c$>"),
        (4, 1, "G$"),
    ]


def test_region_bounds_do_not_change_column_content():
    # [O6]: -r ctg:3-4 prints only columns 3 and 4; the '^' / '$' marks belong to columns 1 and 6 and do not move
    b = batch((0, 0, 60, 0, "6M", "ACGTAC"))
    assert cols(b, 3, 4) == [(3, 1, "G"), (4, 1, "T")]


def test_gaps_print_no_column():
    # no -a: positions without a read have no line
    b = batch((0, 0, 60, 0, "2M", "AC"), (10, 0, 60, 0, "1M", "G"))
    assert [c[0] for c in cols(b, 1, 50)] == [1, 2, 11]


def test_output_extra_hp_column():
    # [O5]: HP:i:1 -> '1', no tag -> '*', one value per read in column order
    b = batch((0, 0, 60, 1, "2M", "AC"), (0, 16, 60, 0, "2M", "AC"), (1, 0, 60, 2, "2M", "GG"))
    rows = list(mpileup_rows(b, 1, 5))
    assert [(p, t, h) for p, _, t, h in rows] == [(1, "^]A^]a", "1,*"), (2, "C$c$^]G", "1,*,2"), (3, "G$", "2")]
    lines = list(mpileup_text(b, "ctg", 1, 5, with_hp=True))
    assert lines[1] == "ctg\t2\tN\t3\tC$c$^]G\tIII\t1,*,2"         # [M1]: six columns + the extra field


def test_positions_bed_filter():
    # [O4]: -l BED (0-based half open rows) keeps columns 2..3 and 5 only; content unchanged
    b = batch((0, 0, 60, 0, "6M", "ACGTAC"))
    assert cols(b, 1, 10, bed=[(1, 3), (4, 5)]) == [(2, 1, "C"), (3, 1, "G"), (5, 1, "A")]


def test_nt16_ambiguity_codes_print_their_letter():
    # SEQ letters other than ACGT print as they are (N, R ...): Clair's parser ignores them for the counts
    b = batch((0, 0, 60, 0, "3M", "ANR"), (0, 16, 60, 0, "3M", "ANR"))
    assert cols(b, 1, 5) == [(1, 2, "^]A^]a"), (2, 2, "Nn"), (3, 2, "R$r$")]


def test_reference_parser_reads_the_hand_written_columns():
    """The consumer of this text is the reference's generate_tensor (create_tensor_pileup.py:85-302): the same
    hand-written column, parsed by the oracle's restatement of it, gives the counts one reads off the string.
    Channel order (param_p.py:31): A C G T I I1 D D1 * a c g t i i1 d d1 #."""
    from oracle import pileup_oracle
    b = batch((0, 0, 60, 0, "2M2D2M", "ACGT"), (0, 16, 60, 0, "2M2D2M", "ACGT"), (1, 0, 60, 0, "1M2I3M", "CAGGGT"),
              (1, 0, 60, 0, "4M", "TGGT"))
    rows = {p: t for p, _, t in cols(b, 1, 10)}
    assert rows[2] == "C-2NNc-2nn^]C+2AG^]T"
    ref = "ACGGGTAAAA"
    vec, alt, depth, pass_af, max_skip = pileup_oracle.column_vector(2, rows[2], ref, 1, None, 0.08, 0.15)
    # forward C x2 (reference base: stored as minus the forward ACGT total), forward T x1, reverse c x1,
    # one insertion (forward), one deletion per strand; '^' counted twice for max_skip
    assert depth == 4 and pass_af and max_skip == 2
    assert vec == [0, -3, 0, 1, 1, 1, 1, 1, 0, 0, -1, 0, 0, 0, 0, 1, 1, 0]
    assert dict(alt) == {"DGG": 2, "ICAG": 1, "XT": 1}          # R = max(0, 4 - 2 del - 1 ins - 1 alt) = 0: no entry
    # position 3 lies under both deletions: '*' forward, '#' reverse, both count in the depth
    vec3, alt3, depth3, _, _ = pileup_oracle.column_vector(3, rows[3], ref, 1, None, 0.08, 0.15)
    assert rows[3] == "*#GG" and depth3 == 4 and vec3[8] == 1 and vec3[17] == 1 and vec3[2] == -2
