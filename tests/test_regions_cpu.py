"""Host logic of the region modes (clair3_rna_b200/regions.py).  The geometry itself is pinned against the
reference by the bed_regions / bed_pad_hifi / known_sites goldens (tests/test_oracle_golden.py runs the oracle
through regions.plan_chunk); this file covers the parsers and the edge cases."""
import gzip

import numpy as np
import pytest

from clair3_rna_b200 import regions
from clair3_rna_b200.synth import chunk_geometry


def test_bed_reader_plain_gzip_and_separators(tmp_path):
    text = "# header\nchr1\t10\t20\textra\nchr2 5 6\nchr1 30 30\n\nchr1\t15\t40\n"
    plain = tmp_path / "a.bed"
    plain.write_text(text)
    gz = tmp_path / "a.bed.gz"
    with gzip.open(gz, "wt") as fp:
        fp.write(text)
    for path in (plain, gz):
        assert regions.read_bed_rows(str(path), "chr1") == [(10, 20), (30, 30), (15, 40)]
        assert regions.read_bed_rows(str(path)) == {"chr1": [(10, 20), (30, 30), (15, 40)], "chr2": [(5, 6)]}
    bad = tmp_path / "bad.bed"
    bad.write_text("chr1\t9\t3\n")
    with pytest.raises(ValueError):
        regions.read_bed_rows(str(bad), "chr1")


def test_known_sites_reader(tmp_path):
    vcf = tmp_path / "k.vcf"
    vcf.write_text("##fileformat=VCFv4.2\n#CHROM\tPOS\nchr1\t100\t.\tA\tC\nchr1\t7\t.\tG\tT\nchr2\t3\t.\tA\tG\nchr1\t100\t.\tA\tT\n")
    assert regions.read_known_positions(str(vcf), "chr1") == [7, 100]
    assert regions.read_known_positions(str(vcf)) == {"chr1": [7, 100], "chr2": [3]}


def test_merge_and_extend():
    rows = [(50, 60), (10, 20), (20, 25), (55, 58), (70, 70), (24, 30)]
    assert regions.merge_intervals(rows).tolist() == [[10, 30], [50, 60]]
    assert regions.merge_intervals([]).shape == (0, 2)
    assert regions.extend_bed_rows([(10, 20), (100, 130)]) == [(0, 53), (67, 163)]
    assert regions.extend_known_rows([5, 33, 34, 35, 500]) == [(0, 67), (1, 68), (466, 533)]
    # empty rows of the confident BED become 1 bp; rows outside the read region are left out
    assert regions.confident_intervals([(70, 70), (10, 20), (500, 600)], 1, 100).tolist() == [[10, 20], [70, 71]]


@pytest.mark.parametrize("length,num", [(1000, 1), (999983, 7), (64444167, 13)])
def test_plain_chunks_equal_chunk_geometry(length, num):
    for cid in range(1, num + 1):
        _, _, s, e, rs, re_ = chunk_geometry(length, cid, num)
        plan = regions.plan_chunk(length, chunk_id=cid, chunk_num=num)
        assert (plan.start1, plan.end1, plan.ref_start1, plan.ref_end1) == (s, e, rs, re_)
        assert plan.site_filter() is None


def test_range_and_whole_contig():
    plan = regions.plan_chunk(5000, ctg_start=100, ctg_end=900)
    assert (plan.start1, plan.end1) == (67, 933)
    assert plan.ref_start1 == 1 and plan.ref_end1 > 900
    plan = regions.plan_chunk(5000)
    assert (plan.start1, plan.end1, plan.ref_start1, plan.ref_end1) == (1, 5033, 1, 5000)
    with pytest.raises(ValueError):                  # the reference has no read region to filter the BED by
        regions.plan_chunk(5000, confident_rows=[(1, 2)], extend_rows=[(0, 35)])


def test_bed_chunks_follow_the_span_of_the_extended_bed():
    conf = [(1000, 1100), (4000, 4500)]
    ext = regions.extend_bed_rows(conf)
    a = regions.plan_chunk(10000, chunk_id=1, chunk_num=2, extend_rows=ext, confident_rows=conf)
    b = regions.plan_chunk(10000, chunk_id=2, chunk_num=2, extend_rows=ext, confident_rows=conf)
    span = (4500 + 33) - (1000 - 33)
    size = span // 2 + (1 if span % 2 else 0)
    assert a.start1 == max(1, 967 + 1 - 33) and a.end1 == 968 + size + 33
    assert b.start1 == 968 + size - 33
    assert a.pileup_bed.tolist() == [[967, 1133], [3967, 4533]]
    assert a.confident.tolist() == [[1000, 1100]] and b.confident.tolist() == [[4000, 4500]]
    with pytest.raises(ValueError):
        regions.plan_chunk(10000, chunk_id=1, chunk_num=2, confident_rows=conf)


def test_known_site_chunks():
    sites = [10, 50, 90, 400, 410, 800, 801]
    plans = [regions.plan_chunk(1000, chunk_id=c, chunk_num=4, known_positions=sites,
                                extend_rows=regions.extend_known_rows(sites)) for c in range(1, 5)]
    assert [None if p is None else p.known.tolist() for p in plans] == [[10, 50], [90, 400], [410, 800], [801]]
    assert plans[0].start1 == 1 and plans[0].end1 == 83
    assert regions.plan_chunk(1000, chunk_id=3, chunk_num=3, known_positions=[5, 6]) is None
    # without chunking the reference never loads the site list: nothing can be a candidate
    p = regions.plan_chunk(1000, known_positions=sites)
    assert p.known.size == 0 and p.site_filter() is not None
