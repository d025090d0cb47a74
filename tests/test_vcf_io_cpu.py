"""bgzip + tabix writer (c3r_vcf_write_bgzf): the file is valid gzip with the BGZF extra field and EOF block, the
index answers region queries exactly (checked with an independent reader of the two formats in vcf_io.py)."""
import gzip
import os

import numpy as np
import pytest

from clair3_rna_b200 import vcf_io
from tests.golden import sort_fixture
from clair3_rna_b200 import sharder

HEADER = "".join(sort_fixture.HEADER)


def merged_rows():
    rows = sort_fixture.chunk_rows()
    shards = [rows[k] for k in sorted(rows, key=lambda k: (sort_fixture.CONTIGS.index(k[0]), k[1]))]
    return sharder.sort_vcf(shards, sort_fixture.CONTIGS, qual=8, show_ref=True)[0]


def test_round_trip_and_region_queries(tmp_path):
    rows = merged_rows()
    # enough text for several BGZF blocks: the fixture rows plus far-away copies of chr1's
    extra = []
    for k in range(1, 400):
        for r in rows:
            c = r.split("\t")
            if c[0] == "chr1":
                c[1] = str(int(c[1]) + 20000 * k)
                extra.append("\t".join(c))
    i1 = max(i for i, r in enumerate(rows) if r.startswith("chr1\t")) + 1
    rows = rows[:i1] + extra + rows[i1:]
    path = str(tmp_path / "out.vcf.gz")
    vcf_io.write_vcf_gz(path, HEADER, rows)
    raw = open(path, "rb").read()
    assert raw[:4] == b"\x1f\x8b\x08\x04" and raw[12:14] == b"BC"
    assert raw[-28:] == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")   # BGZF EOF block
    assert len(raw) < 0.5 * sum(len(r) + 1 for r in rows) and raw.count(b"\x1f\x8b\x08\x04") > 5
    with gzip.open(path, "rt") as fp:
        text = fp.read()
    assert text == HEADER + "".join(r + "\n" for r in rows)
    idx = vcf_io.read_tbi(path + ".tbi")
    assert (idx["format"], idx["col_seq"], idx["col_beg"], idx["col_end"], idx["meta"]) == (2, 1, 2, 0, "#")
    assert list(idx["refs"]) == sharder.contig_output_order(sort_fixture.CONTIGS)
    rng = np.random.default_rng(3)
    by_ctg = {}
    for r in rows:
        c = r.split("\t", 4)
        by_ctg.setdefault(c[0], []).append((int(c[1]), len(c[3]), r))
    n_hits = 0
    for _ in range(300):
        ctg = str(rng.choice(list(by_ctg)))
        recs = by_ctg[ctg]
        p = recs[int(rng.integers(len(recs)))][0]
        a = max(1, p + int(rng.integers(-3000, 50)))
        b = a + int(rng.choice([0, 1, 10, 700, 40000]))
        want = [r for pos, rl, r in recs if pos <= b and pos + max(1, rl) - 1 >= a]
        assert vcf_io.tabix_query(path, ctg, a, b) == want, (ctg, a, b)
        n_hits += len(want)
    assert n_hits > 500
    assert vcf_io.tabix_query(path, "chrNope", 1, 10) == [] and vcf_io.tabix_query(path, "chr2", 10 ** 8, 10 ** 8 + 5) == []


def test_unsorted_input_is_refused(tmp_path):
    rows = merged_rows()
    bad = [rows[1], rows[0]] + rows[2:]
    with pytest.raises(ValueError):
        vcf_io.write_vcf_gz(str(tmp_path / "x.vcf.gz"), HEADER, bad)
    split = [r for r in rows if r.startswith("chr1\t")][:3] + [r for r in rows if r.startswith("chr2\t")][:3] + \
            [r for r in rows if r.startswith("chr1\t")][3:6]
    with pytest.raises(ValueError):
        vcf_io.write_vcf_gz(str(tmp_path / "y.vcf.gz"), HEADER, split)
    vcf_io.write_vcf_gz(str(tmp_path / "empty.vcf.gz"), HEADER, [])
    assert gzip.open(str(tmp_path / "empty.vcf.gz"), "rt").read() == HEADER
