"""Merge stage against the reference's own sort_vcf_from (golden made by tests/golden/make_sort_golden.py)."""
import json
import os

import pytest

from clair3_rna_b200 import sharder
from tests.golden import sort_fixture

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "sort_vcf_golden.json")))


def shards():
    rows = sort_fixture.chunk_rows()
    return [rows[k] for k in sorted(rows, key=lambda k: (sort_fixture.CONTIGS.index(k[0]), k[1]))]


@pytest.mark.parametrize("name", sorted(sort_fixture.VARIANTS))
def test_sort_vcf_matches_reference(name, tmp_path):
    qual, show_ref, tag, filter_tag = sort_fixture.VARIANTS[name]
    redi = None
    if tag:
        sort_fixture.write_files(str(tmp_path))
        redi = sharder.read_rediportal(os.path.join(str(tmp_path), "redi.txt"), sort_fixture.CONTIGS, filter_tag)
    rows, untagged = sharder.sort_vcf(shards(), sort_fixture.CONTIGS, qual=qual, show_ref=show_ref, rediportal=redi)
    want = [l for l in GOLDEN[name]["out"].splitlines() if not l.startswith("#")]
    assert rows == want
    assert [l for l in GOLDEN[name]["out"].splitlines() if l.startswith("#")] == [h.rstrip("\n") for h in sort_fixture.HEADER]
    if tag:
        assert untagged == [l for l in GOLDEN[name]["notag"].splitlines() if not l.startswith("#")]
        assert any("RNAEditing" in r for r in rows) and not any("RNAEditing" in r for r in untagged)
    else:
        assert untagged is None


def test_contig_order_and_gzip_table(tmp_path):
    import gzip
    assert sharder.contig_output_order(["chrM", "2", "chr10", "chr2", "GL1", "chrX", "1"]) == \
        ["chr2", "chr10", "chrX", "1", "2", "chrM", "GL1"]
    p = tmp_path / "t.txt.gz"
    with gzip.open(p, "wt") as fp:
        fp.write("Region\tPosition\tRef\tEd\tStrand\tdb\n" "chr1\t10\tA\tG\t+\tA,D\n" "chr1\t11\tA\tG\t+\tD\n" "chr9\t5\tT\tC\t-\tA\n")
    assert sharder.read_rediportal(str(p), ["chr1"], "A,D:A") == {("chr1", 10): ("A", "G", "A,D")}
    assert len(sharder.read_rediportal(str(p))) == 3
