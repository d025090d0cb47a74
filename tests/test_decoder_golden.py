"""The host-side decoder (probabilities + alt_info -> VCF row) against rows printed by the
reference's own output_with (tests/golden/make_decoder_golden.py)."""
import os

import numpy as np
import pytest

from clair3_rna_b200 import decoder

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["cfg1_ont_drna", "ties_lowdepth", "pad_dense", "phased_noisy"]


def _split_qual(row):
    c = row.split("\t")
    gq = c[9].split(":")
    return c[:5] + c[6:9] + [gq[0]] + gq[2:], float(c[5]), int(gq[1])


@pytest.mark.parametrize("name", CASES)
def test_rows_match_reference(name):
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    d = np.load(os.path.join(HERE, "golden", "decoder_" + name + ".npz"))
    pos = np.concatenate([g["pos"], g["pos"]])
    ref33 = [str(s) for s in g["ref33"]] * 2
    alt = [str(s) for s in g["alt_info"]] * 2
    n_exact = 0
    for i in range(len(pos)):
        want = str(d["rows"][i])
        got = decoder.vcf_row("chr1", int(pos[i]), ref33[i], alt[i], d["probs"][i])
        if want == "":
            assert got is None, i
            continue
        assert got is not None, (i, want)
        if got == want:
            n_exact += 1
            continue
        # everything but QUAL/GQ is strict; QUAL may differ in the last digit (NumPy 1 vs 2 float
        # promotion inside the reference's quality_score_from)
        f_got, q_got, gq_got = _split_qual(got)
        f_want, q_want, gq_want = _split_qual(want)
        assert f_got == f_want, (i, got, want)
        assert abs(q_got - q_want) <= 0.011 and abs(gq_got - gq_want) <= 1, (i, got, want)
    assert n_exact >= 0.99 * len(pos)


def test_header_shape():
    h = decoder.vcf_header([("chr1", 1000)], "S1")
    assert h.splitlines()[0] == "##fileformat=VCFv4.2"
    assert h.splitlines()[-1].endswith("FORMAT\tS1")
    assert "##contig=<ID=chr1,length=1000>" in h
