"""The oracle restatement must reproduce the golden vectors made by the
reference's own code (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import pileup_oracle
from tests.golden import cases as golden_cases

HERE = os.path.dirname(os.path.abspath(__file__))


def load_golden(name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    return {k: z[k] for k in z.files}


def oracle_run(name):
    case = golden_cases.CASES[name]
    batch, ref_bytes, contig = golden_cases.build(name)
    ref_seq = ref_bytes.decode("ascii")
    n = len(ref_seq)
    kw = dict(snp_min_af=case["snp_af"], indel_min_af=case["indel_af"], min_coverage=case["min_cov"],
              min_mq=case["min_mq"], padding=case["padding"], phased=case["phased"],
              head_tail=case.get("head_tail", False))
    # producer call by producer call with the reference's geometry (create_tensor_pileup.py:373-418), results
    # concatenated; region-mode cases carry their BED / known-site filters in the plan
    parts = []
    conf = golden_cases.bed_rows(name)
    for plan in golden_cases.chunk_plans(name):
        if plan is None:
            continue
        s1, e1, rs1, re1 = plan.start1, plan.end1, plan.ref_start1, plan.ref_end1
        parts.append(pileup_oracle.run_region(
            batch.fetch(s1, e1), ref_seq[rs1 - 1:re1], rs1, s1, e1,
            pileup_bed=None if plan.pileup_bed is None else plan.pileup_bed.tolist(),
            confident_bed=conf, known=None if plan.known is None else plan.known.tolist(), **kw))
    return dict(pos=np.concatenate([p["pos"] for p in parts]), depth=np.concatenate([p["depth"] for p in parts]),
                tensor=np.concatenate([p["tensor"] for p in parts]), alt_info=sum((list(p["alt_info"]) for p in parts), []),
                ref33=sum((list(p["ref33"]) for p in parts), []))


@pytest.mark.parametrize("name", sorted(golden_cases.CASES))
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    o = oracle_run(name)
    assert o["pos"].tolist() == g["pos"].tolist()
    assert o["depth"].tolist() == g["depth"].tolist()
    assert list(o["alt_info"]) == [str(s).rstrip("\n") for s in g["alt_info"]]
    assert list(o["ref33"]) == [str(s) for s in g["ref33"]]
    assert np.array_equal(o["tensor"], g["tensor"])
