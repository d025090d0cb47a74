"""Region modes on the GPU: --extend_bed / --bed_fn / --vcf_fn site filters through
c3r_submit_chunk_filtered.  The reference-made goldens (bed_regions, bed_pad_hifi, known_sites) are checked in
tests/test_gpu_parity.py; here: other channel layouts against the oracle on fresh random filters, identities,
argument checks and the two drivers end to end."""
import os

import numpy as np
import pytest

from tests.golden import cases as golden_cases

pytestmark = pytest.mark.gpu


def _engine(case, **kw):
    from clair3_rna_b200 import weights
    from clair3_rna_b200.engine import Engine
    C = 30 if case["phased"] else 18
    eng = Engine(0, C, snp_min_af=case["snp_af"], indel_min_af=case["indel_af"], min_coverage=case["min_cov"],
                 min_mq=case["min_mq"], enable_padding=case["padding"], nn_impl=0, keep_tensor=True,
                 enable_head_tail=case.get("head_tail", False), **kw)
    eng.set_weights(weights.synthetic(C, sharpen=8.0))
    return eng


def _random_rows(batch, n_ref, rng, n):
    rows = []
    for _ in range(n):
        a = max(0, int(batch.pos[int(rng.integers(batch.n_reads))]) + int(rng.integers(-50, 600)))
        rows.append((a, min(n_ref, a + int(rng.choice([0, 1, 3, 35, 120, 400])))))
    return rows


@pytest.mark.parametrize("name,seed", [("phased_noisy", 1), ("pad_dense", 2), ("ties_lowdepth", 3),
                                       ("headtail_ont", 4), ("headtail_pad", 5), ("headtail_phased_bed", 6)])
def test_random_filters_match_oracle(name, seed):
    from oracle import pileup_oracle
    from clair3_rna_b200 import regions
    from clair3_rna_b200.engine import alt_info_strings
    case = golden_cases.CASES[name]
    batch, ref_bytes, _ = golden_cases.build(name)
    ref = np.frombuffer(ref_bytes, np.uint8)
    ref_seq = ref_bytes.decode("ascii")
    n = len(ref_seq)
    rng = np.random.default_rng(seed)
    conf = _random_rows(batch, n, rng, 70)
    ext = regions.extend_bed_rows(conf)
    # with padding the reference divides by the largest depth of the window and raises ZeroDivisionError for a known
    # site whose 33 columns all lie inside an intron (known limits): keep those sites next to read starts
    span = 25 if case["padding"] else 300
    known = sorted({int(batch.pos[int(rng.integers(batch.n_reads))]) + 1 + int(rng.integers(0, span)) for _ in range(150)})
    known = [p for p in known if p <= n]
    kw = dict(snp_min_af=case["snp_af"], indel_min_af=case["indel_af"], min_coverage=case["min_cov"],
              min_mq=case["min_mq"], padding=case["padding"], phased=case["phased"],
              head_tail=case.get("head_tail", False))
    eng = _engine(case)
    combos = [dict(pileup_bed=ext, confident_bed=conf), dict(pileup_bed=ext), dict(confident_bed=conf),
              dict(pileup_bed=regions.extend_known_rows(known), known=known), dict(known=known),
              dict(pileup_bed=ext, confident_bed=conf, known=known)]
    checked = 0
    for combo in combos:
        want = pileup_oracle.run_region(batch, ref_seq, 1, 1, n + 33, **combo, **kw)
        flt = dict(pileup_bed=regions.merge_intervals(combo["pileup_bed"]) if "pileup_bed" in combo else None,
                   confident=regions.confident_intervals(combo["confident_bed"], 1, n + 33) if "confident_bed" in combo else None,
                   known=np.asarray(combo["known"], np.int32) if "known" in combo else None)
        res = eng.call_chunk(batch, ref, 1, 1, n + 33, flt)
        assert res.pos.tolist() == want["pos"].tolist(), sorted(combo)
        assert res.depth.tolist() == want["depth"].tolist()
        assert np.array_equal(res.tensor, want["tensor"])
        assert alt_info_strings(res, batch, ref, 1) == list(want["alt_info"])
        checked += len(want["pos"])
    eng.close()
    assert checked > 100


def test_filter_identities_and_argument_checks():
    from clair3_rna_b200.engine import C3RError
    name = "cfg1_ont_drna"
    case = golden_cases.CASES[name]
    batch, ref_bytes, _ = golden_cases.build(name)
    ref = np.frombuffer(ref_bytes, np.uint8)
    n = len(ref_bytes)
    eng = _engine(case)
    plain = eng.call_chunk(batch, ref, 1, 1, n + 33)
    everything = np.array([[0, n + 64]], np.int32)
    same = eng.call_chunk(batch, ref, 1, 1, n + 33, dict(pileup_bed=everything, confident=everything))
    assert same.pos.tolist() == plain.pos.tolist() and np.array_equal(same.tensor, plain.tensor)
    assert np.array_equal(same.probs, plain.probs)
    # the candidates of the plain run as known sites: the same set, tensors and probabilities
    again = eng.call_chunk(batch, ref, 1, 1, n + 33, dict(known=plain.pos.astype(np.int32)))
    assert again.pos.tolist() == plain.pos.tolist() and np.array_equal(again.tensor, plain.tensor)
    # given but empty: nothing is piled up / nothing is confident / no site is known
    for flt in (dict(pileup_bed=np.zeros((0, 2), np.int32)), dict(confident=np.zeros((0, 2), np.int32)),
                dict(known=np.zeros(0, np.int32))):
        assert eng.call_chunk(batch, ref, 1, 1, n + 33, flt).n_cand == 0
    for bad in (dict(pileup_bed=np.array([[50, 60], [10, 20]], np.int32)),          # unsorted
                dict(confident=np.array([[10, 20], [20, 30]], np.int32)),            # touching
                dict(confident=np.array([[10, 10]], np.int32)),                      # empty interval
                dict(known=np.array([7, 7], np.int32))):                             # repeated site
        with pytest.raises(C3RError):
            eng.call_chunk(batch, ref, 1, 1, n + 33, bad)
    assert eng.call_chunk(batch, ref, 1, 1, n + 33).pos.tolist() == plain.pos.tolist()   # the context is still usable
    eng.close()


def _write_case_files(tmp, name):
    """FASTA, BAM, weights, confident BED (+ its split BED) and genotyping VCF of a golden case"""
    from clair3_rna_b200 import bam, regions, weights
    batch, ref_bytes, contig = golden_cases.build(name)
    fa = os.path.join(tmp, "ref.fa")
    with open(fa, "wb") as fp:
        head = (">%s\n" % contig).encode()
        fp.write(head + ref_bytes + b"\n")
    with open(fa + ".fai", "w") as fp:
        fp.write("%s\t%d\t%d\t%d\t%d\n" % (contig, len(ref_bytes), len(head), len(ref_bytes), len(ref_bytes) + 1))
    bam_fn = os.path.join(tmp, "reads.bam")
    bam.write_bam(bam_fn, [(contig, len(ref_bytes))], {contig: batch}, level=1)
    w_fn = os.path.join(tmp, "w.npz")
    C = 30 if golden_cases.CASES[name]["phased"] else 18
    weights.save(w_fn, weights.synthetic(C, sharpen=8.0))
    files = dict(fa=fa, bam=bam_fn, w=w_fn, contig=contig, n=len(ref_bytes))
    conf, known = golden_cases.bed_rows(name), golden_cases.known_positions(name)
    if conf is not None:
        files["bed"] = os.path.join(tmp, "confident.bed")
        with open(files["bed"], "w") as fp:
            fp.write("".join("%s\t%d\t%d\n" % (contig, a, b) for a, b in conf))
        ext = regions.extend_bed_rows(conf)
    if known is not None:
        files["vcf"] = os.path.join(tmp, "known.vcf")
        with open(files["vcf"], "w") as fp:
            fp.write("##fileformat=VCFv4.2\n" + "".join("%s\t%d\t.\tA\tC\t.\tPASS\t.\n" % (contig, p) for p in known))
        ext = regions.extend_known_rows(known)
    files["split"] = os.path.join(tmp, "split_" + contig)
    with open(files["split"], "w") as fp:
        fp.write("\n".join("%s %d %d" % (contig, a, b) for a, b in ext))
    return files


@pytest.mark.parametrize("name", ["bed_regions", "known_sites"])
def test_drivers_call_the_golden_sites(tmp_path, name):
    """call_var_bam per chunk (the reference's argv) and run_chunks (whole BAM) in region mode: the called
    positions are the golden candidate positions the reference's producer emitted for the same files."""
    from clair3_rna_b200 import call_var_bam, run_chunks, weights
    from clair3_rna_b200.engine import Engine, decode_vcf_rows
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    f = _write_case_files(str(tmp_path), name)
    n_chunks = golden_cases.CASES[name]["chunks"]
    # rows of the same plans through the engine directly
    batch, ref_bytes, contig = golden_cases.build(name)
    ref = np.frombuffer(ref_bytes, np.uint8)
    eng = Engine(0, 18)
    eng.set_weights(weights.load(f["w"]))
    want = []
    for plan in golden_cases.chunk_plans(name):
        if plan is None:
            continue
        sub, r = batch.fetch(plan.start1, plan.end1), ref[plan.ref_start1 - 1:plan.ref_end1]
        res = eng.call_chunk(sub, r, plan.ref_start1, plan.start1, plan.end1, plan.site_filter())
        want += decode_vcf_rows(res, sub, r, plan.ref_start1, contig)
    eng.close()
    want_pos = [int(r.split("\t")[1]) for r in want]
    assert set(want_pos) <= set(g["pos"].tolist()) and len(want_pos) >= 0.9 * len(g["pos"])
    got = []
    for cid in range(1, n_chunks + 1):
        out = os.path.join(str(tmp_path), "c%d.vcf" % cid)
        argv = ["--bam_fn", f["bam"], "--ref_fn", f["fa"], "--chkpnt_fn", f["w"], "--ctgName", f["contig"],
                "--chunk_id", str(cid), "--chunk_num", str(n_chunks), "--call_fn", out, "--pileup",
                "--snp_min_af", "0.08", "--indel_min_af", "0.15",          # what run_clair3_rna passes (its own defaults)
                "--extend_bed", f["split"]]
        argv += ["--bed_fn", f["bed"]] if "bed" in f else ["--vcf_fn", f["vcf"]]
        assert call_var_bam.main(argv) == 0
        if os.path.exists(out):
            got += [l.rstrip("\n") for l in open(out) if not l.startswith("#")]
    assert got == want
    # whole-BAM driver: one contig shorter than a chunk -> one shard; in BED mode its geometry is the BED span with
    # chunk_num 1, so the candidate set can only grow at the former chunk seam; in VCF mode it is the same set
    out = os.path.join(str(tmp_path), "all.vcf")
    kw = dict(bed_fn=f["bed"]) if "bed" in f else dict(vcf_fn=f["vcf"])
    rows = run_chunks.run(f["bam"], f["fa"], f["w"], out, **kw)
    pos = [int(r.split("\t")[1]) for r in rows]
    assert pos == sorted(pos) and set(want_pos) <= set(pos)
    if "vcf" in f:
        assert rows == want
