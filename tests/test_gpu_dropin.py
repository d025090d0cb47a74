"""End to end through the drop-in `call_var_bam` entry point on the GPU: same argv as the
reference's per-chunk driver, VCF compared with rows the reference's own decoder printed for
the same candidates (tests/golden/decoder_*.npz, oracle probabilities)."""
import os

import numpy as np
import pytest

from tests.golden import cases as golden_cases

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _write_inputs(tmp, name):
    from clair3_rna_b200 import weights
    batch, ref_bytes, contig = golden_cases.build(name)
    fa = os.path.join(tmp, "ref.fa")
    with open(fa, "wb") as fp:
        fp.write(b">" + contig.encode() + b"\n")
        for i in range(0, len(ref_bytes), 60):
            fp.write(ref_bytes[i:i + 60] + b"\n")
    with open(fa + ".fai", "w") as fp:
        fp.write("%s\t%d\t%d\t60\t61\n" % (contig, len(ref_bytes), len(contig) + 2))
    batch.save(os.path.join(tmp, "reads.npz"))
    from clair3_rna_b200 import bam
    bam.write_bam(os.path.join(tmp, "reads.bam"), [(contig, len(ref_bytes))], {contig: batch})
    C = 30 if golden_cases.CASES[name]["phased"] else 18
    weights.save(os.path.join(tmp, "w.npz"), weights.synthetic(C, sharpen=8.0))
    return fa, contig


@pytest.mark.parametrize("reads", ["reads.bam", "reads.npz"])
@pytest.mark.parametrize("name", ["cfg1_ont_drna", "pad_dense", "phased_noisy"])
def test_call_var_bam_vcf(tmp_path, name, reads):
    from clair3_rna_b200 import call_var_bam
    case = golden_cases.CASES[name]
    tmp = str(tmp_path)
    fa, contig = _write_inputs(tmp, name)
    out = os.path.join(tmp, "pileup_%s_1.vcf" % contig)
    argv = ["--chkpnt_fn", os.path.join(tmp, "w.npz"), "--bam_fn", os.path.join(tmp, reads), "--ref_fn", fa,
            "--call_fn", out, "--ctgName", contig, "--chunk_id", "1", "--chunk_num", "1", "--platform", case["platform"],
            "--snp_min_af", str(case["snp_af"]), "--indel_min_af", str(case["indel_af"]), "--minMQ", str(case["min_mq"]),
            "--minCoverage", str(case["min_cov"]), "--pileup", "--sampleName", "S"]
    if case["phased"]:
        argv += ["--enable_phasing_model", "True"]
    if case["padding"]:
        argv += ["--enable_padding_in_splice_junction_regions", "True"]
    assert call_var_bam.main(argv) == 0
    lines = open(out).read().splitlines()
    assert lines[0] == "##fileformat=VCFv4.2" and lines[-1][0] != "#"
    rows = [l for l in lines if not l.startswith("#")]
    g = np.load(os.path.join(HERE, "golden", "decoder_" + name + ".npz"))
    n = len(g["rows"]) // 2
    keep = [i for i in range(n) if str(g["rows"][i])]
    want = [str(g["rows"][i]) for i in keep]
    p_ora = g["probs"][:n]                                      # the oracle network's probabilities the golden rows were made from
    assert len(rows) == len(want)
    # the probabilities the GPU produced for the same candidates (same inputs through the engine, bit-deterministic)
    from tests.test_gpu_parity import run_case
    from clair3_rna_b200 import decoder
    res, batch, ref, _ = run_case(name, 1)
    gold = np.load(os.path.join(HERE, "golden", name + ".npz"))
    assert res.pos.tolist() == gold["pos"].tolist()
    p_gpu = res.probs
    alt_info = [str(a) for a in gold["alt_info"]]
    ref33 = [str(r) for r in gold["ref33"]]
    mismatch = 0
    for a, b, i in zip(rows, want, keep):
        ca, cb = a.split("\t"), b.split("\t")
        assert ca[1] == cb[1] == str(int(gold["pos"][i]))      # same candidate positions
        fa_, fb_ = ca[9].split(":"), cb[9].split(":")          # GT:GQ:DP:AD:AF
        same = ca[:5] == cb[:5] and fa_[0] == fb_[0]
        assert fa_[2] == fb_[2]                                 # DP comes from the counts: always identical
        err = float(np.abs(p_gpu[i] - p_ora[i]).max())
        assert err <= 1e-3
        if same:
            assert fa_[3] == fb_[3] and fa_[4] == fb_[4], (a, b)                # AD and AF: integers and their ratio
            assert ca[6] == cb[6] or abs(float(ca[5]) - float(cb[5])) <= 0.2   # FILTER flips only at the QUAL cut-off
            # QUAL = f(p): d QUAL / dp = 4.34 (1/p + 1/(1-p)); compared where that slope is moderate
            best = decoder.decide(ref33[i], p_ora[i], decoder.parse_alt_info(alt_info[i])[1])[3]
            if 0.05 <= best <= 0.95:
                assert abs(float(ca[5]) - float(cb[5])) <= 0.15, (a, b)
            continue
        mismatch += 1
        # a different call is only acceptable as a near-tie: the chosen outcomes' probabilities are within twice the
        # tolerance (an outcome probability is a product of two network outputs), and the row the GPU path wrote is
        # what the reference-semantics decoder makes of the GPU's own probabilities
        b_o = decoder.decide(ref33[i], p_ora[i], decoder.parse_alt_info(alt_info[i])[1])[3]
        b_g = decoder.decide(ref33[i], p_gpu[i], decoder.parse_alt_info(alt_info[i])[1])[3]
        assert abs(float(b_o) - float(b_g)) <= 2e-3, (a, b, float(b_o), float(b_g))
        mine = decoder.vcf_row(contig, int(gold["pos"][i]), ref33[i], alt_info[i], p_gpu[i])
        assert mine == a, (mine, a)
    # calls are identical except where the two best outcomes are within the probability tolerance
    assert mismatch <= max(1, len(rows) // 200), mismatch


def test_no_record_means_no_file(tmp_path):
    from clair3_rna_b200 import call_var_bam, weights
    from clair3_rna_b200.reads import ReadBatch
    tmp = str(tmp_path)
    seq = b"ACGT" * 500
    fa = os.path.join(tmp, "ref.fa")
    with open(fa, "wb") as fp:
        fp.write(b">chr1\n" + seq + b"\n")
    with open(fa + ".fai", "w") as fp:
        fp.write("chr1\t%d\t6\t%d\t%d\n" % (len(seq), len(seq), len(seq) + 1))
    ReadBatch.from_records("chr1", []).save(os.path.join(tmp, "reads.npz"))
    weights.save(os.path.join(tmp, "w.npz"), weights.synthetic(18))
    out = os.path.join(tmp, "pileup_chr1_1.vcf")
    open(out, "w").write("stale\n")
    assert call_var_bam.main(["--chkpnt_fn", os.path.join(tmp, "w.npz"), "--bam_fn", os.path.join(tmp, "reads.npz"),
                              "--ref_fn", fa, "--call_fn", out, "--ctgName", "chr1", "--chunk_id", "1",
                              "--chunk_num", "1", "--pileup"]) == 0
    assert not os.path.exists(out)


def test_tickets_in_flight_equal_sequential_calls():
    """three tickets submitted back to back (their device work overlaps as far as the GPU allows; the network
    scratch is shared, so passes are ordered by an event) give the results of one-at-a-time calls, bit for bit"""
    import numpy as np
    from clair3_rna_b200 import weights
    from clair3_rna_b200.engine import Engine
    from tests.golden import cases as golden_cases
    names = ["cfg1_ont_drna", "ties_lowdepth", "cfg2_ont_cdna"]
    data = [golden_cases.build(n) for n in names]
    eng = Engine(0, 18, keep_tensor=True)
    eng.set_weights(weights.synthetic(18, sharpen=8.0))
    want = []
    for batch, ref_bytes, _ in data:
        ref = np.frombuffer(ref_bytes, np.uint8)
        r = eng.call_chunk(batch, ref, 1, 1, len(ref_bytes) + 33)
        want.append((r.pos.copy(), r.probs.copy(), r.tensor.copy()))
    for _round in range(3):
        tickets = [eng.submit(batch, np.frombuffer(ref_bytes, np.uint8), 1, 1, len(ref_bytes) + 33)
                   for batch, ref_bytes, _ in data]
        for t, (pos, probs, tensor) in zip(tickets, want):
            r = eng.wait(t)
            assert np.array_equal(r.pos, pos) and np.array_equal(r.tensor, tensor)
            assert np.array_equal(r.probs, probs)
    eng.close()


def test_large_batch_sub_passes_equal_small_batches():
    """a batch that needs several sub-passes of the network (more sites than the scratch holds at once) gives, site
    by site, the bits of the same sites sent in small batches"""
    import numpy as np
    from clair3_rna_b200 import weights
    from clair3_rna_b200.engine import Engine
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cfg1_ont_drna.npz"))
    base = g["tensor"]
    rng = np.random.default_rng(5)
    n = 45000
    x = base[rng.integers(0, len(base), n)]
    eng = Engine(0, 18)
    eng.set_weights(weights.synthetic(18, sharpen=8.0))
    big, _ = eng.forward(x)
    for a, b in ((0, 3000), (37000, 39500), (n - 2100, n)):
        small, _ = eng.forward(x[a:b])
        assert np.array_equal(big[a:b], small)
    eng.close()


def test_release_before_wait_and_more_tickets_than_slots():
    """c3r_submit_chunk only queues the position / row stages; the count-dependent stages are queued by later calls.
    A ticket released before anyone waited for it must leave the context usable, a fifth submit with four tickets in
    flight must be refused (not block), and tickets waited for out of order must give the right results."""
    import numpy as np
    import pytest as _pytest
    from clair3_rna_b200 import weights
    from clair3_rna_b200.engine import Engine, C3RError
    from tests.golden import cases as golden_cases
    batch, ref_bytes, _ = golden_cases.build("cfg1_ont_drna")
    ref = np.frombuffer(ref_bytes, np.uint8)
    n = len(ref_bytes)
    eng = Engine(0, 18)
    eng.set_weights(weights.synthetic(18, sharpen=8.0))
    want = eng.call_chunk(batch, ref, 1, 1, n + 33)
    t = eng.submit(batch, ref, 1, 1, n + 33)
    eng.release(t)                                               # never waited for
    ts = [eng.submit(batch, ref, 1, 1, n + 33) for _ in range(4)]
    with _pytest.raises(C3RError, match="tickets in flight"):
        eng.submit(batch, ref, 1, 1, n + 33)
    for t in (ts[2], ts[0], ts[3], ts[1]):                       # out of submit order
        r = eng.wait(t)
        assert np.array_equal(r.pos, want.pos) and np.array_equal(r.probs, want.probs) and np.array_equal(r.alt_n, want.alt_n)
    again = eng.call_chunk(batch, ref, 1, 1, n + 33)
    assert np.array_equal(again.probs, want.probs)
    eng.close()
