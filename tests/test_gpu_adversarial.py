"""Randomised small inputs the synthetic workloads never produce, against the oracle on every count row and on
the candidate set: N / IUPAC codes / '=' in read bases, N and IUPAC codes in the reference, `=`/`X` ops, soft and
hard clips, insertions next to clips, ops cut by the region bounds, reads overhanging the loaded reference
window on both sides, every filter flag, deep and shallow spots, HP tags of every kind.

One reference quirk is deliberately not generated in phased mode: a read base that is an IUPAC code other than N
(or '=') prints a letter the reference's column parser neither counts nor consumes an HP value for
(create_tensor_pileup.py:113-145), which shifts the HP values of all later reads of that column by one.  The
device keeps HP values attached to their reads (DESIGN.md "Known limits")."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

M, I, D, N, S, H, P_, EQ, X = range(9)


def random_case(seed, phased, clean_ref=False):
    from clair3_rna_b200.reads import ReadBatch
    rng = np.random.default_rng(seed)
    L = 3000
    ref = rng.choice(np.frombuffer(b"ACGT", np.uint8), L)
    for _ in range(0 if clean_ref else 6):               # runs of N and single IUPAC codes
        a = int(rng.integers(0, L - 40))
        ref[a:a + int(rng.integers(1, 25))] = ord('N')
    for _ in range(0 if clean_ref else 20):
        ref[int(rng.integers(0, L))] = ord(rng.choice(list("RYKMSW")))
    nt16_of = {65: 1, 67: 2, 71: 4, 84: 8}
    recs = []
    hot = [int(rng.integers(400, 2600)) for _ in range(3)]    # a few deep spots
    for _ in range(int(rng.integers(150, 260))):
        pos = int(rng.integers(0, L - 50)) if rng.random() < 0.6 else max(0, int(rng.choice(hot)) - int(rng.integers(0, 120)))
        cig, codes, x = [], [], pos
        if rng.random() < 0.15: cig.append((int(rng.integers(1, 9)), H))
        if rng.random() < 0.3:
            n = int(rng.integers(1, 12)); cig.append((n, S)); codes += list(rng.choice([1, 2, 4, 8], n))
        if rng.random() < 0.1:
            n = int(rng.integers(1, 4)); cig.append((n, I)); codes += list(rng.choice([1, 2, 4, 8], n))
        n_blocks = int(rng.integers(1, 7))
        for b in range(n_blocks):
            n = int(rng.integers(1, 90))
            n = min(n, L + 200 - x)
            if n <= 0:
                break
            op = M if rng.random() < 0.8 else (EQ if rng.random() < 0.5 else X)
            cig.append((n, op))
            for k in range(n):
                rb = int(ref[x + k]) if x + k < L else 65
                c = nt16_of.get(rb, 1)
                r = rng.random()
                if r < 0.06: c = int(rng.choice([1, 2, 4, 8]))
                elif r < 0.08: c = 15
                elif r < 0.09 and not phased: c = int(rng.choice([3, 5, 6, 10, 0]))     # see the note in the docstring
                codes.append(c)
            x += n
            if b + 1 < n_blocks:
                r = rng.random()
                if r < 0.35:
                    n2 = int(rng.integers(1, 6)); cig.append((n2, I)); codes += list(rng.choice([1, 2, 4, 8, 15], n2, p=[.24, .24, .24, .24, .04]))
                elif r < 0.7:
                    n2 = int(rng.integers(1, 8)); cig.append((n2, D)); x += n2
                else:
                    n2 = int(rng.integers(5, 400)); cig.append((n2, N)); x += n2
        if cig[-1][1] not in (M, EQ, X):
            continue
        if rng.random() < 0.3:
            n = int(rng.integers(1, 12)); cig.append((n, S)); codes += list(rng.choice([1, 2, 4, 8], n))
        flag = int(rng.choice([0, 16, 0, 16, 0, 16, 4, 256, 2048, 1024, 512, 1, 3, 16 | 256]))
        mapq = int(rng.choice([60, 60, 60, 20, 5, 4, 0]))
        hp = int(rng.choice([0, 1, 2, 1, 2, 3, 255])) if phased else 0
        recs.append((pos, flag, mapq, hp, cig, np.array(codes, np.uint8)))
    recs.sort(key=lambda r: r[0])
    return ReadBatch.from_records("c", recs), ref


@pytest.mark.parametrize("seed,phased,padding", [(1, False, False), (2, True, False), (3, False, True), (4, True, True),
                                                  (5, False, False), (6, True, False)])
def test_random_adversarial_inputs_match_oracle(seed, phased, padding):
    from clair3_rna_b200 import weights
    from clair3_rna_b200.engine import Engine, alt_info_strings
    from oracle import mpileup, pileup_oracle
    # with the padding rule on, the reference itself raises KeyError when a padded window position has a
    # non-ACGT reference base (BASE2INDEX lookup, create_tensor_pileup.py:591); the device skips such positions
    batch, ref_full = random_case(seed, phased, clean_ref=padding)
    ref_start1, ref_end1 = 151, 2850                      # the loaded window: reads overhang it on both sides
    ref = ref_full[ref_start1 - 1:ref_end1]
    s1, e1 = 300, 2700
    C = 30 if phased else 18
    snp_af, indel_af, min_cov = (0.08, 0.15, 2) if seed % 2 else (0.3, 0.4, 1)
    eng = Engine(0, C, snp_min_af=snp_af, indel_min_af=indel_af, min_coverage=min_cov, min_mq=5, enable_padding=padding,
                 nn_impl=0, keep_tensor=True, keep_rows=True)
    eng.set_weights(weights.synthetic(C))
    sub = batch.fetch(s1, e1)
    res = eng.call_chunk(sub, ref, ref_start1, s1, e1)
    eng.close()
    ref_seq = ref.tobytes().decode("ascii")
    rows = {int(p): i for i, p in enumerate(res.row_pos)}
    n_checked = 0
    for pos1, depth_col, bases, hps in mpileup.mpileup_rows(sub, s1, e1, 2316, 5):
        vec, _alt, depth, _pass, _ms = pileup_oracle.column_vector(
            pos1, bases, ref_seq, ref_start1, hps.split(",") if phased else None, snp_af, indel_af)
        if pos1 in rows:
            i = rows[pos1]
            assert res.row_counts[i].tolist() == vec, (seed, pos1)
            assert int(res.row_depth[i]) == depth
            n_checked += 1
        else:
            assert depth == 0 and not any(vec), "covered column %d has no row" % pos1
    assert n_checked == len(rows) and n_checked > 500
    ora = pileup_oracle.run_region(sub, ref_seq, ref_start1, s1, e1, snp_min_af=snp_af, indel_min_af=indel_af,
                                   min_coverage=min_cov, min_mq=5, padding=padding, phased=phased)
    assert res.pos.tolist() == ora["pos"].tolist()
    assert np.array_equal(res.tensor, ora["tensor"])
    assert alt_info_strings(res, sub, ref, ref_start1) == ora["alt_info"]


def test_all_n_reference_takes_the_capacity_retry():
    """Against a reference of N every read base is a row event, far beyond the first-attempt event capacity
    (ops + bases/4): the device flags the overflow and c3r_submit_chunk retries with the exact bound.  Rows
    must still equal the oracle's columns (base counts by strand, no reference channel to fold them into)."""
    from clair3_rna_b200 import weights
    from clair3_rna_b200.engine import Engine
    from oracle import mpileup, pileup_oracle
    batch, ref_full = random_case(11, False)
    ref = np.full_like(ref_full, ord('N'))
    ref[1000:1400] = ref_full[1000:1400]                 # a stretch of real bases in the middle
    eng = Engine(0, 18, min_coverage=2, nn_impl=0, keep_rows=True, keep_tensor=True)
    eng.set_weights(weights.synthetic(18))
    s1, e1 = 1, 3000
    res = eng.call_chunk(batch, ref, 1, s1, e1)
    eng.close()
    ref_seq = ref.tobytes().decode("ascii")
    rows = {int(p): i for i, p in enumerate(res.row_pos)}
    n = 0
    for pos1, _d, bases, _hps in mpileup.mpileup_rows(batch, s1, e1, 2316, 5):
        vec, _alt, depth, _pass, _ms = pileup_oracle.column_vector(pos1, bases, ref_seq, 1, None, 0.08, 0.15)
        if pos1 in rows:
            assert res.row_counts[rows[pos1]].tolist() == vec, pos1
            n += 1
    assert n == len(rows) and n > 1500
    ora = pileup_oracle.run_region(batch, ref_seq, 1, s1, e1, min_coverage=2)
    assert res.pos.tolist() == ora["pos"].tolist() and np.array_equal(res.tensor, ora["tensor"])
    assert all(1001 <= p <= 1400 for p in res.pos.tolist())


def test_depth_above_the_mpileup_cap_is_refused():
    """samtools mpileup reads at most 8000 reads per position (-d default; the reference never passes --max-depth,
    create_tensor_pileup.py:442) and drops the rest in read order.  That is not reproduced: the chunk is refused
    (C3R_ERR_CAPACITY with a message that says why) instead of being called on counts the reference never saw."""
    import numpy as np
    import pytest as _pytest
    from clair3_rna_b200 import weights
    from clair3_rna_b200.engine import Engine, C3RError
    from clair3_rna_b200.reads import ReadBatch, encode_seq
    rng = np.random.default_rng(3)
    ref = rng.choice(np.frombuffer(b"ACGT", np.uint8), 400)
    seq = ref[100:200].tobytes().decode()
    def batch(n):
        return ReadBatch.from_records("c", [(100, 16 * (i & 1), 60, 0, [(100, 0)], encode_seq(seq)) for i in range(n)])
    eng = Engine(0, 18)
    eng.set_weights(weights.synthetic(18))
    ok = eng.call_chunk(batch(8000), ref, 1, 1, 433)            # exactly at the cap: called
    assert ok.n_rows == 100
    with _pytest.raises(C3RError, match="8000 reads"):
        eng.call_chunk(batch(8001), ref, 1, 1, 433)
    again = eng.call_chunk(batch(50), ref, 1, 1, 433)           # the context keeps working after the refusal
    assert again.n_rows == 100
    eng.close()
