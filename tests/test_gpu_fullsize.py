"""Full-size runs of the BASELINE.json configs on the GPU, checked through size-independent
properties (the Python oracle needs minutes per thousand candidates at these sizes):

  conservation   sum over rows of the strand base totals / `*`,`#` counts == the number of one-hot
                 bases under M/=/X ops / the total length of D ops of the admitted reads (host, numpy)
  depth          row_depth == base totals + deletion placeholders of the row
  windows        tensor[i][k] == count row (centre - 16 + k) for every candidate (configs without
                 padding; windows deeper than 1.5 x max_depth go through the fp64 rescale)
  run rule       the 33 rows of a window are 33 consecutive positions; candidates strictly ascending
  softmax        both heads of every probability row sum to 1
  determinism    a second pass gives bit-identical integers and probabilities
  shard          calling the contig chunk by chunk (the reference's CHUNK_LIST geometry,
                 run_clair3_rna:441-449) and merging == calling it in one piece
"""
import dataclasses

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

_REF_CONSUME = np.array([1, 0, 1, 1, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0], np.int64)
_QRY_CONSUME = np.array([1, 1, 0, 0, 1, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0], np.int64)


def make(cfg_idx):
    from clair3_rna_b200 import synth
    cfg = synth.config(cfg_idx, scale=1.0)
    cfg = dataclasses.replace(cfg, contigs=cfg.contigs[:1])
    ref = synth.Reference(cfg)
    batch = synth.make_contig_reads(cfg, 0, ref)
    contig, clen = cfg.contigs[0]
    return cfg, batch, ref.fetch(contig, 0, clen), clen


def host_totals(batch, min_mq=5, excl=2316):
    """(one-hot bases under M/=/X ops, total D length, per strand) of the reads mpileup admits"""
    flag = batch.flag.astype(np.int64)
    ok = ((flag & (excl | 0x4 | 0x100 | 0x200 | 0x400)) == 0) & (batch.mapq >= min_mq)
    ok &= ~(((flag & 1) != 0) & ((flag & 2) == 0))
    ops = (batch.cigar & 15).astype(np.int64)
    lens = (batch.cigar >> 4).astype(np.int64)
    n_ops_per = np.diff(batch.cigar_off).astype(np.int64)
    rid = np.repeat(np.arange(batch.n_reads), n_ops_per)
    qlen = lens * _QRY_CONSUME[ops]
    qcs = np.cumsum(qlen) - qlen                                  # query offset before each op, global
    first = batch.cigar_off[:-1].astype(np.int64)
    base0 = np.where(n_ops_per > 0, qcs[np.minimum(first, max(0, ops.size - 1))], 0)
    qstart = batch.seq_off[:-1][rid] + (qcs - base0[rid])         # base index into the nibble pool
    nib = np.empty(batch.seq.size * 2, np.uint8)
    nib[0::2] = batch.seq >> 4
    nib[1::2] = batch.seq & 15
    onehot = np.isin(nib, (1, 2, 4, 8))
    cs = np.concatenate([[0], np.cumsum(onehot, dtype=np.int64)])
    is_m = np.isin(ops, (0, 7, 8)) & ok[rid]
    rev = ((flag[rid] & 16) != 0)
    m_cnt = cs[(qstart + lens)[is_m]] - cs[qstart[is_m]]
    m_rev = rev[is_m]
    is_d = (ops == 2) & ok[rid]
    return (int(m_cnt[~m_rev].sum()), int(m_cnt[m_rev].sum()), int(lens[is_d & ~rev].sum()), int(lens[is_d & rev].sum()))


def ref_channel(ref, row_pos):
    lut = np.zeros(256, np.int64)
    for i, c in enumerate(b"ACGT"):
        lut[c] = i
        lut[c + 32] = i
    return lut[ref[row_pos - 1]]


@pytest.mark.parametrize("cfg_idx", [1, 2, 3, 4])
def test_full_config_properties(cfg_idx):
    from clair3_rna_b200 import weights
    from clair3_rna_b200.engine import Engine
    cfg, batch, ref, clen = make(cfg_idx)
    C = 30 if cfg.phased else 18
    eng = Engine(0, C, enable_padding=cfg.padding, keep_tensor=True, keep_rows=True)
    eng.set_weights(weights.synthetic(C, sharpen=8.0))
    res = eng.call_chunk(batch, ref, 1, 1, clen + 33)
    res2 = eng.call_chunk(batch, ref, 1, 1, clen + 33)
    eng.close()
    assert res.n_cand > 0 and res.n_rows > 0
    # determinism
    assert np.array_equal(res.pos, res2.pos) and np.array_equal(res.tensor, res2.tensor)
    assert np.array_equal(res.row_counts, res2.row_counts) and np.array_equal(res.probs, res2.probs)
    assert np.array_equal(res.alt["count"], res2.alt["count"]) and np.array_equal(res.alt["order"], res2.alt["order"])
    # conservation and depth
    rows = res.row_counts.reshape(-1, C).astype(np.int64)
    ri = ref_channel(ref, res.row_pos.astype(np.int64))
    idx = np.arange(rows.shape[0])
    fsum, rsum = -rows[idx, ri], -rows[idx, 9 + ri]
    assert (fsum >= 0).all() and (rsum >= 0).all()
    mf, mr, df, dr = host_totals(batch)
    assert (int(fsum.sum()), int(rsum.sum()), int(rows[:, 8].sum()), int(rows[:, 17].sum())) == (mf, mr, df, dr)
    others_f = rows[:, 0:4].sum(1) - rows[idx, ri]
    others_r = rows[:, 9:13].sum(1) - rows[idx, 9 + ri]
    assert (others_f <= fsum).all() and (others_r <= rsum).all()
    assert np.array_equal(res.row_depth.astype(np.int64), fsum + rsum + rows[:, 8] + rows[:, 17])
    if C == 30:
        # phased base counts never exceed the strand totals they are drawn from
        assert (rows[:, 18:22].sum(1) + rows[:, 24:28].sum(1) <= fsum + rsum).all()
    # run rule and ordering
    assert (np.diff(res.pos) > 0).all()
    row_of = np.searchsorted(res.row_pos, res.pos)
    assert np.array_equal(res.row_pos[row_of], res.pos)
    assert (row_of >= 16).all() and (row_of + 16 < res.n_rows).all()
    assert np.array_equal(res.row_pos[row_of + 16] - res.row_pos[row_of - 16], np.full(res.n_cand, 32))
    # windows (K4 gather, K5 input) for the configs whose rows are not patched by the padding rule
    if not cfg.padding:
        win = rows[(row_of[:, None] + np.arange(-16, 17)[None, :])]           # [n,33,C]
        deep = res.depth.astype(np.float64) > 144 * 1.5
        assert np.array_equal(res.tensor[~deep].astype(np.int64), win[~deep])
        if deep.any():
            sf = res.depth[deep].astype(np.float64) / 144.0
            want = (win[deep].astype(np.float64) / sf[:, None, None]).astype(np.int32)    # clair3_rna/utils.py:85-92,120
            assert np.array_equal(res.tensor[deep], want)
    else:
        assert np.abs(res.tensor).max() <= max(217, int(np.abs(rows).max()))
    # softmax heads
    assert np.isfinite(res.probs).all() and (res.probs >= 0).all()
    assert np.abs(res.probs[:, :21].sum(1) - 1).max() < 1e-4 and np.abs(res.probs[:, 21:].sum(1) - 1).max() < 1e-4
    # alt_info depth = the centre row's depth
    assert np.array_equal(res.depth, res.row_depth[row_of])


def test_chunked_calls_equal_one_piece():
    """config 2 (chr20-sized contig, 13 reference chunks): per-chunk calls merged like sort_vcf == one call"""
    from clair3_rna_b200 import weights, params as P
    from clair3_rna_b200.engine import Engine
    from clair3_rna_b200.synth import chunk_geometry
    cfg, batch, ref, clen = make(2)
    eng = Engine(0, 18, keep_tensor=True)
    eng.set_weights(weights.synthetic(18, sharpen=8.0))
    whole = eng.call_chunk(batch, ref, 1, 1, clen + P.NO_OF_POSITIONS)
    chunk_num = -(-clen // 5000000)
    assert chunk_num == 13
    merged = {}
    for cid in range(1, chunk_num + 1):
        _, _, s, e, rs, re_ = chunk_geometry(clen, cid, chunk_num)
        sub = batch.fetch(s, e)
        r = eng.call_chunk(sub, ref[rs - 1:re_], rs, s, e)
        for i, p in enumerate(r.pos.tolist()):
            item = (r.tensor[i].tobytes(), r.probs[i].tobytes(), int(r.depth[i]))
            if p in merged:
                assert merged[p][0] == item[0] and merged[p][2] == item[2]      # boundary duplicates are identical
            merged[p] = item
    eng.close()
    assert sorted(merged) == whole.pos.tolist()
    for i, p in enumerate(whole.pos.tolist()):
        assert merged[p][0] == whole.tensor[i].tobytes()
        assert np.abs(np.frombuffer(merged[p][1], np.float32) - whole.probs[i]).max() <= 1e-6


def test_three_tickets_in_flight_at_full_size():
    """config 2 at full size (58 tile pairs: a full round + a remainder round of the recurrent kernels) with three
    tickets in flight: LSTM1 of a pass is then queued in pieces that run beside the previous pass's remainder round
    on the second h1 buffer (tc_forward), and the candidate-count dependent stages of a ticket are queued by later
    calls (advance).  Every ticket must give the result of a call made alone, bit for bit."""
    from clair3_rna_b200 import weights, params as P
    from clair3_rna_b200.engine import Engine
    cfg, batch, ref, clen = make(2)
    eng = Engine(0, 18)
    eng.set_weights(weights.synthetic(18, sharpen=8.0))
    eng.set_reference(ref, 1)
    region = (1, clen + P.NO_OF_POSITIONS)
    alone = eng.call_chunk(batch, None, 1, *region)
    assert alone.n_cand > 2 * 37 * 128                      # more than one round of tile pairs
    from collections import deque
    q = deque(eng.submit(batch, None, 1, *region) for _ in range(2))
    for _ in range(6):
        q.append(eng.submit(batch, None, 1, *region))
        r = eng.wait(q.popleft())
        assert np.array_equal(r.pos, alone.pos) and np.array_equal(r.depth, alone.depth)
        assert np.array_equal(r.probs, alone.probs)
        assert np.array_equal(r.alt_n, alone.alt_n)
    while q:
        r = eng.wait(q.popleft())
        assert np.array_equal(r.probs, alone.probs)
    eng.close()
