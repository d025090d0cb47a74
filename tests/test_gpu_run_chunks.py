"""Whole-BAM driver (config 5 shape: many contigs, sharded by (contig, chunk)) on the GPU: a multi-contig
synthetic BAM + FASTA go through run_chunks (BAM fetch -> GPU -> native decode -> merge) and the merged VCF
must equal calling every contig in one piece; a two-rank split of the same shards must give the same file."""
import dataclasses
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_inputs(tmp, scale=0.03, n_contigs=5):
    from clair3_rna_b200 import synth, bam, weights
    cfg = synth.config(5, scale=scale)
    cfg = dataclasses.replace(cfg, contigs=cfg.contigs[:n_contigs])
    ref = synth.Reference(cfg)
    batches, fa = {}, os.path.join(tmp, "ref.fa")
    with open(fa, "wb") as fp, open(fa + ".fai", "w") as fi:
        off = 0
        for i, (name, length) in enumerate(cfg.contigs):
            batches[name] = synth.make_contig_reads(cfg, i, ref)
            head = (">%s\n" % name).encode()
            fp.write(head)
            fp.write(ref.fetch(name, 0, length).tobytes())
            fp.write(b"\n")
            fi.write("%s\t%d\t%d\t%d\t%d\n" % (name, length, off + len(head), length, length + 1))
            off += len(head) + length + 1
    bam_fn = os.path.join(tmp, "reads.bam")
    bam.write_bam(bam_fn, cfg.contigs, batches, level=1)
    w_fn = os.path.join(tmp, "w.npz")
    weights.save(w_fn, weights.synthetic(18, sharpen=8.0))
    return cfg, ref, batches, fa, bam_fn, w_fn


def test_run_chunks_equals_per_contig_calls(tmp_path):
    from clair3_rna_b200 import run_chunks, weights, params as P
    from clair3_rna_b200.engine import Engine, decode_vcf_rows
    cfg, ref, batches, fa, bam_fn, w_fn = make_inputs(str(tmp_path))
    assert max(l for _, l in cfg.contigs) > P.CHUNK_SIZE          # at least one contig spans several chunks
    out = os.path.join(str(tmp_path), "merged.vcf")
    stats = {}
    merged = run_chunks.run(bam_fn, fa, w_fn, out, stats=stats)
    assert stats["shards"] > len(cfg.contigs) and stats["candidates"] > 0
    eng = Engine(0, 18)
    eng.set_weights(weights.load(w_fn))
    want = []
    for name, length in cfg.contigs:
        r = ref.fetch(name, 0, length)
        res = eng.call_chunk(batches[name], r, 1, 1, length + P.NO_OF_POSITIONS)
        want += decode_vcf_rows(res, batches[name], r, 1, name)
    eng.close()
    assert len(merged) == len(want) and len(want) > 50

    def key(row):                                                   # probabilities are per-site bit-identical, so are the rows
        return row
    assert [key(r) for r in merged] == [key(r) for r in want]
    lines = open(out).read().splitlines()
    assert lines[0] == "##fileformat=VCFv4.2" and [l for l in lines if not l.startswith("#")] == merged

    # two ranks over the same shards (gather emulated in-process): same merged rows
    box = {}
    run_chunks.run(bam_fn, fa, w_fn, out + ".r1", rank=1, world=2, gather=lambda obj: box.setdefault("r1", obj) and None)
    two = run_chunks.run(bam_fn, fa, w_fn, out + ".2", rank=0, world=2, gather=lambda obj: [obj, box["r1"]])
    assert two == merged and len(box["r1"]) > 0
