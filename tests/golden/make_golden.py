#!/usr/bin/env python3
"""Generate golden vectors by running the REFERENCE verbatim (build container only).

For every case below this script
  1. makes synthetic reads + reference (clair3_rna_b200.synth), writes ref.fa(.fai)
     and reads.npz into a scratch directory,
  2. runs /root/reference/clair3_rna.py create_tensor_pileup unmodified, pointed
     at oracle/samtools_shim.py through its own --samtools option,
  3. feeds the producer's text rows to the reference's own
     clair3_rna.utils.tensor_generator_from (stdin mode) to obtain the int32
     batches the network would see,
  4. stores {pos, tensor, alt_info, ref33, depth} as tests/golden/<case>.npz.

/root/reference cannot travel to the GPU box, so only the .npz files (and this
script) are committed.  Usage:  python tests/golden/make_golden.py [case ...]
"""
import io
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_ROOT = "/root/reference"
sys.path.insert(0, ROOT)

from clair3_rna_b200 import synth  # noqa: E402
from clair3_rna_b200.reads import ReadBatch  # noqa: E402
from tests.golden import cases as golden_cases  # noqa: E402


def write_fasta(path, name, seq: bytes):
    with open(path, "wb") as fp:
        fp.write(b">" + name.encode() + b"\n")
        for i in range(0, len(seq), 60):
            fp.write(seq[i:i + 60] + b"\n")
    with open(path + ".fai", "w") as fp:
        fp.write("%s\t%d\t%d\t60\t61\n" % (name, len(seq), len(name) + 2))


def run_reference_producer(tmp, contig, platform, phased, padding, snp_af, indel_af, min_cov, min_mq,
                           chunk_id=1, chunk_num=1, bed_fn=None, extend_bed=None, vcf_fn=None, head_tail=False,
                           ctg_range=None):
    shim = "%s %s" % (sys.executable, os.path.join(ROOT, "oracle", "samtools_shim.py"))
    cmd = [sys.executable, os.path.join(REF_ROOT, "clair3_rna.py"), "create_tensor_pileup",
           "--bam_fn", os.path.join(tmp, "reads.npz"), "--ref_fn", os.path.join(tmp, "ref.fa"),
           "--ctgName", contig, "--platform", platform, "--samtools", shim,
           "--minCoverage", str(min_cov), "--minMQ", str(min_mq),
           "--snp_min_af", str(snp_af), "--indel_min_af", str(indel_af)]
    if ctg_range is not None:
        cmd += ["--ctgStart", str(ctg_range[0]), "--ctgEnd", str(ctg_range[1])]
    else:
        cmd += ["--chunk_id", str(chunk_id), "--chunk_num", str(chunk_num)]
    if phased:
        cmd += ["--add_phasing_feature", "True"]
    if padding:
        cmd += ["--enable_padding_in_splice_junction_regions", "True"]
    if head_tail:
        cmd += ["--enable_variant_calling_at_sequence_head_and_tail", "True"]
    if bed_fn:
        cmd += ["--bed_fn", bed_fn]
    if extend_bed:
        cmd += ["--extend_bed", extend_bed]
    if vcf_fn:
        cmd += ["--vcf_fn", vcf_fn]
    env = dict(os.environ, PYTHONPATH=REF_ROOT)
    out = subprocess.run(cmd, cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    return out.stdout.decode("ascii")


def run_reference_batcher(text, platform):
    """the reference's tensor_generator_from, verbatim, fed through sys.stdin."""
    sys.path.insert(0, REF_ROOT)
    import importlib
    import shared.param_p as param_p
    importlib.reload(param_p)                 # tensor_shape is mutated in place by the phased path
    import clair3_rna.utils as ref_utils
    importlib.reload(ref_utils)
    old = sys.stdin
    sys.stdin = io.StringIO(text)
    try:
        X, pos, alt = [], [], []
        for xb, pb, ab in ref_utils.tensor_generator_from("PIPE", 200, True, platform):
            X.append(np.array(xb))
            pos.extend(pb)
            alt.extend(ab)
    finally:
        sys.stdin = old
    X = np.concatenate(X) if X else np.zeros((0, 33, 18), np.int32)
    return X, pos, alt


def make_case(name):
    case = golden_cases.CASES[name]
    batch, ref_bytes, contig = golden_cases.build(name)
    n_chunks = case.get("chunks", 1)
    texts = []
    with tempfile.TemporaryDirectory() as tmp:
        write_fasta(os.path.join(tmp, "ref.fa"), contig, ref_bytes)
        batch.save(os.path.join(tmp, "reads.npz"))
        # region modes: the files run_clair3_rna would hand to the producer (--bed_fn as given by the user, the
        # per-contig split BED with space separated rows widened by 33 bp, the genotyping VCF)
        from clair3_rna_b200 import regions
        conf, known = golden_cases.bed_rows(name), golden_cases.known_positions(name)
        files = {}
        if conf is not None:
            files["bed_fn"] = os.path.join(tmp, "confident.bed")
            with open(files["bed_fn"], "w") as fp:
                fp.write("# confident regions\n" + "".join("%s\t%d\t%d\n" % (contig, a, b) for a, b in conf))
            ext = regions.extend_bed_rows(conf)
        if known is not None:
            files["vcf_fn"] = os.path.join(tmp, "known.vcf")
            with open(files["vcf_fn"], "w") as fp:
                fp.write("##fileformat=VCFv4.2\n" + "".join("%s\t%d\t.\tA\tC\t.\tPASS\t.\n" % (contig, p) for p in known))
            ext = regions.extend_known_rows(known)
        if files:
            files["extend_bed"] = os.path.join(tmp, contig)
            with open(files["extend_bed"], "w") as fp:
                fp.write("\n".join("%s %d %d" % (contig, a, b) for a, b in ext))
        for cid in range(1, n_chunks + 1):
            texts.append(run_reference_producer(tmp, contig, case["platform"], case["phased"], case["padding"],
                                                case["snp_af"], case["indel_af"], case["min_cov"], case["min_mq"],
                                                chunk_id=cid, chunk_num=n_chunks, head_tail=case.get("head_tail", False),
                                                ctg_range=case.get("ctg_range"), **files))
    chunk_of = np.concatenate([np.full(len(t.splitlines()), cid + 1, np.int32) for cid, t in enumerate(texts)])
    text = "".join(texts)
    rows = [r.split("\t") for r in text.splitlines()]
    raw = np.array([[int(v) for v in r[3].split()] for r in rows], dtype=np.int32)
    X, pos, alt = run_reference_batcher(text, case["platform"])
    C = 30 if case["phased"] else 18
    assert len(pos) == len(rows), (len(pos), len(rows))
    assert X.shape[1:] == (33, C), X.shape
    p = np.array([int(s.split(":")[1]) for s in pos], np.int64)
    ref33 = [s.split(":")[2] for s in pos]
    depth = np.array([int(a.split("-", 1)[0]) for a in alt], np.int64)
    out = os.path.join(HERE, name + ".npz")
    np.savez_compressed(out, pos=p, tensor=X.astype(np.int32), raw=raw.reshape(len(rows), 33, C) if len(rows) else raw,
                        alt_info=np.array(alt), ref33=np.array(ref33), depth=depth, chunk=chunk_of)
    print("%-28s candidates=%6d  reads=%6d  rescaled=%d  -> %s (%d KB)" % (
        name, len(rows), batch.n_reads, int((depth > 216).sum()), os.path.basename(out), os.path.getsize(out) // 1024))


if __name__ == "__main__":
    names = sys.argv[1:] or list(golden_cases.CASES)
    for n in names:
        make_case(n)
