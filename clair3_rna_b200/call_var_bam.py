"""Drop-in for the reference's per-chunk driver `clair3_rna.py call_var_bam`
(/root/reference/clair3_rna/call_var_bam.py:88-333, options :336-518).

Same argv, same contract: exit code 0 and a VCF at --call_fn holding the header of
shared/utils.py:get_header plus one row per candidate, the file being absent when there is
no record (call_variants.py:1594-1599).  Instead of piping create_tensor_pileup text into
call_variants it hands the chunk's flat alignment records to libc3r_b200.so.

Inputs: --bam_fn is an indexed BAM (read by csrc/bam_io.cpp: BGZF inflate + BAI region fetch,
the part of `samtools mpileup -r` that touches the file) or a flat-read .npz
(clair3_rna_b200.reads.ReadBatch.save).  --chkpnt_fn is an .npz of Keras-layout weights
(clair3_rna_b200.weights) or the prefix of a Keras TF-format checkpoint, read natively by tf_bundle.py.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from . import params as P
from . import decoder, fasta, weights as W
from .reads import ReadBatch


def str2bool(v):
    if v is None or isinstance(v, bool):
        return v
    if v.lower() in ('yes', 'ture', 'true', 't', 'y', '1'):
        return True
    if v.lower() in ('no', 'flase', 'false', 'f', 'n', '0'):
        return False
    raise argparse.ArgumentTypeError('Boolean value expected.')


def build_parser():
    ap = argparse.ArgumentParser(description="Call variants of one chunk using the B200 pileup path")
    ap.add_argument('--platform', type=str, default="ont")
    ap.add_argument('--bam_fn', type=str, required=True)
    ap.add_argument('--ref_fn', type=str, required=True)
    ap.add_argument('--call_fn', type=str, required=True)
    ap.add_argument('--chkpnt_fn', type=str, required=True)
    ap.add_argument('--ctgName', type=str, required=True)
    ap.add_argument('--ctgStart', type=int, default=None)
    ap.add_argument('--ctgEnd', type=int, default=None)
    ap.add_argument('--chunk_id', type=int, default=None)
    ap.add_argument('--chunk_num', type=int, default=None)
    ap.add_argument('--sampleName', type=str, default="SAMPLE")
    ap.add_argument('--snp_min_af', type=float, default=P.SNP_MIN_AF)
    ap.add_argument('--indel_min_af', type=float, default=0.08)      # call_var_bam.py:378 (0.15 is the WORKFLOW's default, run_clair3_rna)
    ap.add_argument('--min_af', type=float, default=None)
    ap.add_argument('--minCoverage', type=int, default=P.MIN_COVERAGE)
    ap.add_argument('--minMQ', type=int, default=P.MIN_MQ)
    ap.add_argument('--qual', type=int, default=P.QUAL_CUT_OFF)
    ap.add_argument('--pileup', action='store_true')
    ap.add_argument('--enable_phasing_model', type=str2bool, default=False)
    ap.add_argument('--enable_padding_in_splice_junction_regions', type=str2bool, default=False)
    ap.add_argument('--enable_variant_calling_at_sequence_head_and_tail', type=str2bool, default=False)
    ap.add_argument('--cmd_fn', type=str, default=None)
    ap.add_argument('--device', type=int, default=0, help="CUDA device ordinal")
    ap.add_argument('--bed_fn', type=str, default=None, help="confident regions: candidates must overlap them")
    ap.add_argument('--extend_bed', type=str, default=None, help="BED of the columns to pile up (samtools mpileup -l)")
    ap.add_argument('--vcf_fn', type=str, default=None, help="genotyping mode: call exactly the sites of this VCF")
    # accepted for argv compatibility, not used by this path
    for name in ('--samtools', '--pypy', '--python', '--temp_file_dir', '--tensorflow_threads'):
        ap.add_argument(name, default=None)
    ap.add_argument('--gvcf', type=str2bool, default=False)
    ap.add_argument('--debug', action='store_true')
    return ap


def chunk_plan(args, contig_len):
    """regions.ChunkPlan of this call (geometry + site filters) as create_tensor_pileup.py:373-418, or None when the
    reference would return without output (genotyping chunk without sites)."""
    from . import regions

    def existing(path):                              # file_path_from: a missing file reads as "option not given"
        return path if path and os.path.exists(path) else None
    bed_fn, extend_bed, vcf_fn = existing(args.bed_fn), existing(args.extend_bed), existing(args.vcf_fn)
    chunked = args.chunk_id is not None and args.chunk_num is not None and args.chunk_id <= args.chunk_num
    ranged = args.ctgStart is not None and args.ctgEnd is not None and args.ctgStart <= args.ctgEnd
    return regions.plan_chunk(
        contig_len, chunk_id=args.chunk_id if chunked else None, chunk_num=args.chunk_num if chunked else None,
        ctg_start=args.ctgStart if ranged else None, ctg_end=args.ctgEnd if ranged else None,
        extend_rows=regions.read_bed_rows(extend_bed, args.ctgName) if extend_bed else None,
        confident_rows=regions.read_bed_rows(bed_fn, args.ctgName) if bed_fn else None,
        known_positions=regions.read_known_positions(vcf_fn, args.ctgName) if vcf_fn else None)


def call_chunk_to_rows(eng, batch, ref, ref_start1, start1, end1, contig, qual, native=True, site_filter=None):
    """GPU pass + decode.  native=True decodes through c3r_decode_vcf (C++, all host cores); native=False is
    the per-candidate Python decoder the native one is checked against (tests/test_decode_native_cpu.py)."""
    from .engine import alt_info_strings, flank_strings, decode_vcf_rows
    res = eng.call_chunk(batch, ref, ref_start1, start1, end1, site_filter)
    if native:
        return decode_vcf_rows(res, batch, ref, ref_start1, contig, qual=qual), res
    alts = alt_info_strings(res, batch, ref, ref_start1)
    flanks = flank_strings(res, ref, ref_start1)
    rows = []
    for i in range(res.n_cand):
        if flanks[i][P.FLANK] not in decoder.BASE2ACGT:        # clair3_rna/utils.py:113
            continue
        row = decoder.vcf_row(contig, int(res.pos[i]), flanks[i], alts[i], res.probs[i], qual_for_pass=qual)
        if row is not None:
            rows.append(row)
    return rows, res


def run(args) -> int:
    if args.gvcf:
        sys.exit("[ERROR] --gvcf is outside this path (SURVEY.md §8f)")
    from .engine import Engine
    fai = fasta.read_fai(args.ref_fn)
    if args.ctgName not in fai:
        sys.exit("[ERROR] contig %s not in %s.fai" % (args.ctgName, args.ref_fn))
    contig_len = fai[args.ctgName][0]
    try:
        plan = chunk_plan(args, contig_len)
    except ValueError as e:
        sys.exit(str(e))
    if plan is None:                                 # genotyping chunk without sites: no output file, like the reference
        if os.path.exists(args.call_fn):
            os.remove(args.call_fn)
        return 0
    s1, e1, rs1, re1 = plan.start1, plan.end1, plan.ref_start1, plan.ref_end1
    ref = fasta.fetch(args.ref_fn, fai, args.ctgName, rs1, re1)
    if args.bam_fn.endswith(".npz"):
        batch = ReadBatch.load(args.bam_fn).fetch(s1, e1)
    else:
        from .bam import BamFile
        with BamFile(args.bam_fn) as bf:
            if args.ctgName not in bf.references:
                sys.exit("[ERROR] contig %s not in the header of %s" % (args.ctgName, args.bam_fn))
            batch = bf.fetch(args.ctgName, s1, e1)
    C = P.CHANNEL_SIZE + (P.PHASED_CHANNEL_SIZE if args.enable_phasing_model else 0)
    eng = Engine(args.device, C, snp_min_af=args.snp_min_af, indel_min_af=args.indel_min_af,
                 min_coverage=args.minCoverage, min_mq=args.minMQ,
                 enable_padding=bool(args.enable_padding_in_splice_junction_regions),
                 enable_head_tail=bool(args.enable_variant_calling_at_sequence_head_and_tail))
    eng.set_weights(W.load(args.chkpnt_fn))
    rows, _ = call_chunk_to_rows(eng, batch, ref, rs1, s1, e1, args.ctgName, args.qual, site_filter=plan.site_filter())
    eng.close()
    if os.path.exists(args.call_fn):
        os.remove(args.call_fn)
    if rows:
        cmd = open(args.cmd_fn).read().rstrip() if args.cmd_fn and os.path.exists(args.cmd_fn) else None
        header = decoder.vcf_header([(k, v[0]) for k, v in fai.items()], args.sampleName,
                                    args.ref_fn if os.path.exists(args.ref_fn) else None, cmd)
        with open(args.call_fn, "w") as fp:
            fp.write(header + "\n")
            fp.write("\n".join(rows) + "\n")
    return 0


def main(argv=None):
    args = build_parser().parse_args(argv)
    return run(args)


if __name__ == "__main__":
    sys.exit(main())
