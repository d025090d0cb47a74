#!/usr/bin/env python3
"""Golden for the merge stage: the REFERENCE's src/sort_vcf.py:sort_vcf_from run verbatim (build container only) on
chunk VCFs made by sort_fixture.py, with and without REDIportal tagging / reference calls.  The outputs are committed
as tests/golden/sort_vcf_golden.json.  Usage: python tests/golden/make_sort_golden.py"""
import argparse
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from tests.golden import sort_fixture  # noqa: E402


def run_reference(tmp, qual, show_ref, tag, filter_tag):
    from src import sort_vcf as ref_sort
    out = os.path.join(tmp, "out_%s_%s_%s.vcf" % (qual, show_ref, tag))
    args = argparse.Namespace(
        output_fn=out, input_dir=os.path.join(tmp, "chunks"), vcf_fn_prefix="pileup", vcf_fn_suffix=".vcf",
        sample_name="SAMPLE", ref_fn=None, contigs_fn=os.path.join(tmp, "CONTIGS"), compress_vcf=False, qual=qual,
        output_no_tagging_fn=out + ".notag", show_ref=show_ref, cmd_fn=None, tag_variant_using_readiportal=tag,
        readiportal_source_fn=os.path.join(tmp, "redi.txt") if tag else None, readiportal_database_filter_tag=filter_tag)
    ref_sort.sort_vcf_from(args)
    res = dict(out=open(out).read())
    if tag:
        res["notag"] = open(out + ".notag").read()
    return res


def main():
    golden = {}
    with tempfile.TemporaryDirectory() as tmp:
        sort_fixture.write_files(tmp)
        for name, (qual, show_ref, tag, filter_tag) in sort_fixture.VARIANTS.items():
            golden[name] = run_reference(tmp, qual, show_ref, tag, filter_tag)
            print(name, len(golden[name]["out"].splitlines()), "lines")
    with open(os.path.join(HERE, "sort_vcf_golden.json"), "w") as fp:
        json.dump(golden, fp, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
