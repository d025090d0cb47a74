"""Chunk VCFs, contig list and a REDIportal table for the merge-stage golden (make_sort_golden.py) and its test.
Deterministic: the same rows are regenerated at test time."""
import os

import numpy as np

CONTIGS = ["chr2", "chr11", "chr1", "KI270728.1", "chrX"]           # CONTIGS file order (not the output order)
HEADER = ["##fileformat=VCFv4.2\n", '##FILTER=<ID=PASS,Description="All filters passed">\n',
          "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSAMPLE\n"]
# name: (qual, show_ref, tag_variant_using_readiportal, readiportal_database_filter_tag)
VARIANTS = {"default_ont": (8, False, False, None), "show_ref_hifi": (2, True, False, None),
            "tagged": (8, False, True, "A,D:A,R:A,R,D"), "tagged_all_show_ref": (2, True, True, None),
            "no_qual": (0, True, False, None)}


def chunk_rows():
    """{(contig, chunk id): [row without newline]} - positions overlap at chunk seams (a site called by two chunks
    gets the same row from both, as in the real pipeline), qualities sit around the cut-offs"""
    rng = np.random.default_rng(4242)
    out = {}
    for ci, ctg in enumerate(CONTIGS):
        n_chunks = 1 + ci % 3
        for ch in range(1, n_chunks + 1):
            rows = []
            base = 1000 * ch
            pos = sorted(set(int(x) for x in rng.integers(base - 150, base + 1100, 40)) | {base + 1000 + 7 * k for k in range(8)}
                         | {base + 7 * k for k in range(8)})
            for p in pos:
                r = np.random.default_rng([ci, p])           # the row is a function of the site
                ref = "ACGT"[int(r.integers(4))]
                kind = r.random()
                qual = float(r.choice([0.5, 1.99, 2.0, 2.01, 7.5, 8.0, 8.01, 15.25, 33.1]))
                if kind < 0.3:
                    alt, gt, flt = ".", "0/0", "RefCall"
                elif kind < 0.8:
                    alt = "ACGT".replace(ref, "")[int(r.integers(3))]
                    gt, flt = ("0/1", "PASS") if qual >= 2 else ("0/1", "LowQual")
                elif kind < 0.9:
                    alt, gt, flt = ref + "TG", "1/1", "PASS"
                else:
                    ref, alt, gt, flt = ref + "CA", ref, "0/1", "PASS"
                rows.append("%s\t%d\t.\t%s\t%s\t%.2f\t%s\t.\tGT:GQ:DP:AD:AF\t%s:%d:%d:%d,%d:%.4f" % (
                    ctg, p, ref, alt, qual, flt, gt, int(qual), 30, 20, 10, 0.3333))
            out[(ctg, ch)] = rows
    return out


def redi_rows():
    """REDIportal TABLE1-like rows: contig, pos, ref, alt, strand, db filter, ... for a third of the variant sites"""
    rng = np.random.default_rng(99)
    rows = ["Region\tPosition\tRef\tEd\tStrand\tdb\ttype\tdbsnp"]
    for (ctg, _ch), lines in sorted(chunk_rows().items()):
        for line in lines:
            c = line.split("\t")
            if rng.random() < 0.35:
                same = rng.random() < 0.7
                alt = c[4] if same else "N"
                rows.append("\t".join([ctg, c[1], c[3], alt, "+", str(rng.choice(["A", "A,D", "A,R", "A,R,D", "D"])), "ALU", "-"]))
    rows.append("chrUn\t5\tA\tG\t+\tA\tALU\t-")
    rows.append("chr1\tnot_a_number\tA\tG\t+\tA\tALU\t-")
    return rows


def write_files(tmp):
    os.makedirs(os.path.join(tmp, "chunks"), exist_ok=True)
    for (ctg, ch), rows in chunk_rows().items():
        with open(os.path.join(tmp, "chunks", "pileup_%s_%d.vcf" % (ctg, ch)), "w") as fp:
            fp.write("".join(HEADER))
            fp.write("".join(r + "\n" for r in rows))
    with open(os.path.join(tmp, "CONTIGS"), "w") as fp:
        fp.write("\n".join(CONTIGS))
    with open(os.path.join(tmp, "redi.txt"), "w") as fp:
        fp.write("\n".join(redi_rows()) + "\n")
