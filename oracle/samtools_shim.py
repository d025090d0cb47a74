#!/usr/bin/env python3
"""ORACLE (test infrastructure).  Stand-in `samtools` executable.

The reference shells out to `samtools faidx` (shared/utils.py:168-194) and
`samtools mpileup` (src/create_tensor_pileup.py:436-451); samtools is not in
this image.  This shim answers exactly those two invocations so the reference's
own producer can be run verbatim (`--samtools "python oracle/samtools_shim.py"`)
to make golden vectors.  The "BAM" it reads is a flat-read .npz written by
clair3_rna_b200.reads.ReadBatch.save.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def parse_region(text):
    if ':' not in text:
        return text, None, None
    name, rng = text.rsplit(':', 1)
    a, b = rng.replace(',', '').split('-')
    return name, int(a), int(b)


def faidx(fasta, regions):
    fai = {}
    with open(fasta + ".fai") as fp:
        for row in fp:
            c = row.rstrip("\n").split("\t")
            fai[c[0]] = tuple(int(x) for x in c[1:5])
    with open(fasta, "rb") as fp:
        for region in regions:
            name, a, b = parse_region(region)
            length, offset, linebases, linewidth = fai[name]
            a = 1 if a is None else max(1, a)
            b = length if b is None else min(length, b)
            print(">%s" % region)
            out = []
            p = a - 1
            while p < b:
                line, col = divmod(p, linebases)
                take = min(b - p, linebases - col)
                fp.seek(offset + line * linewidth + col)
                out.append(fp.read(take).decode("ascii"))
                p += take
            seq = "".join(out)
            for i in range(0, len(seq), 60):
                print(seq[i:i + 60])


def mpileup(argv):
    from clair3_rna_b200.reads import ReadBatch
    from oracle.mpileup import mpileup_text
    bam, region = None, None
    min_mq, excl, extra, bed_fn = 0, 0x704, None, None
    i = 0
    while i < len(argv):
        a = argv[i]
        if a == "-r":
            region = argv[i + 1]; i += 2
        elif a == "--min-MQ":
            min_mq = int(argv[i + 1]); i += 2
        elif a == "-l":
            bed_fn = argv[i + 1]; i += 2
        elif a in ("--min-BQ", "--max-depth"):
            i += 2
        elif a == "--excl-flags":
            excl = int(argv[i + 1]); i += 2
        elif a == "--output-extra":
            extra = argv[i + 1]; i += 2
        elif a in ("--reverse-del", "-a"):
            i += 1
        elif a.startswith("-"):
            sys.exit("samtools_shim: unsupported option %s" % a)
        else:
            bam = a; i += 1
    batch = ReadBatch.load(bam)
    name, a, b = parse_region(region)
    if a is None:
        a, b = 1, int(batch.end().max()) if batch.n_reads else 1
    bed = None
    if bed_fn is not None:                      # BED rows of this contig, whitespace separated like samtools' reader
        bed = []
        with open(bed_fn) as fp:
            for row in fp:
                c = row.split()
                if len(c) >= 3 and c[0] == name and not row.startswith("#"):
                    bed.append((int(c[1]), int(c[2])))
    out = sys.stdout
    for line in mpileup_text(batch.fetch(a, b), name, a, b, excl, min_mq, with_hp=(extra == "HP"), bed=bed):
        out.write(line)
        out.write("\n")


def main():
    if len(sys.argv) < 2:
        sys.exit("usage: samtools_shim.py faidx|mpileup ...")
    if sys.argv[1] == "faidx":
        faidx(sys.argv[2], sys.argv[3:])
    elif sys.argv[1] == "mpileup":
        mpileup(sys.argv[2:])
    elif sys.argv[1] == "--version":
        print("samtools 1.17 (shim)")
    else:
        sys.exit("samtools_shim: unsupported command %s" % sys.argv[1])


if __name__ == "__main__":
    main()
