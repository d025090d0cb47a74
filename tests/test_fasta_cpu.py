"""The native indexed-FASTA reader (csrc/fasta_io.cpp, host only) against the numpy form of fasta.fetch: regions over
line boundaries, the short last line, \\r\\n terminators, lower case, clipping at the contig ends."""
import os

import numpy as np
import pytest


@pytest.mark.parametrize("width,term", [(60, "\n"), (7, "\n"), (61, "\r\n"), (1, "\n")])
def test_fetch_into_matches_numpy(tmp_path, width, term):
    from clair3_rna_b200 import build, fasta
    build.build()
    rng = np.random.default_rng(width)
    seqs = {"chrA": "".join(rng.choice(list("ACGTacgtNn"), size=1000)), "chrB": "".join(rng.choice(list("ACGT"), size=width * 5)),
            "chrC": "G"}
    fa = os.path.join(tmp_path, "x.fa")
    off = 0
    with open(fa, "wb") as fp, open(fa + ".fai", "w") as fi:
        for name, s in seqs.items():
            hdr = (">%s%s" % (name, term)).encode()
            fp.write(hdr)
            off += len(hdr)
            fi.write("%s\t%d\t%d\t%d\t%d\n" % (name, len(s), off, width, width + len(term)))
            for i in range(0, len(s), width):
                line = (s[i:i + width] + term).encode()
                fp.write(line)
                off += len(line)
    fai = fasta.read_fai(fa)
    buf = np.empty(2000, np.uint8)
    for name, s in seqs.items():
        n = len(s)
        cases = [(1, n), (1, 1), (n, n), (-5, n + 10), (n + 1, n + 5), (3, 2)]
        cases += [tuple(sorted(int(x) for x in rng.integers(1, n + 1, 2))) for _ in range(40)]
        for a, b in cases:
            want = fasta.fetch(fa, fai, name, a, b)
            got = fasta.fetch_into(fa, fai, name, a, b, buf)
            assert got.tobytes() == want.tobytes(), (name, a, b)
            assert got.tobytes().decode() == s.upper()[max(1, a) - 1:max(0, min(n, b))]
    with pytest.raises(RuntimeError):
        fasta.fetch_into(fa, fai, "chrA", 1, 1000, np.empty(10, np.uint8))      # buffer too small
