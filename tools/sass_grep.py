"""Counts of the tensor-core / bulk-copy SASS mnemonics per kernel of libc3r_b200.so (cuobjdump -sass).
usage: python tools/sass_grep.py > profiles/r2_sass_grep.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "clair3_rna_b200/libc3r_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "UTMALDG", "SYNCS", "MUFU.TANH", "MUFU.EX2", "MUFU.RCP",
        "ELECT", "UCGABAR", "MEMBAR", "ATOMG", "REDG", "SHFL"]
cur, counts = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in line:
        continue
    ins = line.split("*/", 1)[1] if "*/" in line else line
    counts[cur]["_instructions"] += 1
    for p in pats:
        if re.search(r"(?<![A-Z.])" + re.escape(p) + r"(?![A-Z])" if p != "UTCHMMA" else r"UTCHMMA(?!\.2CTA)", ins):
            counts[cur][p] += 1
print("# cuobjdump -sass %s : instruction counts per kernel (arch sm_100a)" % lib)
print("%-64s %7s " % ("kernel", "instr") + " ".join("%12s" % p for p in pats))
for k, c in counts.items():
    if not any(c[p] for p in pats):
        continue
    print("%-64s %7d " % (k[:64], c["_instructions"]) + " ".join("%12d" % c[p] for p in pats))
