"""Constants of the pileup calling path.

Values restate /root/reference/shared/param_p.py (the second `min_coverage`
assignment at :90 wins over :22).  Only what the hot path reads is kept.
"""

VERSION = "0.2.2"            # param_p.py:3 (goes into the VCF header)

# channel order, param_p.py:31
CHANNEL = ('A', 'C', 'G', 'T', 'I', 'I1', 'D', 'D1', '*',
           'a', 'c', 'g', 't', 'i', 'i1', 'd', 'd1', '#')
CHANNEL_SIZE = len(CHANNEL)                  # 18
# create_tensor_pileup.py:181,217
PHASED_CHANNEL = ('AP', 'CP', 'GP', 'TP', 'IP', 'DP', 'AM', 'CM', 'GM', 'TM', 'IM', 'DM')
PHASED_CHANNEL_SIZE = len(PHASED_CHANNEL)    # 12, param_p.py:33
FLANK = 16                                   # param_p.py:34
NO_OF_POSITIONS = 2 * FLANK + 1              # 33
EXPAND_REFERENCE_REGION = 1000               # param_p.py:40
EXCL_FLAGS = 2316                            # param_p.py:41
SKIP_PROPORTION_THRESHOLD = 0.2              # param_p.py:46
PREDICT_BATCH_SIZE = 200                     # param_p.py:51
MAX_DEPTH = 144                              # param_p.py:14-15 (same for ont/hifi)
MIN_MQ = 5                                   # param_p.py:20
MIN_BQ = 0                                   # param_p.py:21
MIN_COVERAGE = 4                             # param_p.py:90
SNP_MIN_AF = 0.08                            # param_p.py:88
INDEL_MIN_AF = 0.15                          # param_p.py:89
CHUNK_SIZE = 5000000                         # param_p.py:91
MIN_THRED_QUAL = {'ont': 8, 'hifi': 2}       # param_p.py:85-86
QUAL_CUT_OFF = 2                             # call_variants.py --qual default
MAX_VARIANT_LENGTH = 50                      # param_p.py:16

# label layout of the 24 network outputs (clair3_rna/task/gt21.py:3-25, genotype.py:3-10)
GT21_LABELS = ('AA', 'AC', 'AG', 'AT', 'CC', 'CG', 'CT', 'GG', 'GT', 'TT',
               'DelDel', 'ADel', 'CDel', 'GDel', 'TDel',
               'InsIns', 'AIns', 'CIns', 'GIns', 'TIns', 'InsDel')
GENOTYPE_LABELS = ('0/0', '1/1', '0/1')

# platform flag -> collapsed platform (run_clair3_rna:603-607)
def collapse_platform(name: str) -> str:
    return 'ont' if name.startswith('ont') else 'hifi'

# SAM flags
FLAG_PAIRED, FLAG_UNMAP, FLAG_MUNMAP, FLAG_REVERSE = 0x1, 0x4, 0x8, 0x10
FLAG_SECONDARY, FLAG_QCFAIL, FLAG_DUP, FLAG_SUPPLEMENTARY = 0x100, 0x200, 0x400, 0x800

# BAM CIGAR op codes
CIG_M, CIG_I, CIG_D, CIG_N, CIG_S, CIG_H, CIG_P, CIG_EQ, CIG_X = range(9)
CIGAR_CHARS = "MIDNSHP=X"
