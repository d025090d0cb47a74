"""ORACLE: CPU restatement of the reference's pileup calling path.

Test infrastructure only.  Nothing under clair3_rna_b200/ imports this package;
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may.
"""
