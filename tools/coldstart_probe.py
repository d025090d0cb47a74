"""What one cold `call_var_bam` process costs (CUDA context, weight packing, buffers) against the same chunk inside a
warm process: the reference workflow starts one such process per (contig, chunk) - 631 at config 5.
usage: python tools/coldstart_probe.py [scale]   (dataset from bench.py's cache)"""
import os, subprocess, sys, time
sys.path.insert(0, os.getcwd())
import bench
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
cores = len(os.sched_getaffinity(0))
cfg, bam, fa, wnpz, prep = bench.cfg5_dataset(scale, 0, 1, cores, lambda: None)
out = "/tmp/coldstart_chunk.vcf"
argv = ["--chkpnt_fn", wnpz, "--bam_fn", bam, "--ref_fn", fa, "--call_fn", out, "--ctgName", "chr1", "--chunk_id", "3",
        "--chunk_num", "50", "--platform", "ont", "--pileup", "--sampleName", "S"]
ts = []
for rep in range(3):
    t = time.time()
    r = subprocess.run([sys.executable, "-m", "clair3_rna_b200.call_var_bam"] + argv, capture_output=True, text=True)
    ts.append(time.time() - t)
    assert r.returncode == 0, r.stderr[-2000:]
rows = sum(1 for l in open(out) if not l.startswith("#"))
print("cold call_var_bam process, chr1 chunk 3/50 (%d rows): %s s" % (rows, ", ".join("%.2f" % x for x in ts)))
# the same chunk in a process that is already up
from clair3_rna_b200 import call_var_bam
t = time.time(); call_var_bam.main(argv); t1 = time.time() - t
t = time.time(); call_var_bam.main(argv); t2 = time.time() - t
print("same chunk through call_var_bam.main() in this process: first %.2f s (context + weights), second %.3f s" % (t1, t2))
# phases of a cold start
t = time.time(); import ctypes; from clair3_rna_b200 import lib; L = lib.load(); t_load = time.time() - t
from clair3_rna_b200.engine import Engine
from clair3_rna_b200 import weights
t = time.time(); e = Engine(0, 18); t_ctx = time.time() - t
t = time.time(); w = weights.load(wnpz); t_w = time.time() - t
t = time.time(); e.set_weights(w); t_pack = time.time() - t
print("in-process phases: dlopen %.3f s, c3r_create (second context in this process) %.3f s, weights.load %.3f s, c3r_set_weights (pack + upload) %.3f s" % (t_load, t_ctx, t_w, t_pack))
