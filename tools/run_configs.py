"""Every BASELINE.json config at full size (config 5: its first two contigs as whole-contig batches): device time of a
pass run alone on resident inputs, and sites/s of a STREAM of such chunks through the public API (three tickets in
flight, host buffers in, results out) - where the remainder round of one pass is filled by the next."""
import dataclasses, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from clair3_rna_b200 import synth, weights, params as P
from clair3_rna_b200.engine import Engine

def one(cfg_idx, contig_idx=0):
    cfg = synth.config(cfg_idx, scale=1.0)
    ref = synth.Reference(cfg)
    batch = synth.make_contig_reads(cfg, contig_idx, ref)
    contig, clen = cfg.contigs[contig_idx]
    C = 30 if cfg.phased else 18
    eng = Engine(0, C, enable_padding=cfg.padding)
    eng.set_weights(weights.synthetic(C, sharpen=8.0))
    r = ref.fetch(contig, 0, clen)
    eng.set_reference(r, 1)
    t = eng.submit(batch, None, 1, 1, clen + P.NO_OF_POSITIONS)
    res = eng.wait(t, release=False)
    ms = min(eng.rerun_resident(t)[0] for _ in range(5))
    st = eng.rerun_resident(t)[1]
    eng.release(t)
    # a stream of the same chunk: three tickets in flight
    from collections import deque
    q = deque(eng.submit(batch, None, 1, 1, clen + P.NO_OF_POSITIONS) for _ in range(2))
    for _ in range(3):                                   # warm
        q.append(eng.submit(batch, None, 1, 1, clen + P.NO_OF_POSITIONS)); tk = q.popleft(); eng.wait(tk, copy=False); eng.release(tk)
    n_it = 12
    t0 = time.time()
    for _ in range(n_it):
        q.append(eng.submit(batch, None, 1, 1, clen + P.NO_OF_POSITIONS)); tk = q.popleft(); eng.wait(tk, copy=False); eng.release(tk)
    dt = (time.time() - t0) / n_it
    while q:
        eng.wait(q.popleft())
    eng.close()
    print("| %d | %s %s | %d | %.1f M | %d | %d | %.3f | %.2f M | %.3f | %.2f M | %s |" % (
        cfg_idx, cfg.name, contig, batch.n_reads, batch.n_aligned_bases() / 1e6, res.n_rows, res.n_cand, ms,
        res.n_cand / ms / 1e3, 1e3 * dt, res.n_cand / dt / 1e6, " ".join("%.3f" % x for x in st[:7])), flush=True)

print("| config | workload | reads | aligned bases | rows | candidates | ms/pass alone | sites/s alone | ms/chunk in a stream | sites/s stream | stage ms alone (memset K1 K2a K2b K3 K4 K5) |")
print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---|")
for c in (1, 2, 3, 4):
    one(c)
one(5, 0)
one(5, 1)
