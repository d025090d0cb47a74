"""host-side timeline of the pipelined public-API loop bench.py's e2e leg uses"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from bench import dataset
from clair3_rna_b200 import weights, params as P
from clair3_rna_b200.engine import Engine
from clair3_rna_b200.reads import ReadBatch
cfg, batch, ref = dataset(2, 1.0, 0)
contig, clen = cfg.contigs[0]
eng = Engine(0, 18); eng.set_weights(weights.synthetic(18, sharpen=8.0))
pin = {}
for k in ("pos", "flag", "mapq", "hp", "cigar_off", "cigar", "seq_off", "seq"):
    a = getattr(batch, k)
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).copy()).pin_memory()
    pin[k] = t.numpy().view(a.dtype)
pb = ReadBatch(batch.contig, **pin)
eng.set_reference(torch.from_numpy(ref.fetch(contig, 0, clen).copy()).pin_memory().numpy(), 1)
reg = (1, clen + P.NO_OF_POSITIONS)
for depth in (1, 2, 3):
    for _ in range(2):
        ts = [eng.submit(pb, None, 1, *reg) for _ in range(depth)]
        for t in ts: eng.wait(t)
    torch.cuda.synchronize()
    n = 12
    t0 = time.time(); sub_t = 0.0; wait_t = 0.0
    q = []
    for i in range(n):
        a = time.time(); q.append(eng.submit(pb, None, 1, *reg)); sub_t += time.time() - a
        if len(q) >= depth:
            a = time.time(); r = eng.wait(q.pop(0)); wait_t += time.time() - a
    while q:
        a = time.time(); r = eng.wait(q.pop(0)); wait_t += time.time() - a
    dt = time.time() - t0
    print("tickets in flight %d: %.3f ms/step  (submit %.3f, wait %.3f)  stage_ms %s" % (
        depth, 1e3 * dt / n, 1e3 * sub_t / n, 1e3 * wait_t / n, [round(x, 3) for x in r.stage_ms]))
