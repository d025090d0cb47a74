"""where the host time of Engine.wait goes (c3r_wait vs building the numpy copies), pipelined loop of depth 2"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from bench import dataset
from clair3_rna_b200 import weights, params as P, lib as L
from clair3_rna_b200 import engine as E
from clair3_rna_b200.reads import ReadBatch
cfg, batch, ref = dataset(2, 1.0, 0)
contig, clen = cfg.contigs[0]
eng = E.Engine(0, 18); eng.set_weights(weights.synthetic(18, sharpen=8.0))
pin = {}
for k in ("pos", "flag", "mapq", "hp", "cigar_off", "cigar", "seq_off", "seq"):
    a = getattr(batch, k)
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).copy()).pin_memory()
    pin[k] = t.numpy().view(a.dtype)
pb = ReadBatch(batch.contig, **pin)
eng.set_reference(torch.from_numpy(ref.fetch(contig, 0, clen).copy()).pin_memory().numpy(), 1)
reg = (1, clen + P.NO_OF_POSITIONS)
for _ in range(3):
    ts = [eng.submit(pb, None, 1, *reg) for _ in range(2)]
    for t in ts: eng.wait(t)
torch.cuda.synchronize()
n = 12
acc = dict(submit=0.0, c3r_wait=0.0, views=0.0, release=0.0)
q = [eng.submit(pb, None, 1, *reg)]
t0 = time.time()
for i in range(n):
    a = time.time(); q.append(eng.submit(pb, None, 1, *reg)); acc["submit"] += time.time() - a
    tk = q.pop(0)
    r = L.Result()
    a = time.time(); eng._check(eng.lib.c3r_wait(eng.ctx, tk, C.byref(r)), "c3r_wait"); acc["c3r_wait"] += time.time() - a
    a = time.time()
    nc = int(r.n_cand)
    alt_off = E._view(r.alt_off, nc, np.int64); alt_n = E._view(r.alt_n, nc, np.int32)
    total = int((alt_off + alt_n).max()) if nc else 0
    out = (E._view(r.pos, nc, np.int32), E._view(r.depth, nc, np.int32), E._view(r.probs, nc * 24, np.float32), E._view(r.alt, total, E.ALT_DTYPE))
    acc["views"] += time.time() - a
    a = time.time(); eng.lib.c3r_release(eng.ctx, tk); acc["release"] += time.time() - a
dt = time.time() - t0
eng.wait(q.pop(0))
print("period %.3f ms; per step: %s; alt entries %d" % (1e3 * dt / n, {k: round(1e3 * v / n, 3) for k, v in acc.items()}, total))
