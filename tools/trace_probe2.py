import os, sys, ctypes as C
os.environ["C3R_TRACE"] = "1"
sys.path.insert(0, os.getcwd())
import numpy as np
from clair3_rna_b200 import weights
from clair3_rna_b200.engine import Engine
g = np.load("tests/golden/cfg1_ont_drna.npz")
x = np.concatenate([g["tensor"]] * 40)[:29696]
eng = Engine(0, 18); eng.set_weights(weights.synthetic(18, sharpen=8.0))
for _ in range(3): p, ms = eng.forward(x)
n = 2 * 2 * 33 * 8 * 8 + 64 * 32
buf = np.zeros(n, np.int64); nb = C.c_int64(0)
eng.lib.c3r_debug_fetch(eng.ctx, 4, buf.ctypes.data, buf.nbytes, C.byref(nb))
T = buf[2 * 2 * 33 * 8 * 8:].reshape(64, 32)
t0 = T[5, 16]
for tc in range(5, 13):
    r = T[tc]
    print("tile %2d | acce_ok %6d | stage ready %s | commit %6d | epi %6d..%6d" % (
        tc, r[16] - t0, [int(v - t0) for v in r[8:12]], r[17] - t0, r[18] - t0, r[19] - t0))
