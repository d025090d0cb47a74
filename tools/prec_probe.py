import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from clair3_rna_b200 import weights
from clair3_rna_b200.engine import Engine
from oracle import model
for name, C in (("cfg1_ont_drna", 18), ("phased_noisy", 30), ("cfg3_hifi_pad", 18)):
    g = np.load("tests/golden/%s.npz" % name)
    x = g["tensor"][:600]
    for sharpen in (8.0, 1.0):
        w = weights.synthetic(C, sharpen=sharpen)
        eng = Engine(0, C, nn_impl=1); eng.set_weights(w)
        p, ms = eng.forward(x)
        ref = model.forward(w, x)
        print(name, "sharpen", sharpen, "n", len(x), "max|dp| %.3e  mean %.3e" % (np.abs(p-ref).max(), np.abs(p-ref).mean()), flush=True)
        eng.close()
