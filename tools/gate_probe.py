"""max |dp| of the tensor-core network against the fp32 oracle on both weight sets (C3R_LIB selects the build)."""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from clair3_rna_b200 import weights
from clair3_rna_b200.engine import Engine
from oracle import model
g = np.load("tests/golden/cfg1_ont_drna.npz")["tensor"]
x = np.concatenate([g, (g[::-1] * 2).astype(np.int32), (g * 3).astype(np.int32)])[:3000]
tag = os.path.basename(os.environ.get("C3R_LIB", "default"))
sets = [("keras_init", weights.synthetic(18, sharpen=8.0))] + [
    ("adversarial rs=%.2f fb=%.1f" % (rs, fb), weights.adversarial(18, recurrent_scale=rs, forget_bias=fb))
    for rs, fb in ((1.0, 3.0), (1.25, 2.0), (1.5, 2.0), (1.5, 3.0), (2.0, 3.0), (3.0, 3.0))]
for kind, w in sets:
    eng = Engine(0, 18, nn_impl=1); eng.set_weights(w)
    p, ms = eng.forward(x)
    inter = {}
    ref = model.forward(w, x, intermediates=inter)
    h1 = eng.debug_fetch(0, len(x)); h2 = eng.debug_fetch(2, len(x))
    e = np.abs(p - ref).max(axis=1)
    print(tag, kind, "max|dp| %.3e mean %.3e  h1 err %.2e h2 err %.2e  max|h1| %.2f" % (
        e.max(), e.mean(), np.abs(h1 - inter["h1"]).max(), np.abs(h2 - inter["h2"]).max(), np.abs(inter["h1"]).max()), flush=True)
    eng.close()
