import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from clair3_rna_b200 import weights
from clair3_rna_b200.engine import Engine
g = np.load("tests/golden/cfg1_ont_drna.npz")
x = np.concatenate([g["tensor"]] * 40)[:14617]
eng = Engine(0, 18); eng.set_weights(weights.synthetic(18, sharpen=8.0))
for _ in range(3): p, ms = eng.forward(x)
t = [eng.forward(x)[1] for _ in range(5)]
print("forward ms", min(t), flush=True)
