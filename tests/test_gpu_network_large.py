"""The tensor-core network on batches of several rounds of the recurrent kernels plus a remainder (>= 20 000
sites), against the fp32 oracle, on two weight sets: Keras-style initialisation with sharpened heads, and the
adversarial set (recurrent kernels x3, forget bias ~3).  Tolerance: |dp| <= 1e-3 absolute (BASELINE.json north_star).
The oracle runs every site: the multi-round / two-stream path of tc_forward is compared with an independent
implementation, not with itself."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
PROB_TOL = 1e-3


def _tensor(n, C):
    name = "cfg1_ont_drna" if C == 18 else "phased_noisy"
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))["tensor"]
    rng = np.random.default_rng(17)
    reps = [g]
    while sum(len(r) for r in reps) < n:
        # variations of real windows: scaled depth and shuffled sites, so that no two tiles are alike
        k = rng.integers(1, 4)
        reps.append((g[rng.permutation(len(g))] * k).astype(np.int32))
    return np.concatenate(reps)[:n]


@pytest.mark.parametrize("kind,C,n", [("keras_init", 18, 21000), ("adversarial", 18, 21000), ("adversarial", 30, 9700)])
def test_large_batch_against_oracle(kind, C, n):
    import torch
    from clair3_rna_b200 import weights
    from clair3_rna_b200.engine import Engine
    from oracle import model
    w = weights.synthetic(C, sharpen=8.0) if kind == "keras_init" else weights.adversarial(C)
    x = _tensor(n, C)
    eng = Engine(0, C, nn_impl=1)
    eng.set_weights(w)
    p, ms = eng.forward(x)
    eng.close()
    torch.set_num_threads(max(1, (os.cpu_count() or 2) - 1))
    ref = np.concatenate([model.forward(w, x[o:o + 2048]) for o in range(0, n, 2048)])
    err = np.abs(p - ref).max(axis=1)
    print(kind, C, n, "max |dp| %.3e  mean %.3e  device ms %.3f" % (err.max(), err.mean(), ms))
    assert err.max() <= PROB_TOL, (kind, float(err.max()), int(err.argmax()))
    # calls: the argmax of both heads agrees wherever the oracle's top two are further apart than the tolerance
    for lo, hi in ((0, 21), (21, 24)):
        a, b = p[:, lo:hi], ref[:, lo:hi]
        top2 = np.sort(b, axis=1)[:, -2:]
        clear = (top2[:, 1] - top2[:, 0]) > 2 * PROB_TOL
        assert np.array_equal(a.argmax(1)[clear], b.argmax(1)[clear])
