"""The tensor-core network on batches of several rounds of the recurrent kernels plus a remainder (>= 20 000
sites), against the fp32 oracle, on three weight sets:

  keras_init  Keras' initialisers, heads sharpened x8                         gate: |dp| <= 1e-3 (north_star)
  stress      forget bias 1.5, random LSTM biases in +-0.5                      gate: |dp| <= 1e-3
  harsh       the same with recurrent kernels x3 and forget bias 3              reported, bounded at 5e-2

The network's operands are fp16 (fp32 accumulate); what limits its accuracy is the fp16 rounding of the RECURRENT
kernels U1 / U2 (measured term by term with a CPU emulation of the operand rounding): a
relative weight perturbation of 2^-12 that the recurrence amplifies by its own gain - 3.9e-4 on keras_init, ~6e-4
on stress, 2e-2 when every step multiplies a perturbation by up to 3.  The harsh set documents that envelope;
nn_impl = 0 (fp32) is the exact path for such weights.
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


@pytest.mark.parametrize("kind,C,n", [("keras_init", 18, 21000), ("stress", 18, 21000), ("stress", 30, 9700), ("harsh", 18, 4000)])
def test_large_batch_against_oracle(kind, C, n):
    import torch
    from clair3_rna_b200 import weights
    from clair3_rna_b200.engine import Engine
    from oracle import model
    w = {"keras_init": lambda: weights.synthetic(C, sharpen=8.0),
         "stress": lambda: weights.adversarial(C, recurrent_scale=1.0, forget_bias=1.5),
         "harsh": lambda: weights.adversarial(C, recurrent_scale=3.0, forget_bias=3.0)}[kind]()
    tol = 5e-2 if kind == "harsh" else PROB_TOL
    x = _tensor(n, C)
    eng = Engine(0, C, nn_impl=1)
    eng.set_weights(w)
    p, ms = eng.forward(x)
    eng.close()
    torch.set_num_threads(max(1, (os.cpu_count() or 2) - 1))
    ref = np.concatenate([model.forward(w, x[o:o + 2048]) for o in range(0, n, 2048)])
    err = np.abs(p - ref).max(axis=1)
    print(kind, C, n, "max |dp| %.3e  mean %.3e  device ms %.3f" % (err.max(), err.mean(), ms))
    assert err.max() <= tol, (kind, float(err.max()), int(err.argmax()))
    # calls: the argmax of both heads agrees wherever the oracle's top two are further apart than the tolerance
    for lo, hi in ((0, 21), (21, 24)):
        a, b = p[:, lo:hi], ref[:, lo:hi]
        top2 = np.sort(b, axis=1)[:, -2:]
        clear = (top2[:, 1] - top2[:, 0]) > 2 * tol
        assert np.array_equal(a.argmax(1)[clear], b.argmax(1)[clear])
