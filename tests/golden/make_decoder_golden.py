#!/usr/bin/env python3
"""Golden VCF rows from the REFERENCE's own decoder (build container only).

Imports /root/reference/clair3_rna/call_variants.py verbatim (with a stub `tensorflow`
module: its module-level import is the only obstacle) and calls its `output_with` on
(position, ref33, alt_info) of the pileup goldens with two probability sources:
the oracle network (synthetic weights) and seeded random peaked distributions that reach
every outcome family.  Output: tests/golden/decoder_<case>.npz {probs, rows}.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_ROOT = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF_ROOT)

tf = types.ModuleType("tensorflow")
tfp = types.ModuleType("tensorflow.python"); tfu = types.ModuleType("tensorflow.python.util")
tfd = types.ModuleType("tensorflow.python.util.deprecation")
sys.modules.update({"tensorflow": tf, "tensorflow.python": tfp, "tensorflow.python.util": tfu,
                    "tensorflow.python.util.deprecation": tfd})
tfu.deprecation = tfd
class _K:
    class Model: pass
    class regularizers:
        @staticmethod
        def l2(x): return None
tf.keras = _K; tf.float32 = "float32"
tf.get_logger = lambda: types.SimpleNamespace(setLevel=lambda *_: None)

import clair3_rna.call_variants as cv       # noqa: E402  (the reference, verbatim)
import shared.param_p as param_p            # noqa: E402
cv.param = param_p

from clair3_rna_b200 import weights         # noqa: E402
from oracle import model                    # noqa: E402


def reference_rows(contig, pos, ref33, alt_info, probs):
    rows = []
    cfg = cv.OutputConfig(is_show_reference=True, is_debug=False, is_haploid_precise_mode_enabled=False,
                          is_haploid_sensitive_mode_enabled=False, is_output_for_ensemble=False,
                          quality_score_for_pass=2, tensor_fn="PIPE", input_probabilities=False,
                          add_indel_length=False, gvcf=False, pileup=True, enable_long_indel=False,
                          maximum_variant_length_that_need_infer=param_p.maximum_variant_length_that_need_infer,
                          keep_iupac_bases=False)
    util = cv.OutputUtilities(None, rows.append, None, None, None)
    out = []
    for i in range(len(pos)):
        n0 = len(rows)
        cv.output_with("%s:%d:%s" % (contig, pos[i], ref33[i]), str(alt_info[i]), probs[i, :21], probs[i, 21:24],
                       0, 0, cfg, util)
        out.append(rows[n0] if len(rows) > n0 else "")
    return out


def random_probs(n, seed):
    rng = np.random.default_rng(seed)
    def soft(k, temp):
        z = rng.standard_normal((n, k)) * temp
        z -= z.max(1, keepdims=True)
        e = np.exp(z)
        return (e / e.sum(1, keepdims=True)).astype(np.float32)
    return np.concatenate([soft(21, 3.0), soft(3, 2.0)], axis=1)


def main():
    for name, C in (("cfg1_ont_drna", 18), ("ties_lowdepth", 18), ("pad_dense", 18), ("phased_noisy", 30)):
        g = np.load(os.path.join(HERE, name + ".npz"))
        pos, ref33, alt = g["pos"], [str(s) for s in g["ref33"]], [str(s) for s in g["alt_info"]]
        p_net = model.forward(weights.synthetic(C, sharpen=8.0), g["tensor"])
        p_rnd = random_probs(len(pos), 4242)
        probs = np.concatenate([p_net, p_rnd])
        rows = reference_rows("chr1", np.concatenate([pos, pos]), ref33 + ref33, alt + alt, probs)
        np.savez_compressed(os.path.join(HERE, "decoder_" + name + ".npz"), probs=probs, rows=np.array(rows))
        kinds = {}
        for r in rows:
            k = r.split("\t")[9].split(":")[0] if r else "none"
            kinds[k] = kinds.get(k, 0) + 1
        print(name, len(rows), kinds)


if __name__ == "__main__":
    main()
