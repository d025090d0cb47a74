"""Network weights in Keras layout (what clair3_rna/model.py:126-156 creates).

Real checkpoints are TF TensorBundles (read by tf_bundle.py; none is on disk here), so tests
and benches use seeded synthetic weights with Keras' default initialisers (Glorot-uniform
kernels, orthogonal recurrent kernels, zero bias with forget-gate bias 1).  The neutral
on-disk format is an .npz of the arrays below.
"""
import numpy as np

U1, U2, NT, DENSE = 128, 160, 33, 128

def shapes(channels: int) -> dict:
    s = {}
    for d in ("forward", "backward"):
        s["LSTM1/%s/kernel" % d] = (channels, 4 * U1)
        s["LSTM1/%s/recurrent_kernel" % d] = (U1, 4 * U1)
        s["LSTM1/%s/bias" % d] = (4 * U1,)
        s["LSTM2/%s/kernel" % d] = (2 * U1, 4 * U2)
        s["LSTM2/%s/recurrent_kernel" % d] = (U2, 4 * U2)
        s["LSTM2/%s/bias" % d] = (4 * U2,)
    s["L4/kernel"] = (NT * 2 * U2, DENSE)
    s["L4/bias"] = (DENSE,)
    for n in ("L5_1", "L5_2"):
        s[n + "/kernel"] = (DENSE, DENSE)
        s[n + "/bias"] = (DENSE,)
    s["Y_gt21_logits/kernel"] = (DENSE, 21)
    s["Y_gt21_logits/bias"] = (21,)
    s["Y_genotype_logits/kernel"] = (DENSE, 3)
    s["Y_genotype_logits/bias"] = (3,)
    return s


def synthetic(channels: int = 18, seed: int = 20260, sharpen: float = 1.0, input_scale: float = 0.25) -> dict:
    """Keras-default initialisation; `sharpen` scales the two head kernels so the
    softmax outputs are not near-uniform; `input_scale` scales LSTM1's input kernel
    (raw read counts reach the hundreds)."""
    rng = np.random.default_rng(seed)
    w = {}
    for name, shp in shapes(channels).items():
        if name.endswith("recurrent_kernel"):
            u = shp[0]
            a = rng.standard_normal((4 * u, u))
            q, r = np.linalg.qr(a)
            q = q * np.sign(np.diag(r))
            w[name] = np.ascontiguousarray(q.T[:u, :]).astype(np.float32)
        elif name.endswith("kernel"):
            lim = np.sqrt(6.0 / (shp[0] + shp[1]))
            k = rng.uniform(-lim, lim, shp)
            if name.startswith("LSTM1") :
                k *= input_scale
            if name.startswith("Y_"):
                k *= sharpen
            w[name] = k.astype(np.float32)
        else:
            b = np.zeros(shp, np.float32)
            if name.startswith("LSTM"):
                u = shp[0] // 4
                b[u:2 * u] = 1.0
            else:
                b += rng.uniform(-0.05, 0.05, shp).astype(np.float32)
            w[name] = b
    return w


def adversarial(channels: int = 18, seed: int = 7, sharpen: float = 8.0, recurrent_scale: float = 3.0,
                forget_bias: float = 3.0) -> dict:
    """A second, harsher weight set for tolerance tests: recurrent kernels scaled up (orthogonal x recurrent_scale, so
    every step amplifies a perturbation of h by up to that factor), forget bias around forget_bias, random non-zero
    LSTM biases - gates saturate and the cell state grows, unlike anything Keras' initialisers produce."""
    w = synthetic(channels, seed=seed, sharpen=sharpen)
    rng = np.random.default_rng(seed + 1)
    for k in list(w):
        if k.endswith("recurrent_kernel"):
            w[k] = (w[k] * recurrent_scale).astype(np.float32)
        elif k.startswith("LSTM") and k.endswith("bias"):
            u = w[k].shape[0] // 4
            b = rng.uniform(-0.5, 0.5, w[k].shape).astype(np.float32)
            b[u:2 * u] += forget_bias
            w[k] = b
    return w


def save(path: str, w: dict) -> None:
    np.savez(path, **{k.replace("/", "__"): v for k, v in w.items()})


def load(path: str) -> dict:
    """.npz in the layout above, or the prefix of a Keras TF-format checkpoint (`<prefix>.index` +
    `<prefix>.data-*`, what the reference's --chkpnt_fn points at) read by tf_bundle.py without TensorFlow."""
    from . import tf_bundle
    if not path.endswith(".npz") and tf_bundle.is_checkpoint_prefix(path):
        return tf_bundle.load_keras_checkpoint(path)
    z = np.load(path)
    return {k.replace("__", "/"): np.ascontiguousarray(z[k], dtype=np.float32) for k in z.files}
