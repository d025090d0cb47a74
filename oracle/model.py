"""ORACLE (test infrastructure).  PyTorch-CPU restatement of Clair3_P's forward.

Source: /root/reference/clair3_rna/model.py:126-216 (Keras defaults: LSTM gate
order i,f,c,o; sigmoid recurrent activation, tanh activation; Bidirectional
concat [fwd, bwd]; Flatten; Dense+SELU; SELU *then* softmax on both heads).
TensorFlow is absent from this image, so this is the float reference the CUDA
network is compared with (|dp| <= 1e-3, BASELINE.json north_star).
"""
import numpy as np
import torch

SELU_SCALE = 1.0507009873554805
SELU_ALPHA = 1.6732632423543772


def selu(x):
    return SELU_SCALE * torch.where(x > 0, x, SELU_ALPHA * torch.expm1(x))


def _lstm_dir(x, W, U, b, reverse):
    """x [n,T,in]; Keras kernels W[in,4u], U[u,4u], b[4u] -> h [n,T,u] in forward time order."""
    n, T, _ = x.shape
    u = U.shape[0]
    h = x.new_zeros(n, u)
    c = x.new_zeros(n, u)
    out = [None] * T
    zx = x @ W + b
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        z = zx[:, t, :] + h @ U
        i = torch.sigmoid(z[:, 0:u])
        f = torch.sigmoid(z[:, u:2 * u])
        g = torch.tanh(z[:, 2 * u:3 * u])
        o = torch.sigmoid(z[:, 3 * u:4 * u])
        c = f * c + i * g
        h = o * torch.tanh(c)
        out[t] = h
    return torch.stack(out, dim=1)


def forward(weights: dict, tensor: np.ndarray, dtype=torch.float32, intermediates: dict | None = None) -> np.ndarray:
    """int32 [n,33,C] -> probabilities [n,24] (21 gt21 + 3 genotype).
    `intermediates`, if given, receives h1 [n,33,256], zx2 [n,33,2,640], h2 [n,33,320], l4 [n,128]."""
    w = {k: torch.from_numpy(np.asarray(v)).to(dtype) for k, v in weights.items()}
    x = torch.from_numpy(np.asarray(tensor)).to(dtype)
    with torch.no_grad():
        for layer in ("LSTM1", "LSTM2"):
            if intermediates is not None and layer == "LSTM2":
                intermediates["zx2"] = torch.stack(
                    [x @ w["LSTM2/%s/kernel" % d] + w["LSTM2/%s/bias" % d] for d in ("forward", "backward")], dim=2).numpy()
            f = _lstm_dir(x, w[layer + "/forward/kernel"], w[layer + "/forward/recurrent_kernel"],
                          w[layer + "/forward/bias"], False)
            b = _lstm_dir(x, w[layer + "/backward/kernel"], w[layer + "/backward/recurrent_kernel"],
                          w[layer + "/backward/bias"], True)
            x = torch.cat([f, b], dim=2)
            if intermediates is not None:
                intermediates["h1" if layer == "LSTM1" else "h2"] = x.numpy()
        v = x.reshape(x.shape[0], -1)
        l4 = selu(v @ w["L4/kernel"] + w["L4/bias"])
        if intermediates is not None:
            intermediates["l4"] = l4.numpy()
        a1 = selu(l4 @ w["L5_1/kernel"] + w["L5_1/bias"])
        a2 = selu(l4 @ w["L5_2/kernel"] + w["L5_2/bias"])
        y1 = torch.softmax(selu(a1 @ w["Y_gt21_logits/kernel"] + w["Y_gt21_logits/bias"]), dim=1)
        y2 = torch.softmax(selu(a2 @ w["Y_genotype_logits/kernel"] + w["Y_genotype_logits/bias"]), dim=1)
        return torch.cat([y1, y2], dim=1).to(torch.float32).numpy()
