"""Pins oracle/model.py (the float reference of the CUDA network) against an independent implementation.

TensorFlow is absent from this image, so the Keras forward of Clair3_P
(/root/reference/clair3_rna/model.py:126-216) cannot be run here.  torch.nn.LSTM is an independent
implementation of the same cell: gate order i, f, g, o (Keras: i, f, c, o), sigmoid recurrent
activation, tanh activation, `bidirectional=True` concatenating [forward, backward] with the backward
outputs put back in forward time order - exactly what Keras' Bidirectional(LSTM(return_sequences=True),
merge_mode="concat") does.  Weight mapping: weight_ih = kernel^T,
weight_hh = recurrent_kernel^T, bias_ih = bias, bias_hh = 0; the `_reverse` parameters are Keras'
backward_layer.  The dense tail is nn.Linear (weight = kernel^T) + torch's own selu + softmax.
Everything in float64; agreement to 1e-9 means the restatement and the library differ only by
summation order.
"""
import numpy as np
import pytest
import torch

from clair3_rna_b200 import weights as W
from oracle import model


def _torch_lstm(w, layer, n_in, units):
    m = torch.nn.LSTM(n_in, units, batch_first=True, bidirectional=True).double()
    with torch.no_grad():
        for d, suf in (("forward", ""), ("backward", "_reverse")):
            getattr(m, "weight_ih_l0" + suf).copy_(torch.from_numpy(np.asarray(w["%s/%s/kernel" % (layer, d)], np.float64).T))
            getattr(m, "weight_hh_l0" + suf).copy_(torch.from_numpy(np.asarray(w["%s/%s/recurrent_kernel" % (layer, d)], np.float64).T))
            getattr(m, "bias_ih_l0" + suf).copy_(torch.from_numpy(np.asarray(w["%s/%s/bias" % (layer, d)], np.float64)))
            getattr(m, "bias_hh_l0" + suf).zero_()
    return m


def _linear(w, name):
    k = np.asarray(w[name + "/kernel"], np.float64)
    m = torch.nn.Linear(k.shape[0], k.shape[1]).double()
    with torch.no_grad():
        m.weight.copy_(torch.from_numpy(k.T))
        m.bias.copy_(torch.from_numpy(np.asarray(w[name + "/bias"], np.float64)))
    return m


def _independent_forward(w, x, channels):
    """Clair3_P.call with library modules only (no code shared with oracle/model.py)."""
    with torch.no_grad():
        t = torch.from_numpy(np.asarray(x, np.float64))
        h1, _ = _torch_lstm(w, "LSTM1", channels, 128)(t)
        h2, _ = _torch_lstm(w, "LSTM2", 256, 160)(h1)
        flat = torch.flatten(h2, 1)                                   # row-major: index = t * 320 + j
        l4 = torch.selu(_linear(w, "L4")(flat))
        a1 = torch.selu(_linear(w, "L5_1")(l4))
        a2 = torch.selu(_linear(w, "L5_2")(l4))
        y1 = torch.softmax(torch.selu(_linear(w, "Y_gt21_logits")(a1)), dim=1)    # SELU before the softmax (model.py:154,195)
        y2 = torch.softmax(torch.selu(_linear(w, "Y_genotype_logits")(a2)), dim=1)
        return h1.numpy(), h2.numpy(), l4.numpy(), torch.cat([y1, y2], 1).numpy()


@pytest.mark.parametrize("channels", [18, 30])
@pytest.mark.parametrize("kind", ["keras_init", "adversarial"])
def test_oracle_forward_equals_torch_nn_lstm(channels, kind):
    w = W.synthetic(channels, sharpen=8.0) if kind == "keras_init" else W.adversarial(channels)
    rng = np.random.default_rng(5)
    # read-count-like inputs: non-negative counts with the reference channel negated, some deep columns
    x = rng.integers(0, 40, size=(24, 33, channels)).astype(np.int32)
    x[:, :, 0] = -x[:, :, 0]
    x[3] *= 6
    inter = {}
    p = model.forward(w, x, dtype=torch.float64, intermediates=inter)
    h1, h2, l4, q = _independent_forward(w, x, channels)
    assert np.abs(inter["h1"] - h1).max() < 1e-9
    assert np.abs(inter["h2"] - h2).max() < 1e-9
    assert np.abs(inter["l4"] - l4).max() < 1e-8
    assert np.abs(p.astype(np.float64) - q).max() < 1e-6            # forward() returns float32
    assert np.allclose(q[:, :21].sum(1), 1.0) and np.allclose(q[:, 21:].sum(1), 1.0)


def test_lstm_dir_is_one_direction_of_nn_lstm():
    """_lstm_dir alone, both directions, against the matching half of nn.LSTM's output."""
    w = W.synthetic(18)
    rng = np.random.default_rng(11)
    x = torch.from_numpy(rng.standard_normal((5, 33, 18)))
    m = _torch_lstm(w, "LSTM1", 18, 128)
    with torch.no_grad():
        ref, _ = m(x)
    for d, rev, sl in (("forward", False, slice(0, 128)), ("backward", True, slice(128, 256))):
        got = model._lstm_dir(x, *(torch.from_numpy(np.asarray(w["LSTM1/%s/%s" % (d, k)], np.float64))
                                   for k in ("kernel", "recurrent_kernel", "bias")), rev)
        assert (got - ref[:, :, sl]).abs().max().item() < 1e-10


def test_float32_oracle_is_close_to_float64():
    """The fp32 oracle the GPU is compared with (|dp| <= 1e-3) is itself within 1e-5 of the fp64 forward."""
    w = W.synthetic(18, sharpen=8.0)
    rng = np.random.default_rng(3)
    x = rng.integers(0, 40, size=(16, 33, 18)).astype(np.int32)
    p32 = model.forward(w, x)
    p64 = model.forward(w, x, dtype=torch.float64)
    assert np.abs(p32 - p64).max() < 1e-5
