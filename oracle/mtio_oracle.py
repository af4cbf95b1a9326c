"""CPU restatement of the MTIO viewport-prediction transformer's inference path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference arm may import
this module; the product path (``mansy_immersivevideostreaming_b200/mtio.py`` -> ``csrc/mansy_mtio.cu``)
never does.

Follows, in numpy float32:
  * ``ViewportTransformerMTIO.sample``            viewport_prediction/models/mtio.py:106-133
  * ``_process_src_tgt``                          viewport_prediction/models/mtio.py:135-148
  * ``ViewportEmbedding`` / ``PositionalEncoding``  viewport_prediction/models/mtio.py:11-45
  * ``Transformer.forward`` + ``DistillLayer``    viewport_prediction/models/customized_transformer.py:13-36,52-70
  * ``to_position_normalized_cartesian``          viewport_prediction/utils/common.py:61-70
and, for the third-party arithmetic on the path (``torch.nn.Transformer`` with the constructor defaults
the reference passes: post-norm, ReLU, nhead 8, ``batch_first=True``; torch pinned ``<=2.0`` by the
reference's README), the published algorithm of ``nn.TransformerEncoderLayer`` /
``nn.TransformerDecoderLayer`` / ``nn.MultiheadAttention`` / ``nn.LayerNorm`` / ``nn.BatchNorm1d`` (eval).

Pinning: ``oracle/make_golden_mtio.py`` imports the UNMODIFIED reference model in the build container,
loads numpy-seeded weights into it, runs ``model.sample`` on CPU (exact fp32) and asserts agreement with
this restatement at 2e-5 before writing ``tests/golden/mtio_kat.npz``.

Reference quirk kept observable: ``customized_transformer.py:47-50`` passes ``device, dtype`` positionally
to ``nn.Transformer.__init__``; from torch 2.1 on that slot is ``bias``, so under a modern torch the
reference builds a transformer WITHOUT biases in its attention / feed-forward / LayerNorm layers, under the
pinned torch (<=2.0) WITH them.  Missing bias keys are therefore treated as zeros here and in the kernels.
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Optional

import numpy as np

F32 = np.float32
D_MODEL = 512
N_HEAD = 8           # nn.Transformer default (customized_transformer.py:40)
IN_CHANNEL = 2
MTIO_HEADS = 3       # ViewportTransformerMTIO(num_head=3) default (mtio.py:49)
LN_EPS = 1e-5
BN_EPS = 1e-5


def positional_encoding(n: int, d: int = D_MODEL) -> np.ndarray:
    """mtio.py:18-26 (float32 like the registered buffer)."""
    pe = np.zeros((n, d), dtype=F32)
    position = np.arange(0, n, dtype=F32)[:, None]
    div_term = np.exp(np.arange(0, d, 2, dtype=F32) * F32(-(math.log(10000.0) / d))).astype(F32)
    pe[:, 0::2] = np.sin(position * div_term)
    pe[:, 1::2] = np.cos(position * div_term)
    return pe


def _get(sd: Mapping[str, np.ndarray], key: str, default: Optional[np.ndarray] = None) -> np.ndarray:
    if key in sd:
        return np.asarray(sd[key], dtype=F32)
    if default is None:
        raise KeyError(key)
    return default


def _linear(x: np.ndarray, w: np.ndarray, b: Optional[np.ndarray]) -> np.ndarray:
    y = x @ w.T
    return y if b is None else y + b


def _layer_norm(x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    mean = x.mean(axis=-1, keepdims=True, dtype=F32)
    var = ((x - mean) ** 2).mean(axis=-1, keepdims=True, dtype=F32)
    return ((x - mean) / np.sqrt(var + F32(LN_EPS))) * w + b


def _mha(sd, prefix: str, q_in: np.ndarray, kv_in: np.ndarray, causal: bool) -> np.ndarray:
    """nn.MultiheadAttention forward (batch_first), q_in [B,Tq,D], kv_in [B,Tk,D]."""
    d = q_in.shape[-1]
    dh = d // N_HEAD
    w = _get(sd, prefix + "in_proj_weight")
    b = _get(sd, prefix + "in_proj_bias", np.zeros(3 * d, F32))
    q = _linear(q_in, w[:d], b[:d])
    k = _linear(kv_in, w[d:2 * d], b[d:2 * d])
    v = _linear(kv_in, w[2 * d:], b[2 * d:])
    B, Tq, _ = q.shape
    Tk = k.shape[1]
    q = q.reshape(B, Tq, N_HEAD, dh).transpose(0, 2, 1, 3)
    k = k.reshape(B, Tk, N_HEAD, dh).transpose(0, 2, 1, 3)
    v = v.reshape(B, Tk, N_HEAD, dh).transpose(0, 2, 1, 3)
    s = (q * F32(1.0 / math.sqrt(dh))) @ k.transpose(0, 1, 3, 2)
    if causal:
        mask = np.triu(np.ones((Tq, Tk), dtype=bool), k=1)
        s = np.where(mask, F32(-np.inf), s)
    s = s - s.max(axis=-1, keepdims=True)
    p = np.exp(s)
    p = p / p.sum(axis=-1, keepdims=True, dtype=F32)
    o = (p @ v).transpose(0, 2, 1, 3).reshape(B, Tq, d)
    return _linear(o, _get(sd, prefix + "out_proj.weight"), _get(sd, prefix + "out_proj.bias", np.zeros(d, F32)))


def _ln(sd, prefix: str, x: np.ndarray) -> np.ndarray:
    d = x.shape[-1]
    return _layer_norm(x, _get(sd, prefix + "weight"), _get(sd, prefix + "bias", np.zeros(d, F32)))


def _ffn(sd, prefix: str, x: np.ndarray) -> np.ndarray:
    w1, w2 = _get(sd, prefix + "linear1.weight"), _get(sd, prefix + "linear2.weight")
    h = np.maximum(_linear(x, w1, _get(sd, prefix + "linear1.bias", np.zeros(w1.shape[0], F32))), F32(0))
    return _linear(h, w2, _get(sd, prefix + "linear2.bias", np.zeros(w2.shape[0], F32)))


def n_layers(sd: Mapping[str, np.ndarray], side: str) -> int:
    n = 0
    while f"transformer.{side}.layers.{n}.linear1.weight" in sd:
        n += 1
    return n


def encode(sd, src: np.ndarray) -> np.ndarray:
    """TransformerEncoder (post-norm layers + final norm) then DistillLayer: [B,T,D] -> [B,T',D]."""
    x = src
    for l in range(n_layers(sd, "encoder")):
        p = f"transformer.encoder.layers.{l}."
        x = _ln(sd, p + "norm1.", x + _mha(sd, p + "self_attn.", x, x, causal=False))
        x = _ln(sd, p + "norm2.", x + _ffn(sd, p, x))
    x = _ln(sd, "transformer.encoder.norm.", x)
    return distill(sd, x)


def distill(sd, x: np.ndarray) -> np.ndarray:
    """customized_transformer.py:19-36: Conv1d(k=3, circular padding 1) -> BatchNorm1d (eval) -> ELU -> MaxPool1d(3, 2, 1)."""
    w = _get(sd, "transformer.distill_layer.downConv.weight")        # [out, in, 3]
    b = _get(sd, "transformer.distill_layer.downConv.bias")
    B, T, d = x.shape
    y = np.zeros((B, T, d), dtype=F32)
    for k in range(3):
        y += np.roll(x, 1 - k, axis=1) @ w[:, :, k].T                  # token t sees x[(t + k - 1) mod T]
    y = y + b
    g = _get(sd, "transformer.distill_layer.norm.weight")
    beta = _get(sd, "transformer.distill_layer.norm.bias")
    rm = _get(sd, "transformer.distill_layer.norm.running_mean")
    rv = _get(sd, "transformer.distill_layer.norm.running_var")
    y = (y - rm) / np.sqrt(rv + F32(BN_EPS)) * g + beta
    y = np.where(y > 0, y, np.expm1(np.minimum(y, F32(0)))).astype(F32)
    t_out = (T + 2 - 3) // 2 + 1
    out = np.empty((B, t_out, d), dtype=F32)
    for o in range(t_out):
        lo, hi = max(2 * o - 1, 0), min(2 * o + 1, T - 1)
        out[:, o] = y[:, lo:hi + 1].max(axis=1)
    return out


def decode(sd, tgt: np.ndarray, memory: np.ndarray) -> np.ndarray:
    x = tgt
    for l in range(n_layers(sd, "decoder")):
        p = f"transformer.decoder.layers.{l}."
        x = _ln(sd, p + "norm1.", x + _mha(sd, p + "self_attn.", x, x, causal=True))
        x = _ln(sd, p + "norm2.", x + _mha(sd, p + "multihead_attn.", x, memory, causal=False))
        x = _ln(sd, p + "norm3.", x + _ffn(sd, p, x))
    return _ln(sd, "transformer.decoder.norm.", x)


def embed(sd, tokens: np.ndarray, pe: np.ndarray) -> np.ndarray:
    """ViewportEmbedding + PositionalEncoding (eval: dropout is the identity), tokens [B,T,6]."""
    x = _linear(tokens, _get(sd, "embedding.linear.weight"), _get(sd, "embedding.linear.bias"))
    return x + pe[None, :tokens.shape[1]]


def wrap_unit(values: np.ndarray) -> np.ndarray:
    """utils/common.py:61-70 (``.to(torch.int)`` truncates toward zero)."""
    out = values.copy()
    neg, big = values < 0, values > 1
    out[neg] = values[neg] - np.trunc(values[neg]) + F32(1)
    out[big] = values[big] - np.trunc(values[big])
    return out


def sample(sd: Mapping[str, np.ndarray], history: np.ndarray, current: np.ndarray, fut_window: int = 15,
           return_tokens: bool = False):
    """``ViewportTransformerMTIO.sample``: history [B,M,2], current [B,1,2] -> ensembled viewports [B,fut,2]."""
    history = np.asarray(history, dtype=F32)
    current = np.asarray(current, dtype=F32)
    pe = _get(sd, "positional_embedding.pe", positional_encoding(64)[None])[0]
    src = np.concatenate([history] * MTIO_HEADS, axis=-1)
    tgt = np.concatenate([current] * MTIO_HEADS, axis=-1)
    pw, pb = _get(sd, "predictor.0.weight"), _get(sd, "predictor.0.bias")
    outs = []
    memory = encode(sd, embed(sd, src, pe))        # eval mode: the per-step re-encoding of mtio.py:120-123 is a pure recomputation
    for _ in range(fut_window):
        out = decode(sd, embed(sd, tgt, pe), memory)
        z = _linear(out[:, -1], pw, pb)
        pred = (F32(1) / (F32(1) + np.exp(-z))).astype(F32)[:, None, :]
        tgt = np.concatenate([tgt, pred], axis=1)
        ens = np.stack([pred[:, :, [c + j * IN_CHANNEL for j in range(MTIO_HEADS)]].sum(axis=-1) / F32(MTIO_HEADS)
                        for c in range(IN_CHANNEL)], axis=-1)
        outs.append(ens.astype(F32))
    res = wrap_unit(np.concatenate(outs, axis=1))
    return (res, tgt) if return_tokens else res


# deterministic weights / synthetic inputs live in the product package (bench.py's measured arm needs them too and may not
# import oracle/); re-exported here for the tests and the golden scripts
from mansy_immersivevideostreaming_b200.mtio import (mtio_state_dict_shapes, seeded_mtio_state_dict,  # noqa: E402,F401
                                                      synthetic_history)


def linreg_sample(history: np.ndarray, current: np.ndarray, fut_window: int = 15) -> np.ndarray:
    """``LinearRegression.sample`` (viewport_prediction/models/linear_regression.py:16-33): sklearn's
    ``LinearRegression(fit_intercept=True)`` over x = 0..P-1 per sample and coordinate, predictions at P..P+F-1, float64
    arithmetic stored into a float32 tensor.  Restated as the closed-form least-squares line."""
    merge = np.concatenate([np.asarray(history, F32), np.asarray(current, F32)], axis=1).astype(np.float64)    # [B, P, 2]
    P = merge.shape[1]
    x = np.arange(P, dtype=np.float64)
    xbar, ybar = x.mean(), merge.mean(axis=1, keepdims=True)
    dx = (x - xbar)[None, :, None]
    slope = (dx * (merge - ybar)).sum(axis=1) / (dx ** 2).sum()
    icpt = ybar[:, 0] - slope * xbar
    fx = np.arange(P, P + fut_window, dtype=np.float64)[None, :, None]
    return (icpt[:, None, :] + slope[:, None, :] * fx).astype(F32)
