"""Generate ``tests/golden/mtio_kat.npz`` from the UNMODIFIED reference MTIO model -- build container only.

    python -m oracle.make_golden_mtio

TEST INFRASTRUCTURE.  Imports ``viewport_prediction/models/mtio.py`` from /root/reference (with a stand-in for the
absent ``munch``), loads numpy-seeded weights (``oracle.mtio_oracle.seeded_mtio_state_dict``; the weights
themselves are NOT stored, only the seed), runs ``model.sample`` on CPU in exact fp32 and asserts that the
numpy restatement agrees before writing the fixture.

Two cases:
  * ``nobias``: the model exactly as the reference constructs it under this container's torch 2.11, where the
    positional ``device, dtype`` arguments of customized_transformer.py:47-50 land on ``bias`` (so: no biases);
  * ``bias``: the same reference module objects with encoder / decoder stacks rebuilt by ``nn.Transformer``
    with ``bias=True`` -- what the reference constructs under its pinned torch (<= 2.0).  The reference's
    ``sample`` / ``_process_src_tgt`` / ``Transformer.forward`` / ``DistillLayer`` code runs unmodified on it.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
import warnings

import numpy as np

from oracle import mtio_oracle as mo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_ROOT = os.environ.get("MANSY_REFERENCE_ROOT", "/root/reference")


def load_reference_mtio():
    if "munch" not in sys.modules:
        m = types.ModuleType("munch")

        class Munch(dict):
            __getattr__ = dict.__getitem__
            __setattr__ = dict.__setitem__

        m.Munch = Munch
        sys.modules["munch"] = m
    vp = os.path.join(REFERENCE_ROOT, "viewport_prediction")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in ("utils", "models")}
    sys.path.insert(0, vp)
    try:
        mtio = importlib.import_module("models.mtio")
        ct = importlib.import_module("models.customized_transformer")
    finally:
        sys.path.remove(vp)
        for k in list(sys.modules):
            if k.split(".")[0] in ("utils", "models"):
                sys.modules.pop(k)
        sys.modules.update(saved)
    return mtio, ct


def build_model(mtio, bias: bool):
    import torch
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = mtio.ViewportTransformerMTIO(in_channel=2, fut_window=15, d_model=512, dim_feedforward=512,
                                           num_encoder_layers=2, num_decoder_layers=2, device="cpu", seed=1)
        has_bias = any(k.endswith("in_proj_bias") for k in net.state_dict())
        if bias and not has_bias:
            stock = torch.nn.Transformer(d_model=512, nhead=8, num_encoder_layers=2, num_decoder_layers=2,
                                         dim_feedforward=512, batch_first=True, bias=True)
            net.transformer.encoder = stock.encoder
            net.transformer.decoder = stock.decoder
        elif not bias and has_bias:
            raise RuntimeError("this torch builds the reference with biases; the nobias case cannot be generated")
    return net.eval()


def run_case(mtio, bias: bool, seed: int, n: int):
    import torch
    net = build_model(mtio, bias)
    sd = mo.seeded_mtio_state_dict(seed, bias=bias)
    tsd = {k: torch.from_numpy(v) for k, v in sd.items()}
    cur = net.state_dict()
    tsd["positional_embedding.pe"] = cur["positional_embedding.pe"]
    tsd["transformer.distill_layer.norm.num_batches_tracked"] = cur["transformer.distill_layer.norm.num_batches_tracked"]
    net.load_state_dict(tsd, strict=True)
    hist, cur_pt = mo.synthetic_history(n, seed + 100)
    torch.backends.mkldnn.enabled = True
    with torch.no_grad():
        ref = net.sample(torch.from_numpy(hist), torch.from_numpy(cur_pt)).numpy()
    got, tokens = mo.sample(sd, hist, cur_pt, 15, return_tokens=True)
    err = float(np.max(np.abs(got - ref)))
    assert err <= 2e-5, f"oracle disagrees with the reference model (bias={bias}): {err}"
    # the registered buffer vs the restated positional encoding
    pe_err = float(np.max(np.abs(cur["positional_embedding.pe"][0, :64].numpy() - mo.positional_encoding(64))))
    assert pe_err <= 1e-5, pe_err      # float32 sin/cos of torch vs numpy
    return hist, cur_pt, ref, tokens, err


def main() -> None:
    mtio, _ = load_reference_mtio()
    out = {}
    for name, bias, seed, n in (("nobias", False, 11, 12), ("bias", True, 12, 12)):
        hist, cur_pt, ref, tokens, err = run_case(mtio, bias, seed, n)
        out[f"{name}_seed"] = np.int64(seed)
        out[f"{name}_history"] = hist
        out[f"{name}_current"] = cur_pt
        out[f"{name}_pred"] = ref
        out[f"{name}_tokens"] = tokens
        print(f"{name}: reference vs restatement max abs err {err:.2e}; pred range [{ref.min():.3f}, {ref.max():.3f}]")
    path = os.path.join(ROOT, "tests", "golden", "mtio_kat.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)
    make_linreg_golden()


def make_linreg_golden() -> None:
    """``--model regression``: the reference's LinearRegression.sample (sklearn fits in a Python loop) on seeded walks."""
    import importlib.util

    import torch
    spec = importlib.util.spec_from_file_location(
        "_ref_linreg", os.path.join(REFERENCE_ROOT, "viewport_prediction", "models", "linear_regression.py"))
    lr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(lr)
    hist, cur = mo.synthetic_history(64, 5)
    ref = lr.LinearRegression(15).sample(torch.from_numpy(hist), torch.from_numpy(cur)).numpy()
    err = float(np.max(np.abs(ref - mo.linreg_sample(hist, cur, 15))))
    assert err <= 5e-7, err
    path = os.path.join(ROOT, "tests", "golden", "linreg_kat.npz")
    np.savez_compressed(path, history=hist, current=cur, pred=ref)
    print(f"linreg: reference vs closed form max abs err {err:.2e}; wrote {path}")


if __name__ == "__main__":
    main()
