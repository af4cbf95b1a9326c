"""MTIO inference oracle (oracle/mtio_oracle.py) against the golden vectors written from the UNMODIFIED reference
model (oracle/make_golden_mtio.py), plus properties of the restatement the CUDA path relies on.  CPU only."""
import numpy as np
import pytest

from helpers import load_golden
from oracle import mtio_oracle as mo


@pytest.mark.parametrize("case,bias", [("nobias", False), ("bias", True)])
def test_oracle_matches_reference_golden(case, bias):
    g = load_golden("mtio_kat.npz")
    sd = mo.seeded_mtio_state_dict(int(g[f"{case}_seed"]), bias=bias)
    pred, tokens = mo.sample(sd, g[f"{case}_history"], g[f"{case}_current"], 15, return_tokens=True)
    np.testing.assert_allclose(pred, g[f"{case}_pred"], rtol=0, atol=2e-5)      # reference: torch fp32 on CPU
    np.testing.assert_allclose(tokens, g[f"{case}_tokens"], rtol=0, atol=2e-5)
    assert pred.shape == (g[f"{case}_history"].shape[0], 15, 2) and pred.min() >= 0 and pred.max() <= 1


def test_decoder_is_causal_and_encoder_step_independent():
    """What the key/value cache of the CUDA path relies on: with a causal mask, the decoder output of token j does
    not change when later tokens are appended (mtio.py:120-123 re-runs the whole prefix every step)."""
    sd = mo.seeded_mtio_state_dict(5, bias=True)
    hist, cur = mo.synthetic_history(3, 9)
    pe = mo.positional_encoding(64)
    memory = mo.encode(sd, mo.embed(sd, np.concatenate([hist] * 3, axis=-1), pe))
    assert memory.shape == (3, 3, 512)                      # DistillLayer: 5 source tokens -> 3
    _, tokens = mo.sample(sd, hist, cur, 6, return_tokens=True)
    full = mo.decode(sd, mo.embed(sd, tokens, pe), memory)
    for j in (1, 3, 6):
        part = mo.decode(sd, mo.embed(sd, tokens[:, :j], pe), memory)
        np.testing.assert_allclose(part, full[:, :j], rtol=0, atol=5e-6)


def test_distill_pooling_windows():
    """MaxPool1d(3, 2, 1) over 5 tokens: windows {0,1}, {1,2,3}, {3,4} (customized_transformer.py:28)."""
    sd = mo.seeded_mtio_state_dict(6, bias=True)
    x = np.random.default_rng(0).normal(size=(2, 5, 512)).astype(np.float32)
    out = mo.distill(sd, x)
    w = sd["transformer.distill_layer.downConv.weight"]
    y = sum(np.roll(x, 1 - k, axis=1) @ w[:, :, k].T for k in range(3)) + sd["transformer.distill_layer.downConv.bias"]
    y = (y - sd["transformer.distill_layer.norm.running_mean"]) / np.sqrt(sd["transformer.distill_layer.norm.running_var"] + 1e-5) \
        * sd["transformer.distill_layer.norm.weight"] + sd["transformer.distill_layer.norm.bias"]
    y = np.where(y > 0, y, np.expm1(np.minimum(y, 0)))
    np.testing.assert_allclose(out[:, 0], np.maximum(y[:, 0], y[:, 1]), atol=1e-5)
    np.testing.assert_allclose(out[:, 1], y[:, 1:4].max(axis=1), atol=1e-5)
    np.testing.assert_allclose(out[:, 2], np.maximum(y[:, 3], y[:, 4]), atol=1e-5)


def test_wrap_unit():
    v = np.array([-0.25, 0.0, 0.5, 1.0, 1.25, -1.5, 2.75], dtype=np.float32)
    np.testing.assert_allclose(mo.wrap_unit(v), [0.75, 0.0, 0.5, 1.0, 0.25, 0.5, 0.75])


def test_host_module_positional_encoding_matches_oracle():
    from mansy_immersivevideostreaming_b200 import mtio
    assert np.array_equal(mtio.positional_encoding(20), mo.positional_encoding(20))


def test_mtio_refuses_without_gpu(built_library):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mansy_immersivevideostreaming_b200 import _capi, mtio
    with pytest.raises(_capi.MansyError):
        mtio.ViewportTransformerMTIO(device="cuda")
    with pytest.raises(_capi.MansyError):
        mtio.ViewportTransformerMTIO(device="cpu")


def test_viewport_windows_follow_the_dataset_cut():
    """load_dataset.py:48-57 with sample_step == frequency: the sample of a chunk ends right before the chunk."""
    from mansy_immersivevideostreaming_b200.mtio import viewport_windows
    g = np.random.default_rng(0).random((2, 4, 5, 2)).astype(np.float32)
    h, c = viewport_windows(g)
    assert h.shape == (8, 5, 2) and c.shape == (8, 1, 2)
    pts = g.reshape(2, 20, 2)
    for p in range(2):
        for i in range(1, 4):
            t = 5 * i - 1                                   # `current` = last point of the previous chunk
            assert np.array_equal(c[p * 4 + i, 0], pts[p, t])
            if i >= 2:
                assert np.array_equal(h[p * 4 + i], pts[p, t - 5:t])
    assert np.array_equal(c[0, 0], pts[0, 0]) and np.array_equal(h[0], np.repeat(pts[0, :1], 5, axis=0))


def test_prediction_files_round_trip(tmp_path):
    """predict.py:50-65 formats: the pickle holds (chunk, gt uint8[64], pred uint8[64], float64) tuples that hmdtrace.py
    indexes, the csv the same rows as text."""
    import os
    import pickle
    from mansy_immersivevideostreaming_b200.mtio import write_prediction_files
    from mansy_immersivevideostreaming_b200.tables import mask_to_bits
    rng = np.random.default_rng(3)
    gt = rng.integers(0, 2, size=(7, 64)).astype(np.uint8)
    pred = rng.integers(0, 2, size=(7, 64)).astype(np.uint8)
    acc = rng.random(7)
    path = write_prediction_files(str(tmp_path), 4, 9, 3, [mask_to_bits(m) for m in gt], [mask_to_bits(m) for m in pred], acc)
    assert path.endswith(os.path.join("video4", "user9.pkl"))
    rows = pickle.load(open(path, "rb"))
    assert [r[0] for r in rows] == list(range(3, 10))
    for i, (chunk, g, p, a) in enumerate(rows):
        assert g.dtype == np.uint8 and g.shape == (64,) and np.array_equal(g, gt[i]) and np.array_equal(p, pred[i]) and a == acc[i]
    lines = open(path.replace(".pkl", ".csv")).read().splitlines()
    assert lines[0] == "chunk,gt,pred,accuracy" and len(lines) == 8
    f = lines[1].split(",")
    assert int(f[0]) == 3 and [int(x) for x in f[1:65]] == list(gt[0]) and [int(x) for x in f[65:129]] == list(pred[0])
    assert float(f[129]) == acc[0]


def test_prediction_files_reproduce_a_shipped_file(tmp_path):
    import os
    import pickle
    src = "/root/reference/datasets/Jin2022/viewports/prediction/video1/user1.pkl"
    if not os.path.exists(src):
        pytest.skip("reference dataset not present (build container only)")
    from mansy_immersivevideostreaming_b200.mtio import write_prediction_files
    from mansy_immersivevideostreaming_b200.tables import mask_to_bits
    rows = pickle.load(open(src, "rb"))
    path = write_prediction_files(str(tmp_path), 1, 1, rows[0][0], [mask_to_bits(r[1]) for r in rows],
                                  [mask_to_bits(r[2]) for r in rows], [r[3] for r in rows])
    back = pickle.load(open(path, "rb"))
    assert len(back) == len(rows)
    for a, b in zip(rows, back):
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3] == b[3] and b[1].dtype == a[1].dtype
    if os.path.exists(src.replace(".pkl", ".csv")):
        assert open(src.replace(".pkl", ".csv")).read() == open(path.replace(".pkl", ".csv")).read()


def test_linreg_oracle_matches_reference_golden():
    g = load_golden("linreg_kat.npz")
    got = mo.linreg_sample(g["history"], g["current"], 15)
    np.testing.assert_allclose(got, g["pred"], rtol=0, atol=5e-7)       # sklearn lstsq vs closed form, float32 storage
