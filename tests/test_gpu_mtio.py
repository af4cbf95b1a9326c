"""MTIO inference on the GPU (csrc/mansy_mtio.cu through the C ABI) against the oracle and the reference golden.

Tolerances: the exact-fp32 CUDA-core path 3e-5 absolute on viewport coordinates in [0,1] (fp32 summation order);
the tcgen05 path 5e-3 absolute -- kind::tf32 keeps 10 mantissa bits of every operand, the precision class the
reference itself runs in (torch.set_float32_matmul_precision('high'), predict.py:100)."""
import numpy as np
import pytest
import torch

from helpers import load_golden
from oracle import mtio_oracle as mo
from oracle import sim_oracle as so

pytestmark = pytest.mark.gpu

TF32_ATOL = 5e-3
FP32_ATOL = 3e-5


def make_model(sd, max_batch=4096, fut=15, **kw):
    from mansy_immersivevideostreaming_b200.mtio import ViewportTransformerMTIO
    net = ViewportTransformerMTIO(in_channel=2, fut_window=fut, d_model=512, dim_feedforward=512, device="cuda:0",
                                  max_batch=max_batch, **kw)
    return net.load_state_dict(sd).eval()


@pytest.mark.parametrize("case,bias", [("nobias", False), ("bias", True)])
def test_reference_golden(case, bias):
    g = load_golden("mtio_kat.npz")
    sd = mo.seeded_mtio_state_dict(int(g[f"{case}_seed"]), bias=bias)
    net = make_model(sd)
    hist, cur = torch.from_numpy(g[f"{case}_history"]).cuda(), torch.from_numpy(g[f"{case}_current"]).cuda()
    net.fp32 = True
    pred, tokens = net.sample(hist, cur, return_tokens=True)
    np.testing.assert_allclose(pred.cpu().numpy(), g[f"{case}_pred"], rtol=0, atol=FP32_ATOL)
    np.testing.assert_allclose(tokens.cpu().numpy(), g[f"{case}_tokens"], rtol=0, atol=FP32_ATOL)
    net.fp32 = False
    pred, tokens = net.sample(hist, cur, return_tokens=True)
    err = np.abs(pred.cpu().numpy() - g[f"{case}_pred"]).max()
    print(f"{case}: tcgen05 TF32 vs reference fp32 max abs err {err:.2e}")
    np.testing.assert_allclose(pred.cpu().numpy(), g[f"{case}_pred"], rtol=0, atol=TF32_ATOL)
    np.testing.assert_allclose(tokens.cpu().numpy(), g[f"{case}_tokens"], rtol=0, atol=TF32_ATOL)


@pytest.mark.parametrize("n", [1, 127, 300])
def test_ragged_batches_vs_oracle(n):
    """Batches that are not a multiple of the 128-row tile (TMA zero-fill + guarded stores), several tiles."""
    sd = mo.seeded_mtio_state_dict(21, bias=True)
    hist, cur = mo.synthetic_history(n, 33)
    want = mo.sample(sd, hist, cur, 15)
    net = make_model(sd)
    for fp32, atol in ((True, FP32_ATOL), (False, TF32_ATOL)):
        net.fp32 = fp32
        got = net.sample(torch.from_numpy(hist).cuda(), torch.from_numpy(cur).cuda()).cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=0, atol=atol)


def test_chunked_and_host_paths_equal_device_path():
    sd = mo.seeded_mtio_state_dict(22, bias=False)
    hist, cur = mo.synthetic_history(300, 34)
    big = make_model(sd, max_batch=512)
    ref = big.sample(torch.from_numpy(hist).cuda(), torch.from_numpy(cur).cuda()).cpu().numpy()
    small = make_model(sd, max_batch=128)          # 300 samples = 3 passes (128 + 128 + 44)
    got = small.sample(torch.from_numpy(hist).cuda(), torch.from_numpy(cur).cuda()).cpu().numpy()
    assert np.array_equal(got, ref)                # rows are independent: the tiling must not change a bit
    host = small.sample(hist, cur)                 # numpy in -> host-buffer entry point -> numpy out
    assert isinstance(host, np.ndarray) and np.array_equal(host, ref)
    again = big.sample(torch.from_numpy(hist).cuda(), torch.from_numpy(cur).cuda()).cpu().numpy()
    assert np.array_equal(again, ref)              # deterministic


def test_other_windows_and_depths():
    """his_window 8 (distilled to 4 memory tokens), 10 prediction steps, 1 encoder / 3 decoder layers."""
    sd = mo.seeded_mtio_state_dict(23, bias=True, n_enc=1, n_dec=3)
    hist, cur = mo.synthetic_history(40, 35, his_window=8)
    want = mo.sample(sd, hist, cur, 10)
    net = make_model(sd, fut=10, num_encoder_layers=1, num_decoder_layers=3, his_window=8)
    net.fp32 = True
    np.testing.assert_allclose(net.sample(torch.from_numpy(hist).cuda(), torch.from_numpy(cur).cuda()).cpu().numpy(), want,
                               rtol=0, atol=FP32_ATOL)
    net.fp32 = False
    np.testing.assert_allclose(net.sample(torch.from_numpy(hist).cuda(), torch.from_numpy(cur).cuda()).cpu().numpy(), want,
                               rtol=0, atol=TF32_ATOL)


def test_predicted_masks_feed_the_simulator_tables():
    """predict.py:33-48 on the device: masks / IoU of the first 5 predicted points are bit-exact functions of the
    positions the kernels produced (a13-a16), and close positions give the oracle's masks."""
    from mansy_immersivevideostreaming_b200.config import SimConfig
    sd = mo.seeded_mtio_state_dict(24, bias=True)
    n = 200
    hist, cur = mo.synthetic_history(n, 36)
    rng = np.random.default_rng(1)
    gt_future = np.mod(cur + np.cumsum(rng.normal(0, 0.03, size=(n, 15, 2)), axis=1), 1.0).astype(np.float32)
    net = make_model(sd)
    gt_m, pred_m, acc, pred = net.predict_chunk_masks(torch.from_numpy(hist), torch.from_numpy(cur), torch.from_numpy(gt_future))
    p = pred.cpu().numpy()
    ogt, opred, oacc = so.chunk_masks(gt_future[:, :5], p[:, :5], SimConfig())
    assert np.array_equal(gt_m.cpu().numpy().view(np.uint64), ogt)
    assert np.array_equal(pred_m.cpu().numpy().view(np.uint64), opred)
    assert np.array_equal(acc.cpu().numpy(), oacc)
    # against the oracle's own predictions: identical masks wherever no point sits within tolerance of a tile edge
    want = mo.sample(sd, hist, cur, 15)
    _, opred2, _ = so.chunk_masks(gt_future[:, :5], want[:, :5], SimConfig())
    same = (opred2 == opred).mean()
    print(f"predicted masks equal to the fp32 oracle's for {same * 100:.1f}% of samples")
    assert same >= 0.9


def test_kernel_timing_hook():
    sd = mo.seeded_mtio_state_dict(25, bias=False)
    hist, cur = mo.synthetic_history(256, 37)
    net = make_model(sd)
    net.sample(torch.from_numpy(hist).cuda(), torch.from_numpy(cur).cuda(), timed=True)
    torch.cuda.synchronize()
    ms, cnt = net.kernel_ms()
    # 2 enc layers x 4 + conv + 2 memory kv + 15 steps x 2 layers x 6 GEMMs; attention 2 + 15 x 2 x 2
    assert cnt[0] == 2 * 4 + 1 + 2 + 15 * 2 * 6 and cnt[1] == 2 + 15 * 2 * 2 and cnt[2] == 5 + 15
    assert (ms > 0).all()


def test_early_stop_equals_prefix_of_full_run():
    """n_steps < fut_window: the first steps are bit-identical to a full run (a prediction never depends on later steps)."""
    sd = mo.seeded_mtio_state_dict(26, bias=True)
    hist, cur = mo.synthetic_history(200, 38)
    net = make_model(sd)
    h, c = torch.from_numpy(hist).cuda(), torch.from_numpy(cur).cuda()
    full = net.sample(h, c).cpu().numpy()
    short = net.sample(h, c, steps=5).cpu().numpy()
    assert np.array_equal(short[:, :5], full[:, :5]) and not short[:, 5:].any()
    assert np.array_equal(net.sample(hist, cur), full)       # host path, all steps: fills the library's staging buffers
    host = net.sample(hist, cur, steps=5)
    assert np.array_equal(host, short)                       # rows >= 5 of the caller's (zeroed) buffer stay untouched
    _, tok_full = net.sample(h, c, return_tokens=True)
    _, tok_short = net.sample(h, c, return_tokens=True, steps=5)
    tok_full, tok_short = tok_full.cpu().numpy(), tok_short.cpu().numpy()
    assert np.array_equal(tok_short[:, :6], tok_full[:, :6]) and not tok_short[:, 6:].any()


def test_config5_predicted_tables_drive_the_environments():
    """BASELINE config 5 end to end at test size: MTIO predictions -> tile masks -> the simulator's tables -> lock-step
    MANSY envs; the CUDA envs on those tables equal the oracle envs on the same tables."""
    from mansy_immersivevideostreaming_b200 import synth
    from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, REWARD_QOE, SimConfig
    from mansy_immersivevideostreaming_b200.mtio import MtioMaskFn, viewport_windows
    from mansy_immersivevideostreaming_b200.simulator import BatchSimulator
    sd = mo.seeded_mtio_state_dict(27, bias=True)
    net = make_model(sd)
    fn = MtioMaskFn(net, n_vp_chunks=54)
    tables, gt_xy, _ = synth.make_synthetic_tables(fn, n_videos=2, n_users=3, n_traces=4, seed=5, trace_len_range=(40, 90),
                                                   return_centres=True)
    # the masks in the tables are the oracle's masks of the positions the kernels predicted
    hist, cur = viewport_windows(gt_xy)
    assert hist.shape == (6 * 54, 5, 2) and np.array_equal(cur[1, 0], gt_xy[0, 0, 4]) and np.array_equal(hist[2, 4], gt_xy[0, 1, 3])
    ogt, opred, oacc = so.chunk_masks(gt_xy.reshape(-1, 5, 2), fn.last_pred, SimConfig())
    valid = (np.arange(54)[None, :] <= (tables.vp_end - tables.vp_start)[:, None]).reshape(-1)
    assert np.array_equal(tables.vp_gt.reshape(-1)[valid], ogt[valid])
    assert np.array_equal(tables.vp_pred.reshape(-1)[valid], opred[valid])
    assert np.array_equal(tables.vp_acc.reshape(-1)[valid], oacc[valid])
    want = mo.sample(sd, hist, cur, 5)
    np.testing.assert_allclose(fn.last_pred, want, rtol=0, atol=TF32_ATOL)
    # and the environments run on them
    n = 32
    tables = tables.with_samples(synth.per_env_samples(tables, n))
    sim = BatchSimulator(tables, n, OBS_MODE_MANSY, REWARD_QOE, seed=1)
    orc = so.OracleVectorEnv(tables, n, OBS_MODE_MANSY, REWARD_QOE, chain="f64", seed=1)
    assert np.array_equal(sim.reset().cpu().numpy(), orc.reset())
    for t in range(55):
        acts = synth.synthetic_actions(n, t, seed=3)
        obs, rew, done = sim.step(torch.from_numpy(acts), auto_reset=True)
        oobs, orew, odone, _ = orc.step(acts, auto_reset=True)
        np.testing.assert_allclose(obs.cpu().numpy(), oobs, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(rew.cpu().numpy(), orew, rtol=1e-5, atol=1e-6)
        assert np.array_equal(done.cpu().numpy().astype(bool), odone)


def test_linear_regression_predictor():
    """predict.py --model regression: the least-squares extrapolation kernel vs the reference golden and the oracle."""
    from mansy_immersivevideostreaming_b200.mtio import LinearRegression
    g = load_golden("linreg_kat.npz")
    net = LinearRegression(fut_window=15, device="cuda:0")
    got = net.sample(torch.from_numpy(g["history"]), torch.from_numpy(g["current"])).cpu().numpy()
    np.testing.assert_allclose(got, g["pred"], rtol=0, atol=5e-7)
    hist, cur = mo.synthetic_history(5000, 77, his_window=7)
    got = LinearRegression(9, "cuda:0").sample(hist, cur).cpu().numpy()
    np.testing.assert_allclose(got, mo.linreg_sample(hist, cur, 9), rtol=0, atol=5e-7)
    assert got.shape == (5000, 9, 2)


def test_two_lane_pass_equals_single_lane(monkeypatch):
    """>= 1024 samples run as two halves on two streams (attention of one half under the GEMMs of the other); rows are
    independent, so the result equals the single-stream pass bit for bit, and a sub-sample equals the oracle."""
    sd = mo.seeded_mtio_state_dict(28, bias=True)
    hist, cur = mo.synthetic_history(1300, 39)
    h, c = torch.from_numpy(hist).cuda(), torch.from_numpy(cur).cuda()
    monkeypatch.setenv("MANSY_MTIO_LANES", "1")
    one = make_model(sd).sample(h, c, return_tokens=True)
    monkeypatch.setenv("MANSY_MTIO_LANES", "2")
    net2 = make_model(sd)
    two = net2.sample(h, c, return_tokens=True)
    assert torch.equal(one[0], two[0]) and torch.equal(one[1], two[1])
    host = net2.sample(hist, cur)
    assert np.array_equal(host, two[0].cpu().numpy())
    idx = np.r_[0:8, 760:776, 1292:1300]              # both lanes, around the cut at row 768
    np.testing.assert_allclose(two[0].cpu().numpy()[idx], mo.sample(sd, hist[idx], cur[idx], 15), rtol=0, atol=TF32_ATOL)


def test_full_size_properties():
    """BASELINE configs[4] size (16,384 samples): results do not depend on how the batch is tiled (one two-lane pass vs
    four sequential passes of 4,096), stay inside the unit square, and a spread of rows equals the oracle."""
    sd = mo.seeded_mtio_state_dict(29, bias=True)
    n = 16384
    hist, cur = mo.synthetic_history(n, 40)
    h, c = torch.from_numpy(hist).cuda(), torch.from_numpy(cur).cuda()
    whole = make_model(sd, max_batch=n).sample(h, c)
    parts = make_model(sd, max_batch=4096).sample(h, c)
    assert torch.equal(whole, parts)
    w = whole.cpu().numpy()
    assert w.shape == (n, 15, 2) and np.isfinite(w).all() and w.min() >= 0.0 and w.max() <= 1.0
    idx = np.arange(0, n, 683)                       # 24 rows across tiles, lanes and passes
    np.testing.assert_allclose(w[idx], mo.sample(sd, hist[idx], cur[idx], 15), rtol=0, atol=TF32_ATOL)


def test_run_to_run_determinism():
    """Twelve passes over 4,096 samples (two lanes, split LayerNorm tiles, residual buffers handed back to TMA) are
    bit-identical; tools/mtio_stress.py runs the long version (400 passes of 16,384)."""
    sd = mo.seeded_mtio_state_dict(30, bias=True)
    hist, cur = mo.synthetic_history(4096, 41)
    h, c = torch.from_numpy(hist).cuda(), torch.from_numpy(cur).cuda()
    net = make_model(sd)
    ref = net.sample(h, c)
    for _ in range(12):
        assert torch.equal(net.sample(h, c), ref)
