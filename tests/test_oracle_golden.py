"""The oracle restatement against the committed golden vectors (which were produced by running
the unmodified reference, oracle/make_golden.py) -- CPU only."""
import numpy as np
import pytest

from helpers import RTOL, assert_rows_match, golden_tables, load_golden, reward_scale
from mansy_immersivevideostreaming_b200.config import (OBS_MODE_MANSY, OBS_MODE_SIMPLE, REWARD_QOE,
                                                       REWARD_QOE_NORM, SimConfig)
from mansy_immersivevideostreaming_b200.tables import environment_test_samples
from oracle import sim_oracle as so

CFG = SimConfig()


def test_geometry_grid_kat():
    g = load_golden("geometry_kat.npz")
    for x, y, m in zip(g["x"], g["y"], g["mask"]):
        assert so.mask_bits(so.fov_tile_mask(int(x), int(y), CFG)) == int(m)


def test_geometry_shipped_gt_masks():
    """Shipped ground-truth masks of the reference dataset, recomputed from the shipped 5 Hz centres."""
    g = load_golden("geometry_kat.npz")
    for chain in ("f64", "f32"):
        gt, _, _ = so.chunk_masks(g["ship_xy"], g["ship_xy"], CFG, chain=chain)
        assert np.array_equal(gt, g["ship_gt"])


def test_chunk_masks_and_iou():
    g = load_golden("geometry_kat.npz")
    gt, pred, acc = so.chunk_masks(g["cm_gt_xy"], g["cm_pred_xy"], CFG)
    assert np.array_equal(gt, g["cm_gt"]) and np.array_equal(pred, g["cm_pred"]) and np.array_equal(acc, g["cm_acc"])


def test_geometry_domain_errors():
    with pytest.raises(ValueError):
        so.fov_tile_mask(-1, 10, CFG)
    with pytest.raises(ValueError):
        so.fov_tile_mask(10, 1441, CFG)


def test_allocate_kat():
    g = load_golden("allocate_kat.npz")
    for m, vers in zip(g["mask"], g["versions"]):
        for a in range(16):
            rin, rout = so.action_to_rates(a)
            assert np.array_equal(so.allocate_tile_versions(rin, rout, int(m), CFG.video_rates), vers[a])


def test_allocate_is_toroidal_chebyshev():
    """BFS level == toroidal Chebyshev distance (property used by the kernels' dilation form)."""
    rng = np.random.default_rng(0)
    for _ in range(50):
        bits = so.mask_bits((rng.random(64) < rng.uniform(0.02, 0.5)).astype(np.uint8)) or 1
        scale = so.tile_scales(bits)
        src = [(t // 8, t % 8) for t in range(64) if (bits >> t) & 1]
        for t in range(64):
            r, c = t // 8, t % 8
            d = min(max(min(abs(r - sr), 8 - abs(r - sr)), min(abs(c - sc), 8 - abs(c - sc))) for sr, sc in src)
            assert scale[t] == d


def test_trace_and_buffer_kat():
    g = load_golden("trace_kat.npz")
    for thr, L, sizes, rec in zip(g["thr"], g["lens"], g["sizes"], g["rec"]):
        idx, tm, buf = 0, 0.0, 3.0
        for s, (dl, ridx, rtm, rb, rbuf) in zip(sizes, rec):
            odl, idx, tm = so.trace_download(int(s), thr, int(L), idx, tm)
            orb, buf = so.buffer_push(buf, 1, odl)
            assert (odl, idx, tm, orb, buf) == (dl, int(ridx), rtm, rb, rbuf)


def _replay(g, tag, tables, obs_mode, reward_mode, chain):
    wid, wnum = (int(x) for x in g[f"{tag}_worker"]) if f"{tag}_worker" in g else (1, 2)
    env = so.OracleEnv(tables, obs_mode, reward_mode, chain, worker_id=wid, worker_num=wnum)
    obs, rew, done, act, aux, vers = (g[f"{tag}_{k}"] for k in ("obs", "reward", "done", "action", "aux", "versions"))
    n_steps = 0
    for i in range(obs.shape[0]):
        if act[i] < 0:
            row = so.flatten_obs(env.reset(), obs_mode)
            assert_rows_match(row, obs[i], obs_mode, chain_exact=True, ctx=f"{tag} reset row {i}")
            continue
        o, r, d, a = env.step(int(act[i]))
        row = so.flatten_obs(o, obs_mode)
        assert d == bool(done[i])
        assert np.array_equal(a["versions"], vers[i])                       # chosen-tile indices: bit-exact
        assert a["chunk_size"] == aux[i, 0] and a["cur_idx"] == aux[i, 4] and a["next_chunk"] == aux[i, 9]
        # download time / buffer / rebuffer: pure float64 in every chain -> exact
        assert (a["download_time"], a["rebuffer"], a["buffer"], a["cur_time"]) == tuple(aux[i, [1, 2, 3, 5]])
        assert_rows_match(row, obs[i], obs_mode, chain_exact=(chain == "f32"), ctx=f"{tag} row {i}")
        if chain == "f32":
            assert float(r) == rew[i]
        else:
            scale = reward_scale(env.w, a["qoe1"], a["qoe2"], a["qoe3"], reward_mode == REWARD_QOE_NORM)
            assert abs(float(r) - rew[i]) <= RTOL * scale
            for k, col in (("qoe1", 6), ("qoe2", 7), ("qoe3", 8)):
                assert abs(a[k] - aux[i, col]) <= RTOL * max(abs(aux[i, col]), 1e-2)
        n_steps += 1
    # episode rows (video, user, trace, weights, steps, sample id)
    eps = g[f"{tag}_episodes"]
    assert len(env.episodes) == eps.shape[0]
    for e, row in zip(env.episodes, eps):
        assert [e["video"], e["user"], e["trace"], *e["w"], e["steps"], e["sample_id"]] == list(row)
    return n_steps


@pytest.mark.parametrize("chain", ["f32", "f64"])
def test_mansy_synth_golden(chain):
    g = load_golden("mansy_synth.npz")
    tables = golden_tables(g)
    assert _replay(g, "train", tables, OBS_MODE_MANSY, REWARD_QOE, chain) > 200
    assert _replay(g, "norm", tables, OBS_MODE_MANSY, REWARD_QOE_NORM, chain) > 200


@pytest.mark.parametrize("chain", ["f32", "f64"])
def test_mansy_real_golden(chain):
    g = load_golden("mansy_real.npz")
    assert _replay(g, "test", golden_tables(g), OBS_MODE_MANSY, REWARD_QOE, chain) > 250


@pytest.mark.parametrize("chain", ["f32", "f64"])
def test_simple_synth_golden(chain):
    g = load_golden("simple_synth.npz")
    tables = golden_tables(g)
    assert _replay(g, "train", tables, OBS_MODE_SIMPLE, REWARD_QOE_NORM, chain) > 150
    test_tables = tables.with_samples(environment_test_samples(tables.n_videos, tables.n_users, tables.n_traces,
                                                               tables.qoe_w.shape[0]))
    assert _replay(g, "test", test_tables, OBS_MODE_SIMPLE, REWARD_QOE, chain) > 150


def test_log_rows_match_reference_csv():
    """`_log` rows (mansy_env.py:271-290): ids, weights and 5-decimal means."""
    g = load_golden("mansy_real.npz")
    tables = golden_tables(g)
    env = so.OracleEnv(tables, OBS_MODE_MANSY, REWARD_QOE, "f64", worker_id=1, worker_num=7)
    for i, a in enumerate(g["test_action"]):
        env.reset() if a < 0 else env.step(int(a))
    lines = str(g["test_log"]).strip().splitlines()[1:]
    assert len(lines) == len(env.episodes)
    for line, e in zip(lines, env.episodes):
        f = line.split(",")
        assert [int(f[0]), int(f[1]), int(f[2])] == [e["video"], e["user"], e["trace"]]
        assert [float(x) for x in f[3:6]] == list(e["w"])
        for val, key in zip(f[6:], ("qoe", "qoe1", "qoe2", "qoe3")):
            assert abs(float(val) - e[key]) <= 2e-5      # 5-decimal rounding of float32- vs float64-chain means


def test_gae_oracle_known_answers():
    """The GAE restatement (tianshou 0.4.8 algorithm; parity unpinned) on hand-computed cases."""
    from oracle import gae_oracle
    rew = np.array([[1.0], [2.0], [3.0]], dtype=np.float32)
    val = np.array([[0.5], [0.25], [0.125]], dtype=np.float32)
    done = np.zeros((3, 1), dtype=np.uint8)
    adv, ret = gae_oracle.gae(rew, val, done, np.array([1.0], np.float32), 0.5, 1.0)
    # returns with lambda = 1: 3 + .5*1 = 3.5; 2 + .5*3.5 = 3.75; 1 + .5*3.75 = 2.875
    np.testing.assert_allclose(ret[:, 0], [2.875, 3.75, 3.5])
    np.testing.assert_allclose(adv[:, 0], [2.875 - 0.5, 3.75 - 0.25, 3.5 - 0.125])
    done[1, 0] = 1        # the episode ends after step 1: no bootstrap across it
    adv, ret = gae_oracle.gae(rew, val, done, np.array([1.0], np.float32), 0.5, 1.0)
    np.testing.assert_allclose(ret[:, 0], [1 + 0.5 * 2.0, 2.0, 3.5])
    ident, mixed = gae_oracle.identifier_reward(np.array([[0.5, 0.25, 0.25]], np.float32), np.array([[0.25, 0.25, 0.5]], np.float32),
                                                np.array([2.0]), 0.25)
    np.testing.assert_allclose(ident, [1 - (0.0625 + 0 + 0.0625) / 3], rtol=1e-7)
    np.testing.assert_allclose(mixed, [0.75 * 2.0 + 0.25 * float(ident[0])])
