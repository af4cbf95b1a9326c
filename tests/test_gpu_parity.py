"""GPU parity: the CUDA path (through the C ABI) against the committed goldens and the oracle.

Bars (BASELINE.json north_star): tile masks, chosen-tile indices, chunk/step counters bit-exact;
download time, buffer level, QoE reward within 1e-5 relative.  The oracle's float64 chain is the
reference under its pinned numpy; the goldens came from the reference under numpy 2 (float32
chain), so rewards are compared with the cancellation-aware scale of tests/helpers.py.
"""
import numpy as np
import pytest
import torch

from helpers import RTOL, assert_rows_match, golden_tables, load_golden, rel_err, reward_scale
from mansy_immersivevideostreaming_b200 import synth
from mansy_immersivevideostreaming_b200.config import (OBS_MODE_MANSY, OBS_MODE_NONE, OBS_MODE_SIMPLE, REWARD_QOE,
                                                       REWARD_QOE_NORM, SimConfig)
from mansy_immersivevideostreaming_b200.simulator import BatchSimulator, ViewportTiler
from mansy_immersivevideostreaming_b200.tables import environment_test_samples
from oracle import sim_oracle as so

pytestmark = pytest.mark.gpu
CFG = SimConfig()

AUX = dict(chunk_size=0, download_time=1, rebuffer=2, buffer=3, cur_idx=4, cur_time=5, qoe=6, qoe1=7, qoe2=8, qoe3=9,
           next_chunk=10, ep_step=11, sample_id=12, reward=13)


# ---------------------------------------------------------------------------------------------
# a13-a16 geometry, a1-a2 allocation
# ---------------------------------------------------------------------------------------------
def test_viewport_tiles_shipped_and_synthetic():
    g = load_golden("geometry_kat.npz")
    tiler = ViewportTiler(CFG)
    gt, _, _ = tiler.chunk_masks(g["ship_xy"], g["ship_xy"])
    assert np.array_equal(gt, g["ship_gt"])                     # shipped ground-truth masks, bit-exact
    gt, pred, acc = tiler.chunk_masks(g["cm_gt_xy"], g["cm_pred_xy"])
    assert np.array_equal(gt, g["cm_gt"]) and np.array_equal(pred, g["cm_pred"])
    assert np.array_equal(acc, g["cm_acc"])                     # IoU: one correctly rounded float64 division


def test_viewport_tiles_grid_kat_and_large():
    g = load_golden("geometry_kat.npz")
    tiler = ViewportTiler(CFG)
    # exact pixel centres: x/2560 and y/1440 are not all representable, so feed centres whose
    # truncation gives the KAT pixel (v = (p + 0.5) / L except at the frame edge)
    xs, ys = g["x"].astype(np.float64), g["y"].astype(np.float64)
    vx = np.where(xs >= 2560, 1.0, (xs + 0.5) / 2560).astype(np.float32)
    vy = np.where(ys >= 1440, 1.0, (ys + 0.5) / 1440).astype(np.float32)
    assert np.array_equal((vx.astype(np.float64) * 2560).astype(np.int64), g["x"])
    assert np.array_equal((vy.astype(np.float64) * 1440).astype(np.int64), g["y"])
    xy = np.stack([vx, vy], axis=1)[:, None, :]
    gt, _, _ = tiler.chunk_masks(xy, xy)
    assert np.array_equal(gt, g["mask"])
    # 200k random chunks against the oracle on a sample, plus structural properties on all
    rng = np.random.default_rng(5)
    gxy = rng.random((200_000, 5, 2)).astype(np.float32)
    pxy = np.mod(gxy + rng.normal(0, 0.05, gxy.shape), 1.0).astype(np.float32)
    gt, pred, acc = tiler.chunk_masks(gxy, pxy)
    idx = rng.choice(gxy.shape[0], 300, replace=False)
    ogt, opred, oacc = so.chunk_masks(gxy[idx], pxy[idx], CFG)
    assert np.array_equal(gt[idx], ogt) and np.array_equal(pred[idx], opred) and np.array_equal(acc[idx], oacc)
    pc = np.array([bin(int(m)).count("1") for m in gt[:5000]])
    assert pc.min() >= 4 and pc.max() <= 64 and np.all((acc >= 0) & (acc <= 1))


def test_allocate_tile_versions_kat():
    g = load_golden("allocate_kat.npz")
    tiler = ViewportTiler(CFG)
    masks = np.repeat(g["mask"], 16).view(np.int64)
    actions = np.tile(np.arange(16, dtype=np.int32), g["mask"].shape[0])
    out = tiler.allocate_tile_versions(torch.from_numpy(masks.copy()), torch.from_numpy(actions)).cpu().numpy()
    assert np.array_equal(out.reshape(-1, 16, 64), g["versions"])


# ---------------------------------------------------------------------------------------------
# golden episodes (reference-run vectors), N = 1 through the same API a gym env uses
# ---------------------------------------------------------------------------------------------
def _replay_golden(g, tag, tables, obs_mode, reward_mode, table=False):
    """table=False: every step gathers (and returns the per-tile versions); table=True: the (pair, chunk, action)
    outcome table built at create answers the download / QoE parts -- both against the same reference-run vectors."""
    wid, wnum = (int(x) for x in g[f"{tag}_worker"]) if f"{tag}_worker" in g else (1, 2)
    sim = BatchSimulator(tables, 1, obs_mode, reward_mode, seed=wid, worker_num=wnum)
    sim.set_outcome_table(table)
    obs, rew, done, act, aux, vers = (g[f"{tag}_{k}"] for k in ("obs", "reward", "done", "action", "aux", "versions"))
    aux_d, ver_d = sim.new_aux(), sim.new_versions()
    max_rel = 0.0
    for i in range(obs.shape[0]):
        if act[i] < 0:
            row = sim.reset().cpu().numpy()[0]
            assert_rows_match(row, obs[i], obs_mode, chain_exact=True, ctx=f"{tag} reset row {i}")
            continue
        o, r, d = sim.step(torch.tensor([int(act[i])], dtype=torch.int32), aux=aux_d, versions=None if table else ver_d)
        row, a = o.cpu().numpy()[0], aux_d.cpu().numpy()[0]
        assert bool(d.item()) == bool(done[i])
        assert table or np.array_equal(ver_d.cpu().numpy()[0], vers[i])             # bit-exact
        assert a[AUX["chunk_size"]] == aux[i, 0] and a[AUX["cur_idx"]] == aux[i, 4] and a[AUX["next_chunk"]] == aux[i, 9]
        # float64 trace/buffer arithmetic in the reference's operation order: bit-exact, far inside 1e-5
        assert (a[AUX["download_time"]], a[AUX["rebuffer"]], a[AUX["buffer"]], a[AUX["cur_time"]]) == tuple(aux[i, [1, 2, 3, 5]])
        assert_rows_match(row, obs[i], obs_mode, chain_exact=False, ctx=f"{tag} row {i}")
        w = tables.qoe_w[tables.samples[int(a[AUX["sample_id"]])][3]]
        scale = reward_scale(w, a[AUX["qoe1"]], a[AUX["qoe2"]], a[AUX["qoe3"]], reward_mode == REWARD_QOE_NORM)
        assert abs(float(r.item()) - rew[i]) <= RTOL * scale
        max_rel = max(max_rel, abs(float(r.item()) - rew[i]) / scale)
        for key, col in (("qoe1", 6), ("qoe2", 7), ("qoe3", 8)):
            assert abs(a[AUX[key]] - aux[i, col]) <= RTOL * max(abs(aux[i, col]), 1e-2)
    sim.close()
    return max_rel


def test_golden_mansy_synth():
    g = load_golden("mansy_synth.npz")
    tables = golden_tables(g)
    _replay_golden(g, "train", tables, OBS_MODE_MANSY, REWARD_QOE)
    _replay_golden(g, "norm", tables, OBS_MODE_MANSY, REWARD_QOE_NORM)
    _replay_golden(g, "train", tables, OBS_MODE_MANSY, REWARD_QOE, table=True)
    _replay_golden(g, "norm", tables, OBS_MODE_MANSY, REWARD_QOE_NORM, table=True)


def test_golden_mansy_real_data():
    g = load_golden("mansy_real.npz")
    _replay_golden(g, "test", golden_tables(g), OBS_MODE_MANSY, REWARD_QOE)
    _replay_golden(g, "test", golden_tables(g), OBS_MODE_MANSY, REWARD_QOE, table=True)


def test_golden_simple_synth():
    g = load_golden("simple_synth.npz")
    tables = golden_tables(g)
    _replay_golden(g, "train", tables, OBS_MODE_SIMPLE, REWARD_QOE_NORM)
    tt = tables.with_samples(environment_test_samples(tables.n_videos, tables.n_users, tables.n_traces, tables.qoe_w.shape[0]))
    _replay_golden(g, "test", tt, OBS_MODE_SIMPLE, REWARD_QOE)
    _replay_golden(g, "test", tt, OBS_MODE_SIMPLE, REWARD_QOE, table=True)


# ---------------------------------------------------------------------------------------------
# many envs in lockstep against the float64 oracle (same chain -> bit-exact or ~1e-16)
# ---------------------------------------------------------------------------------------------
def _oracle_tables(n_videos=4, n_users=5, n_traces=6, seed=11, qoe=None, n_samples=256):
    """Small synthetic dataset (masks from the oracle) with one random sample tuple per env slot: the
    reference indexes ``samples[worker_id]`` directly (mansy_env.py:100,103), so a vector env of N
    envs needs at least N samples or the reference -- and the oracle -- raise IndexError."""
    t = synth.make_synthetic_tables(lambda g, p: so.chunk_masks(g, p, CFG), n_videos=n_videos, n_users=n_users,
                                    n_traces=n_traces, seed=seed, trace_len_range=(30, 120), short_tail_frac=0.25,
                                    qoe_w=qoe)
    return t.with_samples(synth.per_env_samples(t, n_samples, seed=seed + 1))


@pytest.mark.parametrize("obs_mode,reward_mode", [(OBS_MODE_MANSY, REWARD_QOE), (OBS_MODE_SIMPLE, REWARD_QOE_NORM)])
def test_lockstep_vs_oracle_autoreset(obs_mode, reward_mode):
    qoe = np.concatenate([np.asarray(synth.DEFAULT_QOE, np.float32), synth.diverse_qoe_weights(5)])
    tables = _oracle_tables(qoe=qoe)
    N, T = 96, 70                         # > one episode per env: exercises in-kernel auto-reset
    sim = BatchSimulator(tables, N, obs_mode, reward_mode, seed=3)
    orc = so.OracleVectorEnv(tables, N, obs_mode, reward_mode, chain="f64", seed=3)
    obs = sim.reset().cpu().numpy()
    oobs = orc.reset()
    assert np.array_equal(obs, oobs)
    aux_d, ver_d = sim.new_aux(), sim.new_versions()
    worst = dict(reward=0.0, download_time=0.0, buffer=0.0)
    for t in range(T):
        acts = synth.synthetic_actions(N, t, seed=99)
        if t == 5:
            acts[:4] = [15, -1, 99, 14]   # out-of-table actions behave like (0, 0)
        o, r, d = sim.step(torch.from_numpy(acts), auto_reset=True, aux=aux_d, versions=ver_d)
        oo, orr, od, oaux = orc.step(acts, auto_reset=True)
        a = aux_d.cpu().numpy()
        assert np.array_equal(d.cpu().numpy().astype(bool), od)
        assert np.array_equal(ver_d.cpu().numpy(), np.stack([x["versions"] for x in oaux]).astype(np.uint8))
        for key in ("chunk_size", "cur_idx", "next_chunk", "sample_id"):
            assert np.array_equal(a[:, AUX[key]], np.array([x[key] for x in oaux], dtype=np.float64)), key
        gt_bits = a[:, 14].astype(np.uint64) | (a[:, 15].astype(np.uint64) << np.uint64(32))
        assert np.array_equal(gt_bits, np.array([x["gt_bits"] for x in oaux], dtype=np.uint64))   # tile masks bit-exact
        for key in ("download_time", "rebuffer", "buffer", "cur_time", "qoe", "qoe1", "qoe2", "qoe3"):
            ref = np.array([x[key] for x in oaux])
            assert np.all(rel_err(a[:, AUX[key]], ref) <= RTOL), key
            if key in worst:
                worst[key] = max(worst[key], float(rel_err(a[:, AUX[key]], ref).max()))
        assert np.all(rel_err(a[:, AUX["reward"]], orr) <= RTOL)
        worst["reward"] = max(worst["reward"], float(rel_err(a[:, AUX["reward"]], orr).max()))
        np.testing.assert_allclose(r.cpu().numpy(), orr.astype(np.float32), rtol=1e-6, atol=0)   # float32 rounding of the same value
        got = o.cpu().numpy()
        for i in range(N):
            assert_rows_match(got[i], oo[i], obs_mode, chain_exact=False, ctx=f"t={t} env={i}")
    assert sim.error_flag() == 0
    # per-episode records (what `_log` writes) for every env's last finished episode
    stats = sim.episode_stats().cpu().numpy()
    for i, env in enumerate(orc.envs):
        e = env.episodes[-1]
        assert stats[i, 4] == e["steps"] and stats[i, 5] == e["sample_id"] and stats[i, 11] == len(env.episodes)
        np.testing.assert_allclose(stats[i, 0:4], e["sums"], rtol=1e-12)
    print("worst relative errors vs float64 oracle:", worst)
    sim.close()


def test_step_by_ids_and_explicit_reset():
    """tianshou protocol: step a subset by id, terminal observation returned, caller resets ids."""
    tables = _oracle_tables(seed=21)
    N = 12
    sim = BatchSimulator(tables, N, OBS_MODE_MANSY, REWARD_QOE, seed=0)
    orc = so.OracleVectorEnv(tables, N, OBS_MODE_MANSY, REWARD_QOE, chain="f64", seed=0)
    sim.reset(); orc.reset()
    rng = np.random.default_rng(2)
    for t in range(130):
        ids = np.sort(rng.choice(N, size=int(rng.integers(1, N + 1)), replace=False)).astype(np.int32)
        acts = rng.integers(0, 15, size=ids.shape[0]).astype(np.int32)
        o, r, d = sim.step(torch.from_numpy(acts), env_ids=ids)
        o, d = o.cpu().numpy(), d.cpu().numpy().astype(bool)
        done_ids = []
        for j, (e, a) in enumerate(zip(ids, acts)):
            oo, orr, od, _ = orc.envs[e].step(int(a))
            assert od == d[j]
            assert_rows_match(o[j], so.flatten_obs(oo, OBS_MODE_MANSY), OBS_MODE_MANSY, chain_exact=False, ctx=f"t={t} env={e}")
            if od:
                done_ids.append(int(e))
        if done_ids:
            rows = sim.reset(done_ids).cpu().numpy()
            for j, e in enumerate(done_ids):
                assert np.array_equal(rows[j], so.flatten_obs(orc.envs[e].reset(), OBS_MODE_MANSY))
    sim.close()


def test_shard_invariance():
    """N envs on one handle == the same envs split over shards (env_offset / global worker_num)."""
    tables = _oracle_tables(seed=31)
    N, T = 64, 60
    full = BatchSimulator(tables, N, OBS_MODE_MANSY, REWARD_QOE, seed=5, worker_num=N)
    shards = [BatchSimulator(tables, N // 4, OBS_MODE_MANSY, REWARD_QOE, seed=5, worker_num=N, env_offset=k * (N // 4))
              for k in range(4)]
    a = full.reset()
    b = torch.cat([s.reset() for s in shards])
    assert torch.equal(a, b)
    for t in range(T):
        acts = torch.from_numpy(synth.synthetic_actions(N, t, seed=7))
        o, r, d = full.step(acts, auto_reset=True)
        parts = [s.step(acts[k * 16:(k + 1) * 16], auto_reset=True) for k, s in enumerate(shards)]
        assert torch.equal(o, torch.cat([p[0] for p in parts]))
        assert torch.equal(r, torch.cat([p[1] for p in parts])) and torch.equal(d, torch.cat([p[2] for p in parts]))
    assert torch.equal(full.episode_stats(), torch.cat([s.episode_stats() for s in shards]))


@pytest.mark.parametrize("obs_mode", [OBS_MODE_MANSY, OBS_MODE_SIMPLE, OBS_MODE_NONE])
def test_multistep_rollout_equals_single_steps(obs_mode):
    """The persistent multi-step launch (state in registers) == T single-step launches with the
    same hashed action stream; also covers env_offset in the action hash."""
    tables = _oracle_tables(seed=41)
    N, T = 80, 75
    a = BatchSimulator(tables, N, obs_mode, REWARD_QOE, seed=1, env_offset=1000, worker_num=5000)
    b = BatchSimulator(tables, N, obs_mode, REWARD_QOE, seed=1, env_offset=1000, worker_num=5000)
    a.reset(); b.reset()
    stride = max(a.obs_stride, 4)
    obs_a = torch.zeros((T, N, stride), dtype=torch.float32, device="cuda") if obs_mode != OBS_MODE_NONE else None
    rew_a = torch.zeros((T, N), dtype=torch.float32, device="cuda")
    done_a = torch.zeros((T, N), dtype=torch.uint8, device="cuda")
    a.rollout_random(T, seed=1234, step0=10, obs=obs_a, reward=rew_a, done=done_a, per_step_outputs=True)
    for t in range(T):
        acts = torch.from_numpy(synth.synthetic_actions(N, 10 + t, seed=1234, env_offset=1000))
        o, r, d = b.step(acts, auto_reset=True)
        if obs_mode != OBS_MODE_NONE:
            assert torch.equal(obs_a[t], o), t
        assert torch.equal(rew_a[t], r) and torch.equal(done_a[t], d), t
    assert torch.equal(a.episode_stats(), b.episode_stats())
    sa, sb = a.episode_state_host(), b.episode_state_host()
    for f in ("cur_time", "buf", "next_chunk", "cur_idx", "ep_step", "cursor", "sample_id"):
        assert np.array_equal(sa[f], sb[f]), f


def test_host_buffer_path_equals_device_path():
    tables = _oracle_tables(seed=51)
    N = 40
    a = BatchSimulator(tables, N, OBS_MODE_MANSY, REWARD_QOE, seed=0)
    b = BatchSimulator(tables, N, OBS_MODE_MANSY, REWARD_QOE, seed=0)
    host = a.make_host_buffers()
    a.reset_host(host)
    assert np.array_equal(host["obs"].numpy(), b.reset().cpu().numpy())
    for t in range(60):
        acts = synth.synthetic_actions(N, t, seed=5)
        host["actions"].copy_(torch.from_numpy(acts))
        a.step_host(host, auto_reset=True)
        o, r, d = b.step(torch.from_numpy(acts), auto_reset=True)
        assert np.array_equal(host["obs"].numpy(), o.cpu().numpy())
        assert np.array_equal(host["reward"].numpy(), r.cpu().numpy()) and np.array_equal(host["done"].numpy(), d.cpu().numpy())


def test_full_size_properties():
    """BASELINE sizes (65,536 envs, diverse QoE weights): size-independent properties instead of
    the scalar oracle -- episode lengths equal the table-derived chunk counts, counters advance in
    lockstep, rewards recompute from the emitted QoE terms, masks/one-hots are valid, and a sample
    of envs is checked against the oracle."""
    cfg = SimConfig()
    tiler = ViewportTiler(cfg)
    qoe = synth.diverse_qoe_weights(65_536)
    base = synth.make_synthetic_tables(tiler.chunk_masks, n_videos=24, n_users=60, n_traces=40, qoe_w=qoe)
    N = 65_536
    tables = base.with_samples(np.concatenate([synth.per_env_samples(base, N)[:, :3], np.arange(N, dtype=np.int32)[:, None]], axis=1))
    sim = BatchSimulator(tables, N, OBS_MODE_MANSY, REWARD_QOE, seed=0)
    obs = sim.reset()
    aux = sim.new_aux()
    pair = tables.samples[:, 0] * tables.n_users + tables.samples[:, 1]
    end = np.minimum(tables.vp_end[pair], tables.video_time[tables.samples[:, 0]] - 1)
    ep_len = end - cfg.startup_download                      # chunks 6..end
    check = np.arange(0, N, 4099)
    orc = [so.OracleEnv(tables, OBS_MODE_MANSY, REWARD_QOE, "f64", worker_id=int(i), worker_num=N) for i in check]
    for o in orc:
        o.reset()
    steps_done = np.zeros(N, dtype=np.int64)
    for t in range(52):
        acts = synth.synthetic_actions(N, t, seed=77)
        o, r, d = sim.step(torch.from_numpy(acts), auto_reset=False, aux=aux)
        a = aux.cpu().numpy()
        live = steps_done < ep_len
        steps_done += live
        d = d.cpu().numpy().astype(bool)
        assert np.array_equal(d, steps_done >= ep_len)                       # episode ends exactly at its last chunk
        assert np.array_equal(a[live, AUX["ep_step"]], steps_done[live]) and np.array_equal(a[live, AUX["next_chunk"]], 6 + steps_done[live])
        w = tables.qoe_w[tables.samples[:, 3]].astype(np.float64)
        recomputed = w[:, 0] * a[:, AUX["qoe1"]] - w[:, 1] * a[:, AUX["qoe2"]] - w[:, 2] * a[:, AUX["qoe3"]]
        assert np.array_equal(recomputed[live], a[live, AUX["qoe"]])         # same float64 expression
        assert np.all(a[live, AUX["download_time"]] > 0) and np.all(a[live, AUX["buffer"]] >= 1.0)
        rows = o[:, 760:775].sum(dim=1)
        assert torch.all(rows[torch.from_numpy(live).cuda()] == 1.0)         # exactly one action bit
        for o_env, i in zip(orc, check):
            if not o_env.done:
                oo, orr, od, oa = o_env.step(int(acts[i]))
                assert_rows_match(o[i].cpu().numpy(), so.flatten_obs(oo, OBS_MODE_MANSY), OBS_MODE_MANSY, chain_exact=False)
                assert rel_err(a[i, AUX["reward"]], orr) <= RTOL and rel_err(a[i, AUX["download_time"]], oa["download_time"]) <= RTOL
    assert steps_done.min() >= 3 and np.all(steps_done == ep_len)
    stats = sim.episode_stats().cpu().numpy()
    assert np.array_equal(stats[:, 4], ep_len) and np.all(stats[:, 11] == 1)
    sim.close()


def test_full_size_simple_rl_one_million_envs():
    """BASELINE config 4 (run_simple_rl.py baseline over 1,048,576 envs, simulator only): size-independent
    properties over a full episode span with auto-reset -- every env finishes episodes of exactly the
    table-derived length, step counters stay in lockstep, the reward is the same float64 expression of the
    emitted QoE terms -- plus a strided sample of envs replayed by the scalar oracle."""
    cfg = SimConfig()
    N = 1 << 20
    base = synth.make_synthetic_tables(ViewportTiler(cfg).chunk_masks, n_videos=24, n_users=60, n_traces=40)
    tables = base.with_samples(synth.per_env_samples(base, N))
    sim = BatchSimulator(tables, N, OBS_MODE_SIMPLE, REWARD_QOE_NORM, seed=0)
    sim.reset()
    aux = sim.new_aux()
    check = np.arange(0, N, 65_537)
    orc = [so.OracleEnv(tables, OBS_MODE_SIMPLE, REWARD_QOE_NORM, "f64", worker_id=int(i), worker_num=N) for i in check]
    for o in orc:
        o.reset()
    pair = tables.samples[:, 0] * tables.n_users + tables.samples[:, 1]
    ep_len = np.minimum(tables.vp_end[pair], tables.video_time[tables.samples[:, 0]] - 1) - cfg.startup_download
    step_in_ep = np.zeros(N, dtype=np.int64)
    w = tables.qoe_w[tables.samples[:, 3]].astype(np.float64)
    for t in range(12):
        acts = synth.synthetic_actions(N, t, seed=5)
        o, r, d = sim.step(torch.from_numpy(acts), auto_reset=True, aux=aux)
        a = aux.cpu().numpy()
        step_in_ep += 1
        assert np.array_equal(a[:, AUX["ep_step"]], step_in_ep)
        assert np.array_equal(d.cpu().numpy().astype(bool), step_in_ep >= ep_len)
        qoe = w[:, 0] * a[:, AUX["qoe1"]] - w[:, 1] * a[:, AUX["qoe2"]] - w[:, 2] * a[:, AUX["qoe3"]]
        assert np.array_equal(qoe, a[:, AUX["qoe"]])
        assert np.array_equal(qoe / ((w[:, 0] + w[:, 1]) + w[:, 2]), a[:, AUX["reward"]])       # simple_rl_env.py:124-127
        step_in_ep[step_in_ep >= ep_len] = 0                                                      # auto-reset
        rows = o[torch.from_numpy(check).cuda()].cpu().numpy()
        for k, (o_env, i) in enumerate(zip(orc, check)):
            oo, orr, od, _ = o_env.step(int(acts[i]))
            if od:
                oo = o_env.reset()
            assert_rows_match(rows[k], so.flatten_obs(oo, OBS_MODE_SIMPLE), OBS_MODE_SIMPLE, chain_exact=False)
            assert rel_err(a[i, AUX["reward"]], orr) <= RTOL
    assert sim.error_flag() == 0
    sim.close()


@pytest.mark.parametrize("obs_mode", [OBS_MODE_MANSY, OBS_MODE_SIMPLE])
def test_outcome_table_equals_gather_path(obs_mode):
    """The (viewport pair, chunk, action) outcome table is filled by the step's own gather code: 20 000 environments stepped
    120 times (two episodes each, out-of-table actions included) with and without it give identical bits everywhere --
    observations, rewards, float64 internals (aux), state records, statistics."""
    n = 20000
    tables = _oracle_tables(n_videos=5, n_users=7, n_traces=9, seed=31, n_samples=n)
    a = BatchSimulator(tables, n, obs_mode, REWARD_QOE, seed=3)
    b = BatchSimulator(tables, n, obs_mode, REWARD_QOE, seed=3)
    b.set_outcome_table(False)
    assert torch.equal(a.reset(), b.reset())
    aux_a, aux_b = a.new_aux(), b.new_aux()
    rng = np.random.default_rng(5)
    for t in range(120):
        acts = torch.from_numpy(rng.integers(0, 16 if t % 7 == 0 else 15, size=n).astype(np.int32)).cuda()
        oa, ra, da = a.step(acts, auto_reset=True, aux=aux_a)
        ob, rb, db = b.step(acts, auto_reset=True, aux=aux_b)
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db) and torch.equal(aux_a, aux_b), t
    assert torch.equal(a.episode_stats(), b.episode_stats())
    sa, sb = a.episode_state_host(), b.episode_state_host()
    assert sa.tobytes() == sb.tobytes()
    assert a.error_flag() == 0 and b.error_flag() == 0
