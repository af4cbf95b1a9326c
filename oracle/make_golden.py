"""Generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE -- build container only.

Usage (from the repo root, in the container that has /root/reference):

    python -m oracle.make_golden

Every fixture is produced by the reference's own code (imported through oracle/ref_loader.py;
numpy 2.x here, so the reference computes its QoE chain in float32 -- see oracle/sim_oracle.py
``chain="f32"``).  While generating, the script also asserts that the oracle restatement
reproduces the reference bit for bit (f32 chain), which is what pins the oracle.

Fixtures (all small, committed):
  geometry_kat.npz   a13-a16: reference masks for a grid of pixel centres incl. every tile/
                     wrap boundary, plus shipped ground-truth masks + their 5 Hz centres.
  allocate_kat.npz   a1-a2: reference tile_rate_versions for masks x 15 actions.
  trace_kat.npz      a4-a5: reference NetworkTrace / PlaybackBuffer sequences.
  mansy_synth.npz    a3,a6-a10,a12: MANSYEnv episodes on a synthetic dataset (tables inside).
  simple_synth.npz   a11: SimpleRLEnv episodes on the same synthetic dataset.
  mansy_real.npz     MANSYEnv episodes on a slice of the shipped Jin2022/4G data (tables inside).
  policy_kat.npz     a17-a18: reference Actor/Critic/QoEIdentifier forward (fp32, CPU) on
                     numpy-seeded weights and observations taken from the synthetic episodes.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mansy_immersivevideostreaming_b200 import synth                                      # noqa: E402
from mansy_immersivevideostreaming_b200.config import (OBS_MODE_MANSY, OBS_MODE_SIMPLE, REWARD_QOE,  # noqa: E402
                                                       REWARD_QOE_NORM, SimConfig)
from mansy_immersivevideostreaming_b200.tables import (SimTables, environment_test_samples, masks_to_u64,  # noqa: E402
                                                       pack_from_reference_layout)
from oracle import sim_oracle as so                                                       # noqa: E402
from oracle.ref_loader import REFERENCE_ROOT, load_reference, silence_prints              # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CFG = SimConfig()


def _ref_mask(ref, x, y):
    return ref.vp_common.find_tiles_covered_by_viewport(int(x), int(y), CFG.video_width, CFG.video_height,
                                                        CFG.tile_width, CFG.tile_height, CFG.tile_num_width,
                                                        CFG.tile_num_height).reshape(-1)


def make_geometry(ref):
    # (1) reference function on a boundary-heavy grid of pixel centres
    xs = sorted(set([0, 1, 19, 20, 21, 299, 300, 301, 319, 320, 321, 339, 340, 341, 620, 639, 640, 641, 940,
                     1279, 1280, 1281, 1600, 1939, 1940, 1941, 2239, 2240, 2241, 2259, 2260, 2261, 2540, 2559,
                     2560] + list(range(7, 2560, 97))))
    ys = sorted(set([0, 1, 29, 30, 31, 149, 150, 151, 179, 180, 181, 209, 210, 211, 329, 330, 331, 360, 539, 540,
                     541, 720, 1079, 1080, 1110, 1111, 1259, 1260, 1261, 1289, 1290, 1291, 1410, 1439, 1440]
                    + list(range(5, 1440, 83))))
    gx, gy = np.meshgrid(np.asarray(xs), np.asarray(ys), indexing="ij")
    gx, gy = gx.reshape(-1).astype(np.int32), gy.reshape(-1).astype(np.int32)
    masks = np.zeros(gx.shape[0], dtype=np.uint64)
    for i in range(gx.shape[0]):
        m = _ref_mask(ref, gx[i], gy[i])
        masks[i] = so.mask_bits(m)
        assert masks[i] == so.mask_bits(so.fov_tile_mask(int(gx[i]), int(gy[i]), CFG))
    # (2) shipped ground-truth masks + the 5 Hz centres that produce them (predict.py:33-48;
    #     sample i uses rows 16+5i..20+5i of the npy, chunk id i+3: load_dataset.py:36-37,50)
    import pickle
    pairs = [(21, 3), (14, 10), (16, 24), (1, 22), (9, 44), (24, 60), (5, 7), (12, 13)]
    ship_xy, ship_gt, ship_chunk, ship_pair = [], [], [], []
    vp_root = os.path.join(REFERENCE_ROOT, "datasets", "Jin2022", "viewports")
    for v, u in pairs:
        arr = np.load(os.path.join(vp_root, f"video{v}", "5Hz", f"simple_5Hz_user{u}.npy"))
        lst = pickle.load(open(os.path.join(vp_root, "prediction", f"video{v}", f"user{u}.pkl"), "rb"))
        for chunk, gt, _pred, _acc in lst:
            pts = arr[5 * chunk + 1: 5 * chunk + 6, 1:3].astype(np.float32)
            assert pts.shape == (5, 2)
            ship_xy.append(pts); ship_gt.append(so.mask_bits(gt)); ship_chunk.append(chunk); ship_pair.append((v, u))
    ship_xy = np.stack(ship_xy)
    ship_gt = np.asarray(ship_gt, dtype=np.uint64)
    # the oracle must reproduce every shipped mask from the shipped centres (both numeric chains)
    for chain in ("f64", "f32"):
        g, _p, _a = so.chunk_masks(ship_xy, ship_xy, CFG, chain=chain)
        assert np.array_equal(g, ship_gt), f"oracle({chain}) does not reproduce the shipped gt masks"
    # (3) reference per-chunk OR + IoU on synthetic centres (predict.py:33-48 restated around the
    #     reference's own mask function; pixel conversion per the pinned-stack float64 product)
    rng = np.random.default_rng(99)
    gt_xy = rng.random((256, 5, 2)).astype(np.float32)
    gt_xy[:8] = np.array([0.0, 1.0, 0.5, 0.1171875, 0.8828125, 0.125, 0.875, 0.25])[:, None, None]
    pred_xy = np.mod(gt_xy + rng.normal(0, 0.05, gt_xy.shape), 1.0).astype(np.float32)
    cm_gt = np.zeros(256, np.uint64); cm_pred = np.zeros(256, np.uint64); cm_acc = np.zeros(256, np.float64)
    for i in range(256):
        g = np.zeros(64, np.uint8); p = np.zeros(64, np.uint8)
        for j in range(5):
            g |= _ref_mask(ref, int(float(gt_xy[i, j, 0]) * CFG.video_width), int(float(gt_xy[i, j, 1]) * CFG.video_height))
            p |= _ref_mask(ref, int(float(pred_xy[i, j, 0]) * CFG.video_width), int(float(pred_xy[i, j, 1]) * CFG.video_height))
        cm_gt[i], cm_pred[i] = so.mask_bits(g), so.mask_bits(p)
        cm_acc[i] = np.sum(g & p) / np.sum(g | p)
    og, op, oa = so.chunk_masks(gt_xy, pred_xy, CFG)
    assert np.array_equal(og, cm_gt) and np.array_equal(op, cm_pred) and np.array_equal(oa, cm_acc)
    np.savez_compressed(os.path.join(GOLDEN, "geometry_kat.npz"), x=gx, y=gy, mask=masks,
                        ship_xy=ship_xy, ship_gt=ship_gt, ship_chunk=np.asarray(ship_chunk, np.int32),
                        ship_pair=np.asarray(ship_pair, np.int32),
                        cm_gt_xy=gt_xy, cm_pred_xy=pred_xy, cm_gt=cm_gt, cm_pred=cm_pred, cm_acc=cm_acc)
    print(f"geometry_kat: {gx.shape[0]} grid centres, {ship_gt.shape[0]} shipped chunks, 256 chunk-mask cases")


def make_allocate(ref):
    rng = np.random.default_rng(5)
    masks = [0, (1 << 64) - 1, 1, 1 << 63, 1 << 27, 0x00003C3C3C000000, 0x8100000000000081, 0xFF, 0xFF << 56,
             0x0101010101010101, 0x8080808080808080]
    # FoV-shaped masks from the geometry and random sparse/dense masks
    for _ in range(40):
        x, y = int(rng.integers(0, 2561)), int(rng.integers(0, 1441))
        masks.append(so.mask_bits(so.fov_tile_mask(x, y, CFG)))
    for dens in (0.05, 0.2, 0.5):
        for _ in range(12):
            masks.append(so.mask_bits((rng.random(64) < dens).astype(np.uint8)))
    masks = np.asarray(masks, dtype=np.uint64)
    out = np.zeros((masks.shape[0], 16, 64), dtype=np.uint8)     # action 15 = out-of-table action
    for i, m in enumerate(masks):
        pv = np.array([(int(m) >> t) & 1 for t in range(64)], dtype=np.float32)
        for a in range(16):
            rin, rout = ref.common.action2rates(a)
            ver, _ = ref.common.allocate_tile_rates(rin, rout, pv, list(CFG.video_rates), 8, 8)
            out[i, a] = ver
            assert np.array_equal(ver, so.allocate_tile_versions(*so.action_to_rates(a), int(m), CFG.video_rates))
    np.savez_compressed(os.path.join(GOLDEN, "allocate_kat.npz"), mask=masks, versions=out)
    print(f"allocate_kat: {masks.shape[0]} masks x 16 actions")


def make_trace(ref, root):
    import pickle
    rng = np.random.default_rng(11)
    cases = []
    for k in range(12):
        L = int(rng.integers(3, 40))
        thr = rng.integers(0, 9_000_000, size=L)
        thr[rng.random(L) < 0.25] = 0
        if k == 0:
            thr[:] = 0; thr[-1] = 1_000_000                     # long stall, wraps the trace
        if not np.any(thr > 0):
            thr[0] = 4_000_000
        if k % 3 == 2:
            thr = thr.astype(np.float64) * 0.37 + 0.5               # non-integer (rescaled-trace style)
            path = os.path.join(root, f"kat_trace_{k}.pkl")
            pickle.dump([(i, float(thr[i])) for i in range(L)], open(path, "wb"))
        else:
            path = os.path.join(root, f"kat_trace_{k}.pkl")
            pickle.dump([(i, int(thr[i])) for i in range(L)], open(path, "wb"))
        net = ref.simulator.NetworkTrace(path)
        buf = ref.simulator.PlaybackBuffer(5, 1)
        sizes = rng.integers(200_000, 9_000_000, size=40)
        rec = []
        idx, tm, b = 0, 0.0, 3.0
        for s in sizes:
            dl = net.simulate_download(int(s))
            rb = buf.push_chunk(1, dl)
            odl, idx, tm = so.trace_download(int(s), np.asarray(thr, dtype=np.float64), L, idx, tm)
            orb, b = so.buffer_push(b, 1, odl)
            assert (dl, net.cur_idx, net.cur_time, rb, buf.get_buffer_size()) == (odl, idx, tm, orb, b)
            rec.append((dl, net.cur_idx, net.cur_time, rb, buf.get_buffer_size()))
        cases.append((np.asarray(thr, dtype=np.float64), sizes.astype(np.int64), np.asarray(rec, dtype=np.float64)))
    Lmax = max(c[0].shape[0] for c in cases)
    thr = np.zeros((len(cases), Lmax)); lens = np.zeros(len(cases), np.int32)
    for i, c in enumerate(cases):
        thr[i, : c[0].shape[0]] = c[0]; lens[i] = c[0].shape[0]
    np.savez_compressed(os.path.join(GOLDEN, "trace_kat.npz"), thr=thr, lens=lens,
                        sizes=np.stack([c[1] for c in cases]), rec=np.stack([c[2] for c in cases]))
    print(f"trace_kat: {len(cases)} traces x 40 downloads")


def _run_env_episodes(ref, env, oracle_f32, oracle_f64, obs_mode, n_episodes, rng, action_fn=None):
    """Teacher-forced episodes: the same action stream drives the reference and both oracles."""
    rows, rews, dones, acts, aux_rows, vers, ep_rows = [], [], [], [], [], [], []
    max_rel = 0.0
    for ep in range(n_episodes):
        s_ref = env.reset(); s32 = oracle_f32.reset(); oracle_f64.reset()
        for k in s_ref:
            assert np.array_equal(np.asarray(s_ref[k]), s32[k]), k
        rows.append(so.flatten_obs(s_ref, obs_mode)); rews.append(0.0); dones.append(False); acts.append(-1)
        aux_rows.append(np.zeros(10)); vers.append(np.zeros(64, np.uint8))
        done = False
        step = 0
        while not done:
            a = int(rng.integers(0, 15)) if action_fn is None else action_fn(ep, step)
            s_ref, r_ref, done, _ = env.step(a)
            s32, r32, d32, aux32 = oracle_f32.step(a)
            _s64, r64, d64, aux64 = oracle_f64.step(a)
            assert done == d32 == d64
            for k in s_ref:
                assert np.array_equal(np.asarray(s_ref[k]), s32[k]), (k, ep, step)
            assert float(r_ref) == float(r32), (r_ref, r32)
            assert list(env.tile_rates) == list(aux32["versions"])
            assert env.simulator.net_trace.cur_time == aux32["cur_time"] and env.simulator.net_trace.cur_idx == aux32["cur_idx"]
            assert env.simulator.get_buffer_size() == aux32["buffer"]
            w = oracle_f64.w
            scale = abs(float(w[0]) * aux64["qoe1"]) + abs(float(w[1]) * aux64["qoe2"]) + abs(float(w[2]) * aux64["qoe3"])
            max_rel = max(max_rel, abs(aux64["qoe"] - aux32["qoe"]) / max(scale, 1e-30))
            rows.append(so.flatten_obs(s_ref, obs_mode)); rews.append(float(r_ref)); dones.append(done); acts.append(a)
            # reference-side internals (python floats, identical in both numeric chains)
            aux_rows.append(np.array([aux32["chunk_size"], aux32["download_time"], aux32["rebuffer"], aux32["buffer"],
                                      aux32["cur_idx"], aux32["cur_time"], float(env.qoe_model.qoe1),
                                      float(env.qoe_model.qoe2), float(env.qoe_model.qoe3), aux32["next_chunk"]]))
            vers.append(np.asarray(env.tile_rates, dtype=np.uint8))
            step += 1
        e = oracle_f32.episodes[-1]
        ep_rows.append([e["video"], e["user"], e["trace"], *e["w"], e["steps"], e["sample_id"]])
    return dict(obs=np.stack(rows), reward=np.asarray(rews, np.float64), done=np.asarray(dones),
                action=np.asarray(acts, np.int32), aux=np.stack(aux_rows), versions=np.stack(vers),
                episodes=np.asarray(ep_rows, np.float64)), max_rel


def make_env_goldens(ref, root):
    tables = synth.make_synthetic_tables(lambda g, p: so.chunk_masks(g, p, CFG), n_videos=3, n_users=4,
                                         n_traces=5, seed=7, trace_len_range=(40, 90), short_tail_frac=0.3)
    cfg_path = synth.write_reference_layout(tables, root)
    config = ref.common.get_config_from_yml(cfg_path)
    qoe_weights = [[float(x) for x in w] for w in tables.qoe_w]
    # the packer must give back the same tables from the files the reference reads
    packed = pack_from_reference_layout(config, "Synth", "SynthNet", [int(v) for v in tables.video_ids],
                                        [int(u) for u in tables.user_ids], [int(t) for t in tables.trace_ids],
                                        qoe_weights, mode="train")
    for k in SimTables._ARRAYS:
        assert np.array_equal(getattr(packed, k), getattr(tables, k)), k

    out = {}
    # MANSY: train mode (reward = qoe: use_identifier False) and the "use_identifier" normalised reward
    for tag, mode, use_id, rmode, worker in (("train", "train", False, REWARD_QOE, (1, 3)),
                                             ("norm", "train", True, REWARD_QOE_NORM, (0, 1))):
        with silence_prints():
            env = ref.mansy_env.MANSYEnv(config, "Synth", "SynthNet", qoe_weights, None, 0.5,
                                         os.path.join(root, f"log_{tag}.csv"), CFG.startup_download, mode=mode,
                                         seed=0, worker_num=worker[1], use_identifier=use_id)
        env.seed(worker[0])
        o32 = so.OracleEnv(tables, OBS_MODE_MANSY, rmode, "f32", worker_id=worker[0], worker_num=worker[1])
        o64 = so.OracleEnv(tables, OBS_MODE_MANSY, rmode, "f64", worker_id=worker[0], worker_num=worker[1])
        rec, max_rel = _run_env_episodes(ref, env, o32, o64, OBS_MODE_MANSY, 5, np.random.default_rng(3))
        print(f"mansy_synth[{tag}]: {rec['obs'].shape[0]} rows, f64-vs-f32 reward gap (magnitude-aware) {max_rel:.3e}")
        for k, v in rec.items():
            out[f"{tag}_{k}"] = v
        out[f"{tag}_worker"] = np.asarray(worker, np.int32)
        out[f"{tag}_log"] = np.asarray(open(os.path.join(root, f"log_{tag}.csv")).read())
    out.update(tables.to_npz_dict())
    np.savez_compressed(os.path.join(GOLDEN, "mansy_synth.npz"), **out)

    out = {}
    for tag, mode, rmode in (("train", "train", REWARD_QOE_NORM), ("test", "test", REWARD_QOE)):
        tb = tables if mode == "train" else tables.with_samples(
            environment_test_samples(tables.n_videos, tables.n_users, tables.n_traces, tables.qoe_w.shape[0]))
        with silence_prints():
            env = ref.simple_rl_env.SimpleRLEnv(config, "Synth", "SynthNet", qoe_weights,
                                                os.path.join(root, f"slog_{tag}.csv"), CFG.startup_download,
                                                mode=mode, seed=0, worker_num=2)
        env.seed(1)
        o32 = so.OracleEnv(tb, OBS_MODE_SIMPLE, rmode, "f32", worker_id=1, worker_num=2)
        o64 = so.OracleEnv(tb, OBS_MODE_SIMPLE, rmode, "f64", worker_id=1, worker_num=2)
        rec, max_rel = _run_env_episodes(ref, env, o32, o64, OBS_MODE_SIMPLE, 4, np.random.default_rng(4))
        print(f"simple_synth[{tag}]: {rec['obs'].shape[0]} rows, gap {max_rel:.3e}")
        for k, v in rec.items():
            out[f"{tag}_{k}"] = v
    out.update(tables.to_npz_dict())
    np.savez_compressed(os.path.join(GOLDEN, "simple_synth.npz"), **out)
    return tables, config


def make_real_golden(ref):
    """A slice of the shipped data: test videos 21/14, users 3/10/14, traces 31/33, test QoE."""
    config = ref.common.get_config_from_yml(os.path.join(REFERENCE_ROOT, "config.yml"))
    # paths in config.yml are relative to bitrate_selection/ (utils/common.py:10)
    for d in (config.viewport_datasets_dir, config.video_datasets_dir, config.network_datasets_dir):
        for k in d:
            d[k] = os.path.normpath(os.path.join(REFERENCE_ROOT, "bitrate_selection", d[k]))
    videos, users, traces = [21, 14], [3, 10, 14], [31, 33]
    qoe_weights = config.qoe_split["test"]
    config.video_split["Jin2022"]["test"] = videos
    config.user_split["Jin2022"]["test"] = users
    config.network_split["4G"]["test"] = traces
    tables = pack_from_reference_layout(config, "Jin2022", "4G", videos, users, traces, qoe_weights, mode="test")
    log = os.path.join(tempfile.mkdtemp(), "real_log.csv")
    with silence_prints():
        env = ref.mansy_env.MANSYEnv(config, "Jin2022", "4G", qoe_weights, None, 0.5, log, CFG.startup_download,
                                     mode="test", seed=1, worker_num=1)
    env.seed(1)
    # stride through the 48 test samples with worker_num=7 so videos/users/traces/weights all vary
    env.worker_num, env.worker_id = 7, 1
    o32 = so.OracleEnv(tables, OBS_MODE_MANSY, REWARD_QOE, "f32", worker_id=1, worker_num=7)
    o64 = so.OracleEnv(tables, OBS_MODE_MANSY, REWARD_QOE, "f64", worker_id=1, worker_num=7)
    rec, max_rel = _run_env_episodes(ref, env, o32, o64, OBS_MODE_MANSY, 6, np.random.default_rng(8))
    print(f"mansy_real: {rec['obs'].shape[0]} rows, gap {max_rel:.3e}")
    out = {f"test_{k}": v for k, v in rec.items()}
    out["test_worker"] = np.asarray((1, 7), np.int32)
    out["test_log"] = np.asarray(open(log).read())
    out.update(tables.to_npz_dict())
    np.savez_compressed(os.path.join(GOLDEN, "mansy_real.npz"), **out)


from mansy_immersivevideostreaming_b200.policy import seeded_state_dict as numpy_state_dict  # noqa: E402


def make_policy(ref):
    import torch
    torch.manual_seed(0)
    g = np.load(os.path.join(GOLDEN, "mansy_synth.npz"))
    gs = np.load(os.path.join(GOLDEN, "simple_synth.npz"))
    from mansy_immersivevideostreaming_b200.config import MANSY_OBS_SEGMENTS, SIMPLE_OBS_SEGMENTS

    def unflatten(rows, segs):
        return {k: rows[:, off:off + int(np.prod(shape))].reshape((-1,) + tuple(shape)).copy() for k, off, shape in segs}

    rows = g["train_obs"][::7][:32]
    obs = unflatten(rows, MANSY_OBS_SEGMENTS)
    M = ref.models_mansy
    fn = M.FeatureNet(8, 64, 5, 128, device="cpu")
    actor = M.Actor(fn, 1280, 128, 15, "cpu")
    critic = M.Critic(fn, 1280, 128, "cpu")
    ifn = M.QoEIdentifierFeatureNet(8, 64, 5, 15, 128, device="cpu")
    ident = M.QoEIdentifier(ifn, 1280, 128, "cpu")
    out = {"mansy_rows": rows}
    for tag, mod, seed in (("actor", actor, 101), ("critic", critic, 102), ("ident", ident, 103)):
        sd = mod.state_dict()
        if tag == "critic":      # the feature net is SHARED with the actor (run_mansy.py:207-209)
            new = {k: (actor.state_dict()[k] if k.startswith("feature_net.") else None) for k in sd}
            gen = numpy_state_dict([(k, tuple(v.shape)) for k, v in sd.items() if not k.startswith("feature_net.")], seed)
            new.update({k: torch.from_numpy(v) for k, v in gen.items()})
        else:
            gen = numpy_state_dict([(k, tuple(v.shape)) for k, v in sd.items()], seed)
            new = {k: torch.from_numpy(v) for k, v in gen.items()}
        mod.load_state_dict(new)
        out[f"{tag}_names"] = np.asarray(list(sd.keys()))
        out[f"{tag}_shapes"] = np.asarray([",".join(map(str, v.shape)) for v in sd.values()])
    with torch.no_grad():
        logits, _ = actor(obs)
        value = critic(obs)
        probs = ident(obs, obs["action_one_hot"])
    out.update(actor_logits=logits.numpy(), critic_value=value.numpy(), ident_out=probs.numpy())

    srows = gs["train_obs"][::5][:32]
    sobs = unflatten(srows, SIMPLE_OBS_SEGMENTS)
    S = ref.models_simple
    sfn = S.FeatureNet(8, 64, 5, device="cpu")
    sactor = S.Actor(sfn, 640, 15, "cpu")
    scritic = S.Critic(sfn, 640, "cpu")
    sd = sactor.state_dict()
    gen = numpy_state_dict([(k, tuple(v.shape)) for k, v in sd.items()], 201)
    sactor.load_state_dict({k: torch.from_numpy(v) for k, v in gen.items()})
    sdc = scritic.state_dict()
    genc = numpy_state_dict([(k, tuple(v.shape)) for k, v in sdc.items() if not k.startswith("feature_net.")], 202)
    newc = {k: sactor.state_dict()[k] for k in sdc if k.startswith("feature_net.")}
    newc.update({k: torch.from_numpy(v) for k, v in genc.items()})
    scritic.load_state_dict(newc)
    with torch.no_grad():
        sprobs, _ = sactor(sobs)
        svalue = scritic(sobs)
    out.update(simple_rows=srows, simple_actor_names=np.asarray(list(sd.keys())),
               simple_actor_shapes=np.asarray([",".join(map(str, v.shape)) for v in sd.values()]),
               simple_critic_names=np.asarray(list(sdc.keys())),
               simple_critic_shapes=np.asarray([",".join(map(str, v.shape)) for v in sdc.values()]),
               simple_probs=sprobs.numpy(), simple_value=svalue.numpy())
    np.savez_compressed(os.path.join(GOLDEN, "policy_kat.npz"), **out)
    print("policy_kat: actor/critic/identifier + simple_rl actor/critic on 32 observations each")


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    ref = load_reference()
    root = tempfile.mkdtemp(prefix="mansy_golden_")
    make_geometry(ref)
    make_allocate(ref)
    make_trace(ref, root)
    make_env_goldens(ref, root)
    make_real_golden(ref)
    make_policy(ref)
    for f in sorted(os.listdir(GOLDEN)):
        print(f"  {f}: {os.path.getsize(os.path.join(GOLDEN, f)) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
