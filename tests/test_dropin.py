"""Drop-in env modules (mansy_immersivevideostreaming_b200/dropin): the reference's constructor signatures
and gym-style behaviour on a dataset written in the reference's on-disk formats."""
from __future__ import annotations

import inspect
import os
import tempfile

import numpy as np
import pytest

from mansy_immersivevideostreaming_b200 import synth
from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, OBS_MODE_SIMPLE, REWARD_QOE, REWARD_QOE_NORM, SimConfig
from mansy_immersivevideostreaming_b200.dropin.envs import MANSYEnv, SimpleRLEnv
from mansy_immersivevideostreaming_b200.dropin.envs._common import tables_for
from mansy_immersivevideostreaming_b200.refconfig import load_config_yml
from mansy_immersivevideostreaming_b200.tables import SimTables
from oracle import sim_oracle as so

from helpers import assert_rows_match, segments

CFG = SimConfig()


def _dataset(root):
    t = synth.make_synthetic_tables(lambda g, p: so.chunk_masks(g, p, CFG), n_videos=2, n_users=3, n_chunks=30,
                                    n_traces=3, seed=11, trace_len_range=(30, 60))
    return t, synth.write_reference_layout(t, root)


def test_signatures_match_reference():
    """Parameter names/defaults of bitrate_selection/envs/mansy_env.py:19-20 and simple_rl_env.py:15-16."""
    p = list(inspect.signature(MANSYEnv.__init__).parameters)[1:]
    assert p == ["config", "dataset", "network_dataset", "qoe_weights", "identifier", "lamb", "log_path",
                 "startup_download", "mode", "seed", "worker_num", "device", "use_identifier"]
    d = {k: v.default for k, v in inspect.signature(MANSYEnv.__init__).parameters.items() if v.default is not inspect._empty}
    assert d == {"mode": "train", "seed": 0, "worker_num": 1, "device": "cpu", "use_identifier": False}
    p = list(inspect.signature(SimpleRLEnv.__init__).parameters)[1:]
    assert p == ["config", "dataset", "network_dataset", "qoe_weights", "log_path", "startup_download", "mode",
                 "seed", "worker_num", "device"]


def test_config_loader_and_table_cache():
    root = tempfile.mkdtemp()
    t, cfg_path = _dataset(root)
    config = load_config_yml(cfg_path)
    assert config.video_datasets_dir["Synth"].startswith(os.path.join(root, "datasets"))      # utils/common.py:21-25
    assert config.bs_results_dir == os.path.join(root, "results") + os.sep + "bitrate_selection"
    w = config.qoe_split["test"]
    packed = tables_for(config, "Synth", "SynthNet", w, "test", config.startup_download)
    assert packed is tables_for(config, "Synth", "SynthNet", w, "test", config.startup_download)   # packed once
    for k in SimTables._ARRAYS:
        if k != "samples":
            assert np.array_equal(getattr(packed, k), getattr(t, k)), k
    assert packed.n_samples == 2 * 3 * 3 * len(w)                                               # utils/common.py:96-97


@pytest.mark.gpu
@pytest.mark.parametrize("cls,obs_mode,mode", [(MANSYEnv, OBS_MODE_MANSY, "test"), (SimpleRLEnv, OBS_MODE_SIMPLE, "train")])
def test_dropin_env_matches_oracle(cls, obs_mode, mode):
    root = tempfile.mkdtemp()
    _, cfg_path = _dataset(root)
    config = load_config_yml(cfg_path)
    w = config.qoe_split[mode]
    log = os.path.join(root, "results.csv")
    if cls is MANSYEnv:
        env = cls(config, "Synth", "SynthNet", w, None, 0.5, log, config.startup_download, mode=mode, seed=3, worker_num=2)
        reward_mode = REWARD_QOE
    else:
        env = cls(config, "Synth", "SynthNet", w, log, config.startup_download, mode=mode, seed=3, worker_num=2)
        reward_mode = REWARD_QOE_NORM
    tables = tables_for(config, "Synth", "SynthNet", w, mode, config.startup_download)
    orc = so.OracleEnv(tables, obs_mode, reward_mode, "f64", worker_id=3 % 2, worker_num=2)
    assert env.sample_count() == tables.n_samples and env.action_space.n == (15 if cls is MANSYEnv else 5)
    rng = np.random.default_rng(0)
    for ep in range(3):
        state = env.reset()
        ref = so.flatten_obs(orc.reset(), obs_mode)
        assert env.current_video == int(tables.video_ids[tables.samples[orc.sample_id][0]])
        assert env.current_trace == int(tables.trace_ids[tables.samples[orc.sample_id][2]])
        done = False
        while not done:
            for key, off, shape in segments(obs_mode):
                n = int(np.prod(shape))
                assert state[key].shape == tuple(shape) and state[key].dtype == np.float32
                np.testing.assert_allclose(state[key].reshape(-1), ref[off:off + n], rtol=1e-5, atol=1e-7, err_msg=key)
            a = int(rng.integers(0, 15))
            state2, r, done, info = env.step(a)
            obs_o, r_ref, d_ref, _ = orc.step(a)
            ref = so.flatten_obs(obs_o, obs_mode)
            assert state2 is state and info == {} and isinstance(done, bool) and done == d_ref   # old-gym 4-tuple
            assert abs(r - r_ref) <= 1e-5 * max(abs(r_ref), 1e-3)
    env.close()
    lines = open(log).read().strip().splitlines()
    assert lines[0] == "video,user,trace,qoe_w1,qoe_w2,qoe_w3,qoe,qoe1,qoe2,qoe3" and len(lines) == 4
