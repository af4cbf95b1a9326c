"""The oracle's MPC expert against the golden decisions of the UNMODIFIED reference ExpertEnv (CPU only)."""
import numpy as np
import pytest

from helpers import golden_tables, load_golden
from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, REWARD_QOE
from oracle import sim_oracle as so


@pytest.mark.parametrize("horizon", [1, 2])
def test_oracle_reproduces_reference_expert(horizon):
    g = load_golden("expert_kat.npz")
    tables = golden_tables(g)
    ref_actions, f64_actions = g[f"h{horizon}_actions_ref"], g[f"h{horizon}_actions_f64"]
    o32 = so.OracleEnv(tables, OBS_MODE_MANSY, REWARD_QOE, "f32", worker_id=0, worker_num=1)
    o64 = so.OracleEnv(tables, OBS_MODE_MANSY, REWARD_QOE, "f64", worker_id=0, worker_num=1)
    o32.reset(); o64.reset()
    for k in range(min(len(ref_actions), 110)):
        assert so.expert_choose_action(o32, horizon) == int(ref_actions[k]), k     # the chain the reference ran in
        assert so.expert_choose_action(o64, horizon) == int(f64_actions[k]), k
        _, _, d, _ = o32.step(int(ref_actions[k]))
        o64.step(int(ref_actions[k]))
        if d:
            o32.reset(); o64.reset()


def test_short_tail_uses_the_chunks_that_are_left():
    """expert_env.py:362: horizon = min(horizon, end_chunk - next_chunk + 1); at the last chunk every horizon agrees."""
    g = load_golden("expert_kat.npz")
    env = so.OracleEnv(golden_tables(g), OBS_MODE_MANSY, REWARD_QOE, "f64", worker_id=0, worker_num=1)
    env.reset()
    while env.next_chunk < env.end_chunk:
        env.step(3)
    assert so.expert_choose_action(env, 1) == so.expert_choose_action(env, 3)
