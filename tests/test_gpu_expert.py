"""MPC expert (csrc/mansy_sim.cu expert_mpc_kernel) against the oracle and the golden written from the UNMODIFIED
reference ExpertEnv (oracle/make_golden_expert.py).  Actions are integers: bit-exact."""
import numpy as np
import pytest
import torch

from helpers import golden_tables, load_golden
from mansy_immersivevideostreaming_b200 import synth
from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, REWARD_QOE
from mansy_immersivevideostreaming_b200.simulator import BatchSimulator
from oracle import sim_oracle as so

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("horizon", [1, 2, 3, 4])      # 4 = the reference's default horizon (10 decisions, 50 625 sequences each)
def test_reference_golden_decisions(horizon):
    """Teacher-forced with the reference's own actions; the kernel's decisions equal the float64-chain oracle's at every
    step (the chain step_env follows; the reference under numpy 2 runs float32, whose decisions the fixture also holds)."""
    g = load_golden("expert_kat.npz")
    tables = golden_tables(g)
    sim = BatchSimulator(tables, 1, OBS_MODE_MANSY, REWARD_QOE, seed=0, worker_num=1)
    sim.reset()
    ref_actions, want, value = g[f"h{horizon}_actions_ref"], g[f"h{horizon}_actions_f64"], g[f"h{horizon}_value_f64"]
    for k in range(len(want)):
        a, v = sim.expert_actions(horizon, return_value=True)
        assert int(a[0]) == int(want[k]), (horizon, k)
        assert float(v[0]) == float(value[k])               # same float64 operations in the same order
        sim.step(torch.tensor([int(ref_actions[k])], dtype=torch.int32), auto_reset=True)
    agree = float(np.mean(ref_actions == want))
    print(f"horizon {horizon}: {len(want)} decisions; float64 chain == reference's float32 chain on {agree * 100:.1f}%")


def test_batch_vs_oracle_following_the_expert():
    """64 environments following their own expert decisions (horizon 2) for 60 steps with auto-reset."""
    g = load_golden("mansy_synth.npz")
    tables = golden_tables(g)
    n, horizon = 64, 2
    tables = tables.with_samples(synth.per_env_samples(tables, n))
    sim = BatchSimulator(tables, n, OBS_MODE_MANSY, REWARD_QOE, seed=3)
    orc = so.OracleVectorEnv(tables, n, OBS_MODE_MANSY, REWARD_QOE, chain="f64", seed=3)
    sim.reset(); orc.reset()
    for t in range(60):
        a = sim.expert_actions(horizon)
        want = np.array([so.expert_choose_action(e, horizon) for e in orc.envs], dtype=np.int32)
        assert np.array_equal(a.cpu().numpy(), want), t
        _, rew, done = sim.step(a, auto_reset=True)
        _, orew, odone, _ = orc.step(want, auto_reset=True)
        np.testing.assert_allclose(rew.cpu().numpy(), orew, rtol=1e-5, atol=1e-6)
        assert np.array_equal(done.cpu().numpy().astype(bool), odone)


def test_horizon_4_matches_oracle_and_state_is_untouched():
    g = load_golden("mansy_synth.npz")
    tables = golden_tables(g)
    n = 3
    sim = BatchSimulator(tables, n, OBS_MODE_MANSY, REWARD_QOE, seed=1)
    orc = so.OracleVectorEnv(tables, n, OBS_MODE_MANSY, REWARD_QOE, chain="f64", seed=1)
    sim.reset(); orc.reset()
    for t in range(2):
        before = sim.episode_state_host().copy()
        a, v = sim.expert_actions(4, return_value=True)
        assert np.array_equal(sim.episode_state_host(), before)           # virtual downloads leave no trace
        res = [so.expert_choose_action(e, 4, return_value=True) for e in orc.envs]
        assert np.array_equal(a.cpu().numpy(), np.array([r[0] for r in res], dtype=np.int32))
        assert np.array_equal(v.cpu().numpy(), np.array([r[1] for r in res]))
        sim.step(a, auto_reset=True); orc.step(a.cpu().numpy(), auto_reset=True)


def test_expert_beats_fixed_actions_at_full_width():
    """Property at a size the oracle cannot reach: 4,096 envs, horizon 4 (50,625 sequences each): the winning sum is at
    least the QoE sum of the greedy (horizon-1) policy really stepped over the same chunks -- one of the sequences."""
    g = load_golden("mansy_synth.npz")
    tables = golden_tables(g)
    n = 4096
    tables = tables.with_samples(synth.per_env_samples(tables, n))
    sim = BatchSimulator(tables, n, OBS_MODE_MANSY, REWARD_QOE, seed=2)
    sim.reset()
    a4, v4 = sim.expert_actions(4, return_value=True)
    a3, v3 = sim.expert_actions(3, return_value=True)
    assert int(a4.min()) >= 0 and int(a4.max()) < 15 and sim.error_flag() == 0
    # greedy 4-step roll-out with real steps: its reward sum cannot exceed the exhaustive optimum over the same 4 chunks
    aux = sim.new_aux()
    total = torch.zeros(n, dtype=torch.float64, device=sim.device)
    alive = torch.ones(n, dtype=torch.bool, device=sim.device)
    for t in range(4):
        a1 = sim.expert_actions(1)
        _, _, done = sim.step(a1, auto_reset=False, aux=aux)
        total += torch.where(alive, aux[:, 6], torch.zeros_like(total))          # MANSY_AUX_QOE
        alive &= ~done.bool()
    assert bool((total <= v4 + 1e-9).all())
