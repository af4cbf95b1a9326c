"""The per-rollout exchange of episode totals (csrc/mansy_peer.cu) and the host-side guards of the vector env.

World size 1 runs everywhere; the 2-rank NVLink path runs when the box has two GPUs (`gpurun --gpus 2`), as a
`torch.distributed.run` child job (tools/peer_check.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from mansy_immersivevideostreaming_b200 import synth
from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, REWARD_QOE, SimConfig
from mansy_immersivevideostreaming_b200.rollout import PeerGroup, gather_episode_stats, summarise_stats
from mansy_immersivevideostreaming_b200.simulator import BatchSimulator, ViewportTiler
from mansy_immersivevideostreaming_b200.vector_env import B200VectorEnv

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = SimConfig()


def _tables(n):
    t = synth.make_synthetic_tables(ViewportTiler(CFG).chunk_masks, n_videos=3, n_users=4, n_traces=5, seed=21,
                                    trace_len_range=(40, 90))
    return t.with_samples(synth.per_env_samples(t, n))


def test_totals_pack_and_single_rank_peer_gather_equal_the_statistics_rows():
    n = 777
    sim = BatchSimulator(_tables(n), n, OBS_MODE_MANSY, REWARD_QOE, seed=2)
    sim.reset()
    sim.rollout_random(70, seed=5)                       # > 51 steps: every env has finished episodes
    full = sim.episode_stats()
    want = full[:, 6:12].contiguous()
    assert float(want[:, 5].min()) >= 1.0
    assert torch.equal(sim.episode_totals(), want)
    assert torch.equal(gather_episode_stats(sim), want)  # no process group: the pack kernel
    peers = PeerGroup(n, 0)
    for _ in range(3):                                   # both mailbox parities
        got = peers.gather_episode_stats(sim)
        assert got.shape == (n, 6) and got.dtype == torch.float64 and torch.equal(got, want)
    peers.barrier()
    torch.cuda.synchronize()
    assert not peers.timed_out()
    s = summarise_stats(got)
    assert s["episodes"] == float(want[:, 5].sum()) and s["steps"] == float(want[:, 4].sum())
    peers.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_rank_peer_gather_over_nvlink():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29731", os.path.join(ROOT, "tools", "peer_check.py")],
                       env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "peer_check ok" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


def test_env_ids_are_validated_on_the_host():
    n = 16
    t = _tables(n)
    sim = BatchSimulator(t, n, OBS_MODE_MANSY, REWARD_QOE, seed=0)
    sim.reset()
    acts = torch.zeros(2, dtype=torch.int32, device="cuda")
    with pytest.raises(IndexError):
        sim.step(acts, env_ids=[0, 16])
    with pytest.raises(IndexError):
        sim.reset([-1])
    with pytest.raises(ValueError):
        sim.step(acts, env_ids=[3, 3])
    venv = B200VectorEnv(t, n, OBS_MODE_MANSY, REWARD_QOE)
    venv.reset()
    with pytest.raises(IndexError):
        venv.step([1], id=[n])
    with pytest.raises(ValueError):
        venv.reset(id=[2, 2])
    obs, rew, done, info = venv.step([4, 5], id=[7, 2])
    assert rew.dtype == np.float64 and [i["env_id"] for i in info] == [7, 2]
    assert sim.error_flag() == 0
    # an empty id list is a call that touches nothing (tianshou passes one when no environment is ready)
    before = sim.episode_state_host().copy()
    rows, r0, d0 = sim.step(acts[:0], env_ids=[])
    assert rows.shape[0] == 0 and r0.numel() == 0 and d0.numel() == 0
    assert sim.reset([]).shape[0] == 0
    obs0, rew0, done0, info0 = venv.step([], id=[])
    assert len(rew0) == 0 and len(done0) == 0 and len(info0) == 0
    assert before.tobytes() == sim.episode_state_host().tobytes() and sim.error_flag() == 0


def test_seed_moves_the_cursor_without_touching_a_running_episode():
    """mansy_env.py:253-256: seed() only sets worker_id; an episode in flight keeps going."""
    n = 8
    t = _tables(n)
    a = BatchSimulator(t, n, OBS_MODE_MANSY, REWARD_QOE, seed=1)
    b = BatchSimulator(t, n, OBS_MODE_MANSY, REWARD_QOE, seed=1)
    a.reset(); b.reset()
    acts = torch.arange(n, dtype=torch.int32, device="cuda")
    a.seed(5)                                            # mid-episode
    oa, ra, da = a.step(acts)
    ob, rb, db = b.step(acts)
    assert torch.equal(oa, ob) and torch.equal(ra, rb) and not bool(da.any())
    sa, sb = a.episode_state_host(), b.episode_state_host()
    assert np.array_equal(sa["cursor"], (5 + np.arange(n)) % n) and not np.array_equal(sa["cursor"], sb["cursor"])
    assert np.array_equal(sa["next_chunk"], sb["next_chunk"])


def test_finished_env_stepped_again_is_logged_once(tmp_path):
    n = 4
    t = _tables(n)
    log = tmp_path / "log.csv"
    venv = B200VectorEnv(t, n, OBS_MODE_MANSY, REWARD_QOE, log_path=str(log))
    venv.reset()
    done = np.zeros(n, bool)
    steps = 0
    while not done.all():
        _, _, d, _ = venv.step(np.full(n, 3))
        done |= d
        steps += 1
        assert steps < 80
    for _ in range(3):                                   # finished envs stepped again before their reset
        _, rew, d, _ = venv.step(np.full(n, 3))
        assert d.all() and (rew == 0).all()
    venv.close()
    assert len(open(log).read().strip().splitlines()) == 1 + n


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_handles_run_on_their_own_device_whatever_the_current_one_is():
    """One process driving two GPUs: every entry point switches to its handle's device (the stream it is given belongs to
    it).  The same rollout on device 1 -- with device 0 current throughout -- must equal the one on device 0 bit for bit."""
    from mansy_immersivevideostreaming_b200.policy import PolicyNet, mansy_state_dict_shapes, seeded_state_dict
    from mansy_immersivevideostreaming_b200.rollout import PolicyRollout
    n, steps = 300, 60
    a, c = mansy_state_dict_shapes()
    out = []
    torch.cuda.set_device(0)
    for dev in (0, 1):
        t = synth.make_synthetic_tables(ViewportTiler(CFG, device=dev).chunk_masks, n_videos=3, n_users=4, n_traces=5, seed=21,
                                        trace_len_range=(40, 90))
        t = t.with_samples(synth.per_env_samples(t, n))
        sim = BatchSimulator(t, n, OBS_MODE_MANSY, REWARD_QOE, seed=3, device=dev)
        pol = PolicyNet(seeded_state_dict(a, 1), seeded_state_dict(c, 2), OBS_MODE_MANSY, device=dev)
        roll = PolicyRollout(sim, pol, 4, seed=77)
        assert torch.cuda.current_device() == 0
        roll.run(steps)
        torch.cuda.synchronize(dev)
        out.append((sim.episode_totals().cpu(), roll.buf.actions.cpu(), roll.buf.obs.cpu(), sim.error_flag()))
        peers = PeerGroup(n, dev)
        assert torch.equal(peers.gather_episode_stats(sim).cpu(), out[-1][0])
        peers.close(); sim.close(); pol.close()
    assert torch.cuda.current_device() == 0
    assert out[0][3] == 0 and out[1][3] == 0
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2])
    assert float(out[0][0][:, 5].sum()) > 0          # episodes finished: the comparison covers resets
