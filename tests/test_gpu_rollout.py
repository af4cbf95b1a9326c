"""Device rollout loop (policy forward -> sample -> step): the C-driven loop equals the same steps issued
one by one from Python, the host-storage variant equals the device ring, and the simulator inside the
rollout matches the oracle when the oracle is teacher-forced with the actions the device policy sampled."""
import numpy as np
import pytest
import torch

from mansy_immersivevideostreaming_b200 import synth
from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, OBS_MODE_SIMPLE, REWARD_QOE, SimConfig
from mansy_immersivevideostreaming_b200.policy import (PolicyNet, mansy_state_dict_shapes, seeded_state_dict,
                                                       simple_state_dict_shapes)
from mansy_immersivevideostreaming_b200.rollout import PolicyRollout
from mansy_immersivevideostreaming_b200.simulator import BatchSimulator, ViewportTiler
from oracle import sim_oracle as so

pytestmark = pytest.mark.gpu
CFG = SimConfig()


def _setup(kind, n, tensor_cores, slabs=5, seed=3):
    tables = synth.make_synthetic_tables(ViewportTiler(CFG).chunk_masks, n_videos=3, n_users=4, n_traces=5, seed=21,
                                         trace_len_range=(40, 90))
    tables = tables.with_samples(synth.per_env_samples(tables, n))
    shapes = mansy_state_dict_shapes() if kind == OBS_MODE_MANSY else simple_state_dict_shapes()
    policy = PolicyNet(seeded_state_dict(shapes[0], 1), seeded_state_dict(shapes[1], 2), kind)
    sim = BatchSimulator(tables, n, kind, REWARD_QOE, seed=seed)
    return tables, policy, PolicyRollout(sim, policy, slabs, seed=77, tensor_cores=tensor_cores)


@pytest.mark.parametrize("kind,tensor_cores", [(OBS_MODE_MANSY, True), (OBS_MODE_MANSY, False), (OBS_MODE_SIMPLE, True)])
def test_c_loop_equals_python_loop_and_oracle(kind, tensor_cores):
    n, steps = 160, 60                      # > 51 steps: every env finishes an episode and auto-resets
    tables, _, a = _setup(kind, n, tensor_cores, slabs=steps + 1)
    _, _, b = _setup(kind, n, tensor_cores, slabs=steps + 1)
    a.run(steps, timed=True)
    for _ in range(steps):
        b.step()
    torch.cuda.synchronize()
    assert torch.equal(a.buf.obs, b.buf.obs)
    for name in ("actions", "reward", "done", "value", "logp"):      # slab `steps` of the per-step rings is never written
        assert torch.equal(getattr(a.buf, name)[:steps], getattr(b.buf, name)[:steps]), name
    pm, sm, k = a.kernel_ms()
    assert k == steps and pm > 0 and sm > 0
    assert int(a.buf.done[:steps].sum()) >= n                       # auto-reset happened in the loop
    # teacher-forced oracle on the sampled actions
    orc = so.OracleVectorEnv(tables, n, kind, REWARD_QOE, chain="f64", seed=3)
    np.testing.assert_array_equal(a.buf.obs[0].cpu().numpy(), orc.reset())
    acts = a.buf.actions.cpu().numpy()
    for t in range(steps):
        oobs, orew, odone, _ = orc.step(acts[t], auto_reset=True)
        np.testing.assert_allclose(a.buf.obs[t + 1].cpu().numpy(), oobs, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(a.buf.reward[t].cpu().numpy(), orew, rtol=1e-5, atol=1e-6)
        assert np.array_equal(a.buf.done[t].cpu().numpy().astype(bool), odone)
    assert a.sim.error_flag() == 0


def test_host_rollout_equals_device_rollout():
    n, steps = 200, 12
    _, _, a = _setup(OBS_MODE_MANSY, n, True, slabs=4)
    _, _, b = _setup(OBS_MODE_MANSY, n, True, slabs=steps + 1)
    host = a.make_host_buffers(host_slabs=steps)
    a.run_host(steps, host)
    b.run(steps)
    torch.cuda.synchronize()
    for t in range(steps):
        assert torch.equal(host["obs"][t], b.buf.obs[t + 1].cpu())
        for name in ("actions", "reward", "done", "value", "logp"):
            assert torch.equal(host[name][t], getattr(b.buf, name)[t].cpu()), name
    h2d, d2h = a.host_bytes_per_step()
    assert h2d == 4 * n and d2h == n * (784 * 4 + 17)


@pytest.mark.parametrize("kind,n,steps,slabs", [(OBS_MODE_MANSY, 160, 60, 61), (OBS_MODE_MANSY, 4096, 12, 3),
                                                (OBS_MODE_SIMPLE, 333, 60, 7), (OBS_MODE_MANSY, 1, 55, 4),
                                                (OBS_MODE_MANSY, 8192 + 77, 7, 3),      # > 33 tiles: clusters walk several tiles per step
                                                (OBS_MODE_SIMPLE, 13000, 7, 3)])
def test_fused_rollout_equals_two_kernel_rollout(kind, n, steps, slabs, monkeypatch):
    """ONE launch of the fused policy+step cluster kernel for all steps == two launches per step (with and without
    programmatic dependent launch), bit for bit, including the ring wrap of the slabs and a continued rollout.  (More
    tiles than resident clusters: the kernel walks several tiles per cluster -- the default up to two tiles per cluster,
    forced here beyond that.)"""
    monkeypatch.setenv("MANSY_FUSED_MULTI_TILE", "1")
    _, _, a = _setup(kind, n, True, slabs=slabs)
    _, _, b = _setup(kind, n, True, slabs=slabs)
    _, _, c = _setup(kind, n, True, slabs=slabs)
    for r in (a, b, c):
        r.policy.set_tc_split(4)        # the same split-K cluster kernel on both sides (above 37 tiles the PDL loop would pick one CTA per tile,
                                        # whose accumulation order differs in the last bits)
    launches0 = a.sim.lib.mansy_kernel_launches()
    a.run(steps - 5)
    a.run(5)                               # continues at t = steps - 5
    torch.cuda.synchronize()
    assert a.sim.lib.mansy_kernel_launches() - launches0 == 3       # fused: one launch per call (+ the policy memo table, once)
    b.run(steps, fused=False)
    c.run(steps, fused=False, pdl=False)
    torch.cuda.synchronize()
    w = min(steps, slabs)                  # slabs the per-step rings have written (the buffers start uninitialised)
    for other in (b, c):
        assert torch.equal(a.buf.obs[:min(steps + 1, slabs)], other.buf.obs[:min(steps + 1, slabs)])
        for name in ("actions", "reward", "done", "value", "logp"):
            assert torch.equal(getattr(a.buf, name)[:w], getattr(other.buf, name)[:w]), name
        assert torch.equal(a.sim.episode_stats(), other.sim.episode_stats())
    assert a.sim.error_flag() == 0
