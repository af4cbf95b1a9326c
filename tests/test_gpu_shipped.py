"""Fixtures the reference itself ships, read from the oracle/_ref archive on the GPU box:

* ``best_policy.pth`` / ``best_identifier.pth`` -- real trained weights through ``PolicyNet.from_policy_state_dict`` /
  ``IdentifierNet`` against the outputs of the unmodified reference ``Actor`` / ``Critic`` / ``QoEIdentifier`` on real-data
  observation rows (``policy_shipped_kat.npz``, made by oracle/make_golden_shipped.py);
* ``valid_log.csv`` / ``train_log.csv`` -- the episode ORDER the unmodified scripts produce, which pins the
  "every reset() call advances the sample cursor" rule on ``B200VectorEnv``.
"""
from __future__ import annotations

import io

import numpy as np
import pytest
import torch

from helpers import load_golden
from oracle import ref_loader

pytestmark = pytest.mark.gpu

FP32_ATOL = 2e-5      # exact-fp32 CUDA-core kernels vs torch fp32 on the CPU
TF32_ATOL = 5e-3      # tcgen05 kind::tf32 (the reference's own precision class, run_mansy.py:253)


def _fixture(name: str) -> bytes:
    if not ref_loader.code_available():
        pytest.fail("oracle/_ref/mansy_reference.zip is missing: run `python -m oracle.make_ref` where /root/reference exists "
                    "(build() does it) -- the archive travels with gpurun")
    return ref_loader.read_member(f"fixtures/{name}")


def test_shipped_policy_checkpoint_fp32_and_tf32():
    from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY
    from mansy_immersivevideostreaming_b200.policy import PolicyNet
    g = load_golden("policy_shipped_kat.npz")
    sd = torch.load(io.BytesIO(_fixture("best_policy.pth")), map_location="cpu", weights_only=False)
    net = PolicyNet.from_policy_state_dict(sd, OBS_MODE_MANSY)
    obs = torch.from_numpy(g["rows"]).cuda()
    logits, value = net.forward(obs)
    np.testing.assert_allclose(logits.cpu().numpy()[:, :15], g["actor_logits"], rtol=0, atol=FP32_ATOL)
    np.testing.assert_allclose(value.cpu().numpy(), g["critic_value"].reshape(-1), rtol=0, atol=FP32_ATOL)
    for split in (1, 4):                       # one CTA per tile / split-K 4-CTA cluster (the fused rollout's kernel)
        net.set_tc_split(split)
        tl, tv, _, _ = net.forward_tc(obs, sample=False)
        np.testing.assert_allclose(tl.cpu().numpy()[:, :15], g["actor_logits"], rtol=0, atol=TF32_ATOL)
        np.testing.assert_allclose(tv.cpu().numpy(), g["critic_value"].reshape(-1), rtol=0, atol=TF32_ATOL)
    # the trained policy is far from uniform on these rows: the comparison is not vacuous
    assert np.abs(g["actor_logits"]).max() > 0.5 and np.ptp(g["critic_value"]) > 0.5


def test_shipped_identifier_checkpoint():
    from mansy_immersivevideostreaming_b200.policy import IdentifierNet
    g = load_golden("policy_shipped_kat.npz")
    obs = torch.from_numpy(g["rows"]).cuda()
    isd = torch.load(io.BytesIO(_fixture("best_identifier.pth")), map_location="cpu", weights_only=False)
    ident = IdentifierNet(isd)
    np.testing.assert_allclose(ident.forward(obs, tensor_cores=False).cpu().numpy(), g["ident_out"], rtol=0, atol=FP32_ATOL)
    np.testing.assert_allclose(ident.forward(obs, tensor_cores=True).cpu().numpy(), g["ident_out"], rtol=0, atol=TF32_ATOL)
    # the copy saved inside best_policy.pth (identifier.*)
    sd = torch.load(io.BytesIO(_fixture("best_policy.pth")), map_location="cpu", weights_only=False)
    ident2 = IdentifierNet({k[len("identifier."):]: v for k, v in sd.items() if k.startswith("identifier.")})
    np.testing.assert_allclose(ident2.forward(obs, tensor_cores=False).cpu().numpy(), g["ident_in_policy_out"], rtol=0, atol=FP32_ATOL)


def _log_samples(text: str, videos, users, traces, qoe):
    """Map the rows of a shipped episode log to sample indices of generate_environment_samples (utils/common.py:60-84)."""
    V, U, T, Q = len(videos), len(users), len(traces), len(qoe)
    max_len = max(V, U, T, Q)
    total = max(max_len, V * Q * (-(-max_len // (V * Q))))
    lookup = {(videos[i % V], users[i % U], traces[i % T], tuple(float(x) for x in qoe[i % Q])): i for i in range(total)}
    assert len(lookup) == total
    out = []
    for line in text.strip().splitlines()[1:]:
        f = line.split(",")
        out.append(lookup[(int(f[0]), int(f[1]), int(f[2]), (float(f[3]), float(f[4]), float(f[5])))])
    return out, total


def test_shipped_logs_pin_the_cursor_per_reset_rule(tmp_path):
    """``valid_log.csv`` of the shipped run starts at sample 5 because tianshou's Collector resets every env once when it
    is built and once more when a validation pass starts, and EVERY reset consumes a sample (mansy_env.py:100-101);
    ``train_log.csv`` starts at sample 0.  Replay that protocol on ``B200VectorEnv`` (synthetic tables carrying the
    shipped split's ids; equal-length episodes so the lock-step order is the env order) and compare, per environment,
    the sequence of samples with the one the shipped log shows for that environment."""
    import dataclasses
    import yaml
    from mansy_immersivevideostreaming_b200 import synth
    from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, REWARD_QOE, SimConfig
    from mansy_immersivevideostreaming_b200.tables import environment_samples
    from mansy_immersivevideostreaming_b200.vector_env import B200VectorEnv
    from oracle import sim_oracle as so
    cfg = yaml.safe_load(ref_loader.read_member("config.yml"))
    scfg = SimConfig()

    def split(mode):
        return (cfg["video_split"]["Jin2022"][mode], cfg["user_split"]["Jin2022"][mode], cfg["network_split"]["4G"][mode],
                cfg["qoe_split"][mode])

    def replay(mode, n_envs, seed, resets_before_play, n_rounds, log):
        videos, users, traces, qoe = split(mode)
        V, U, T, Q = len(videos), len(users), len(traces), len(qoe)
        t = synth.make_synthetic_tables(lambda g, p: so.chunk_masks(g, p, scfg), n_videos=V, n_users=U, n_chunks=20, n_traces=T,
                                        seed=5, trace_len_range=(30, 50), short_tail_frac=0.0, vp_last_chunk=19)
        t = dataclasses.replace(t, video_ids=np.asarray(videos), user_ids=np.asarray(users), trace_ids=np.asarray(traces),
                                qoe_w=np.asarray(qoe, dtype=np.float32)).with_samples(environment_samples(V, U, T, Q))
        venv = B200VectorEnv(t, n_envs, OBS_MODE_MANSY, REWARD_QOE, seed=0, worker_num=n_envs, log_path=str(log))
        venv.seed(seed)                                        # run_mansy.py:55-56
        for _ in range(resets_before_play):
            venv.reset()
        rng = np.random.default_rng(0)
        for _ in range(n_rounds):
            done = np.zeros(n_envs, bool)
            for _ in range(t.n_chunks):                        # a finished env stepped again stays finished (logged once)
                _, _, d, _ = venv.step(rng.integers(0, 15, size=n_envs))
                done |= d
            assert done.all()
            venv.reset(np.flatnonzero(done))                   # Collector: reset the finished ids
        venv.close()
        return _log_samples(open(log).read(), videos, users, traces, qoe)[0]

    valid_ids, total = _log_samples(_fixture("valid_log.csv").decode(), *split("valid"))
    train_ids, total_train = _log_samples(_fixture("train_log.csv").decode(), *split("train"))
    assert (total, total_train) == (48, 72) and valid_ids[:4] == [5, 6, 7, 4] and train_ids[:3] == [0, 1, 2]

    # validation: 4 envs (one per QoE weight), worker_num 4, seed 5; Collector.__init__ reset + reset_env of the pass
    ours = replay("valid", 4, 5, 2, 5, tmp_path / "valid.csv")
    for k in range(4):
        res = (5 + k) % 4                                      # env k keeps the residue of its first worker_id
        shipped_k = [s for s in valid_ids if s % 4 == res]
        ours_k = ours[k::4]
        assert len(ours_k) == 5 and ours_k == shipped_k[:5], (k, ours_k, shipped_k[:6])
    assert ours[:8] == valid_ids[:8] == [5, 6, 7, 4, 9, 10, 11, 8]
    # training: 1 env, worker_num 1 (run_mansy.py:37), seed 5: the Collector's one reset starts sample 0
    ours = replay("train", 1, 5, 1, 6, tmp_path / "train.csv")
    assert ours == train_ids[:6] == [0, 1, 2, 3, 4, 5]
