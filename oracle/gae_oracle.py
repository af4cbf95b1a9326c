"""CPU restatement of the PPO learner inputs the rollout feeds (TEST INFRASTRUCTURE: only tests/, smoke() and
bench.py's cpu_baseline may import anything under oracle/).

``gae``: tianshou==0.4.8 (pinned by the reference's README.md:21; NOT vendored under /root/reference and not
installed here, so **parity unpinned** at this boundary) ``BasePolicy.compute_episodic_return`` with
``_gae_return``, restated from the published algorithm:

    v_s_  = V(s_{t+1}) * value_mask            (0 after a terminal transition)
    delta = rew + v_s_ * gamma - v_s
    discount = (1 - end_flag) * (gamma * gae_lambda)
    gae_t = delta_t + discount_t * gae_{t+1}    (backwards, float64)
    advantage = gae,  returns = gae + v_s       (cast to float32 by to_torch_as)

call sites in the reference: PPOPolicy.process_fn through ``models/mansy_ppo.py:55`` (``self.process_fn``),
``run_mansy.py:240-251`` (gamma, gae_lambda).

``identifier_reward``: ``utils/mansy_utils.py:42-49`` + ``models/mansy_ppo.py:40-49`` in numpy float32 / float64.
"""
import numpy as np


def gae(reward, value, done, last_value, gamma, lam):
    """reward, value, done: [T, N]; last_value: [N].  Returns (advantage, returns) float32 [T, N]."""
    rew = np.asarray(reward, dtype=np.float64)
    v_s = np.asarray(value, dtype=np.float64)
    live = 1.0 - np.asarray(done, dtype=np.float64)
    v_next = np.concatenate([v_s[1:], np.asarray(last_value, dtype=np.float64)[None]], axis=0) * live
    delta = rew + v_next * gamma - v_s
    discount = live * (gamma * lam)
    out = np.zeros_like(rew)
    g = np.zeros(rew.shape[1])
    for t in range(rew.shape[0] - 1, -1, -1):
        g = delta[t] + discount[t] * g
        out[t] = g
    return out.astype(np.float32), (out + v_s).astype(np.float32)


def identifier_reward(pred, qoe_weight, qoe_reward, lamb):
    """pred, qoe_weight: [N, 3] float32; qoe_reward [N].  Returns (identifier reward float32 [N], blended float64 [N])."""
    d = (np.asarray(pred, np.float32) - np.asarray(qoe_weight, np.float32)).astype(np.float32)
    sq = (d * d).astype(np.float32)
    mse = ((sq[:, 0] + sq[:, 1]).astype(np.float32) + sq[:, 2]).astype(np.float32) / np.float32(3.0)
    ident = (np.float32(1.0) - mse).astype(np.float32)
    mixed = (1.0 - lamb) * np.asarray(qoe_reward, np.float64) + lamb * ident.astype(np.float64)
    return ident, mixed
