"""Tensor-core (tcgen05, TF32) policy forward against a plain PyTorch fp32 restatement of the same op
and against the reference model's own outputs (tests/golden/policy_kat.npz).

Tolerance: TF32 inputs (10-bit mantissa, the precision class the reference itself runs these layers in:
torch.set_float32_matmul_precision('high'), run_mansy.py:253) with fp32 accumulation.  Observed error on
logits of magnitude ~1 is a few 1e-4; the bound asserted here is 5e-3 absolute (+5e-3 relative).
Intermediate stages (layer-1 features, hidden activations) are checked first so that a failure names the
stage that broke."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import load_golden
from mansy_immersivevideostreaming_b200.config import (MANSY_OBS_SEGMENTS, OBS_MODE_MANSY, OBS_MODE_SIMPLE,
                                                       SIMPLE_OBS_SEGMENTS)
from mansy_immersivevideostreaming_b200.policy import (MANSY_BRANCHES, SIMPLE_BRANCHES, PolicyNet,
                                                       mansy_state_dict_shapes, seeded_state_dict,
                                                       simple_state_dict_shapes)
from test_gpu_policy import MANSY_KEYS, SIMPLE_KEYS, _shapes, torch_reference

pytestmark = pytest.mark.gpu
RTOL, ATOL = 5e-3, 5e-3


def stages_reference(rows, actor, critic, kind):
    segs = {k: (off, int(np.prod(shape))) for k, off, shape in (MANSY_OBS_SEGMENTS if kind == OBS_MODE_MANSY else SIMPLE_OBS_SEGMENTS)}
    names, keys = (MANSY_BRANCHES, MANSY_KEYS) if kind == OBS_MODE_MANSY else (SIMPLE_BRANCHES, SIMPLE_KEYS)
    x = torch.from_numpy(rows)
    feats = []
    for name, key in zip(names, keys):
        off, n = segs[key]
        w = torch.from_numpy(actor[f"feature_net.{name}.0.weight"]).reshape(128, -1)
        feats.append(F.leaky_relu(F.linear(x[:, off:off + n], w, torch.from_numpy(actor[f"feature_net.{name}.0.bias"])), 0.01))
    f = torch.cat(feats, dim=-1)
    res = feats[-1] if kind == OBS_MODE_MANSY else 0.0
    ha = F.leaky_relu(F.linear(f, torch.from_numpy(actor["fc.0.weight"]), torch.from_numpy(actor["fc.0.bias"])), 0.01) + res
    hc = F.leaky_relu(F.linear(f, torch.from_numpy(critic["fc.0.weight"]), torch.from_numpy(critic["fc.0.bias"])), 0.01) + res
    return [t.numpy() for t in feats], torch.cat([ha, hc], dim=1).numpy()


@pytest.mark.parametrize("split", [1, 4])
@pytest.mark.parametrize("kind,n", [(OBS_MODE_MANSY, 128), (OBS_MODE_MANSY, 1), (OBS_MODE_MANSY, 1000), (OBS_MODE_SIMPLE, 333),
                                    (OBS_MODE_MANSY, 4096), (OBS_MODE_MANSY, 148 * 128 * 2 + 77)])
def test_tc_policy_vs_torch_fp32(kind, n, split):
    shapes = mansy_state_dict_shapes() if kind == OBS_MODE_MANSY else simple_state_dict_shapes()
    actor, critic = seeded_state_dict(shapes[0], 7), seeded_state_dict(shapes[1], 8)
    stride = 784 if kind == OBS_MODE_MANSY else 400
    nb = 10 if kind == OBS_MODE_MANSY else 5
    rng = np.random.default_rng(n)
    rows = rng.random((n, stride)).astype(np.float32)
    net = PolicyNet(actor, critic, kind)
    net.set_tc_split(split)
    feat = torch.full((n, nb * 128), float("nan"), device="cuda")
    hid = torch.full((n, 256), float("nan"), device="cuda")
    logits, value, actions, logp = net.forward_tc(torch.from_numpy(rows).cuda(), seed=5, step=9, feat_dbg=feat, hid_dbg=hid)
    torch.cuda.synchronize()
    ref_feats, ref_hid = stages_reference(rows, actor, critic, kind)
    got_feat = feat.cpu().numpy()
    for i, canon in enumerate(PolicyNet.TC_BRANCH_ORDER[kind]):
        np.testing.assert_allclose(got_feat[:, i * 128:(i + 1) * 128], ref_feats[canon], rtol=RTOL, atol=ATOL,
                                   err_msg=f"layer-1 features of branch {canon} (processing slot {i})")
    if split == 4 and kind == OBS_MODE_MANSY:     # the cluster kernel applies the residual through the heads
        ref_hid = ref_hid - np.concatenate([ref_feats[-1], ref_feats[-1]], axis=1)
    np.testing.assert_allclose(hid.cpu().numpy(), ref_hid, rtol=RTOL, atol=ATOL, err_msg="hidden activations")
    ref_logits, ref_value = torch_reference(rows, actor, critic, kind)
    np.testing.assert_allclose(logits.cpu().numpy()[:, :15], ref_logits, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(value.cpu().numpy(), ref_value, rtol=RTOL, atol=ATOL)
    # the fused sampler == the stand-alone sampler on the same logits
    a2, lp2 = net.sample(logits, seed=5, step=9)
    assert torch.equal(actions, a2)
    np.testing.assert_allclose(logp.cpu().numpy(), lp2.cpu().numpy(), rtol=1e-6, atol=1e-6)
    # fp32 CUDA-core kernel vs tensor-core kernel
    l32, v32 = net.forward(torch.from_numpy(rows).cuda())
    np.testing.assert_allclose(logits.cpu().numpy()[:, :15], l32.cpu().numpy()[:, :15], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("split", [0, 1, 4])
def test_tc_policy_vs_reference_golden(split):
    g = load_golden("policy_kat.npz")
    actor = seeded_state_dict(_shapes(g["actor_names"], g["actor_shapes"]), 101)
    critic = seeded_state_dict([(n, s) for n, s in _shapes(g["critic_names"], g["critic_shapes"])
                                if not n.startswith("feature_net.")], 102)
    net = PolicyNet(actor, critic, OBS_MODE_MANSY)
    net.set_tc_split(split)
    rows = np.ascontiguousarray(g["mansy_rows"])
    logits, value, _, _ = net.forward_tc(torch.from_numpy(rows).cuda(), sample=False)
    np.testing.assert_allclose(logits.cpu().numpy()[:, :15], g["actor_logits"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(value.cpu().numpy(), g["critic_value"].reshape(-1), rtol=RTOL, atol=ATOL)


def test_tc_policy_strided_rows_and_repeat_calls():
    """Rows inside a wider buffer (obs_stride > 784) and back-to-back launches on the same handle."""
    shapes = mansy_state_dict_shapes()
    actor, critic = seeded_state_dict(shapes[0], 3), seeded_state_dict(shapes[1], 4)
    net = PolicyNet(actor, critic, OBS_MODE_MANSY)
    rng = np.random.default_rng(0)
    wide = rng.random((300, 800)).astype(np.float32)
    dev = torch.from_numpy(wide).cuda()
    ref_logits, _ = torch_reference(np.ascontiguousarray(wide[:, :784]), actor, critic, OBS_MODE_MANSY)
    for _ in range(3):
        logits, _, _, _ = net.forward_tc(dev[:, :784], sample=False)
        np.testing.assert_allclose(logits.cpu().numpy()[:, :15], ref_logits, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("split", [1, 4])
@pytest.mark.parametrize("kind,n", [(OBS_MODE_MANSY, 700), (OBS_MODE_SIMPLE, 333), (OBS_MODE_MANSY, 4096)])
def test_memoised_table_branches_equal_the_full_forward(kind, n, split):
    """``forward_tc_sim`` (what every rollout runs): the 320-input branches come from the (video, chunk) memo computed in
    exact fp32.  On the simulator's own observations -- mid-episode, after auto-resets, different videos / chunks per env --
    it must agree with the torch fp32 restatement (same 5e-3 bound, in practice tighter than the all-TF32 forward) and with
    the full tensor-core forward of the same rows."""
    from mansy_immersivevideostreaming_b200 import synth
    from mansy_immersivevideostreaming_b200.config import REWARD_QOE, SimConfig
    from mansy_immersivevideostreaming_b200.simulator import BatchSimulator, ViewportTiler
    shapes = mansy_state_dict_shapes() if kind == OBS_MODE_MANSY else simple_state_dict_shapes()
    actor, critic = seeded_state_dict(shapes[0], 17), seeded_state_dict(shapes[1], 18)
    net = PolicyNet(actor, critic, kind)
    net.set_tc_split(split)
    t = synth.make_synthetic_tables(ViewportTiler(SimConfig()).chunk_masks, n_videos=5, n_users=6, n_traces=7, seed=33,
                                    trace_len_range=(40, 90))
    t = t.with_samples(synth.per_env_samples(t, n))
    sim = BatchSimulator(t, n, kind, REWARD_QOE, seed=4)
    obs = sim.new_obs()
    sim.reset(None, obs)
    for steps in (0, 9, 61):                          # reset rows, mid-episode, after every env wrapped into a new episode
        if steps:
            sim.rollout_random(steps, seed=steps, obs=obs)
        rows = obs.cpu().numpy()
        ref_logits, ref_value = torch_reference(rows, actor, critic, kind)
        lg = torch.empty((n, 16), device="cuda"); va = torch.empty(n, device="cuda")
        ac = torch.empty(n, dtype=torch.int32, device="cuda"); lp = torch.empty(n, device="cuda")
        net.forward_tc_sim(sim, obs, lg, va, ac, lp, seed=3, step=steps)
        np.testing.assert_allclose(lg.cpu().numpy()[:, :15], ref_logits, rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(va.cpu().numpy(), ref_value, rtol=RTOL, atol=ATOL)
        full_l, full_v, _, _ = net.forward_tc(obs, sample=False)
        np.testing.assert_allclose(lg.cpu().numpy()[:, :15], full_l.cpu().numpy()[:, :15], rtol=RTOL, atol=ATOL)
        a2, _ = net.sample(lg, seed=3, step=steps)
        assert torch.equal(ac, a2)
    # a second simulator with other tables: the memo follows the simulator it is asked about
    t2 = synth.make_synthetic_tables(ViewportTiler(SimConfig()).chunk_masks, n_videos=3, n_users=4, n_traces=5, seed=99,
                                     trace_len_range=(40, 90))
    sim2 = BatchSimulator(t2.with_samples(synth.per_env_samples(t2, n)), n, kind, REWARD_QOE, seed=1)
    obs2 = sim2.new_obs()
    sim2.reset(None, obs2)
    sim2.rollout_random(5, seed=2, obs=obs2)
    ref_logits, _ = torch_reference(obs2.cpu().numpy(), actor, critic, kind)
    net.forward_tc_sim(sim2, obs2, lg, va, ac, lp)
    np.testing.assert_allclose(lg.cpu().numpy()[:, :15], ref_logits, rtol=RTOL, atol=ATOL)
    net.forward_tc_sim(sim, obs, lg, va, ac, lp)
    np.testing.assert_allclose(lg.cpu().numpy()[:, :15], torch_reference(obs.cpu().numpy(), actor, critic, kind)[0], rtol=RTOL, atol=ATOL)
