"""Policy / value forward (models/mansy.py, models/simple_rl.py) on the GPU against
(1) the reference model's own outputs stored in tests/golden/policy_kat.npz (weights regenerated
from the recorded numpy seeds) and (2) a plain PyTorch fp32 restatement of the same op.

Tolerance: fp32 with a different summation order than torch's GEMM -> rtol 1e-4, atol 2e-5
(the reference itself runs these layers in TF32, run_mansy.py:253, which is far coarser)."""
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

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 2e-5

MANSY_KEYS = ("throughput", "next_chunk_size", "next_chunk_quality", "pred_viewport", "viewport_acc",
              "past_viewport_qualities", "past_quality_variances", "past_rebuffering", "buffer", "qoe_weight")
SIMPLE_KEYS = ("throughput", "chunk_sizes", "rebuffer", "last_bitrates", "pred_viewport")


def torch_reference(rows, actor, critic, kind):
    """Plain fp32 PyTorch restatement (CPU, float32 'highest' precision)."""
    segs = {k: (off, int(np.prod(shape))) for k, off, shape in (MANSY_OBS_SEGMENTS if kind == OBS_MODE_MANSY else SIMPLE_OBS_SEGMENTS)}
    names, keys = (MANSY_BRANCHES, MANSY_KEYS) if kind == OBS_MODE_MANSY else (SIMPLE_BRANCHES, SIMPLE_KEYS)
    x = torch.from_numpy(rows)
    feats = []
    for name, key in zip(names, keys):
        off, n = segs[key]
        w = torch.from_numpy(actor[f"feature_net.{name}.0.weight"]).reshape(128, -1)
        b = torch.from_numpy(actor[f"feature_net.{name}.0.bias"])
        feats.append(F.leaky_relu(F.linear(x[:, off:off + n], w, b), 0.01))
    f = torch.cat(feats, dim=-1)
    res = feats[-1] if kind == OBS_MODE_MANSY else 0.0
    ha = F.leaky_relu(F.linear(f, torch.from_numpy(actor["fc.0.weight"]), torch.from_numpy(actor["fc.0.bias"])), 0.01) + res
    hc = F.leaky_relu(F.linear(f, torch.from_numpy(critic["fc.0.weight"]), torch.from_numpy(critic["fc.0.bias"])), 0.01) + res
    logits = F.linear(ha, torch.from_numpy(actor["out.weight"]), torch.from_numpy(actor["out.bias"]))
    value = F.linear(hc, torch.from_numpy(critic["out.weight"]), torch.from_numpy(critic["out.bias"]))
    if kind == OBS_MODE_SIMPLE:
        logits = torch.softmax(logits, dim=1)
    return logits.numpy(), value.numpy().reshape(-1)


def _shapes(names, shapes):
    return [(str(n), tuple(int(x) for x in str(s).split(","))) for n, s in zip(names, shapes)]


def test_mansy_policy_vs_reference_golden():
    g = load_golden("policy_kat.npz")
    actor = seeded_state_dict(_shapes(g["actor_names"], g["actor_shapes"]), 101)
    critic_all = _shapes(g["critic_names"], g["critic_shapes"])
    critic = seeded_state_dict([(n, s) for n, s in critic_all if not n.startswith("feature_net.")], 102)
    net = PolicyNet(actor, critic, OBS_MODE_MANSY)
    rows = np.ascontiguousarray(g["mansy_rows"])
    logits, value = net.forward(torch.from_numpy(rows).cuda())
    np.testing.assert_allclose(logits.cpu().numpy()[:, :15], g["actor_logits"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(value.cpu().numpy(), g["critic_value"].reshape(-1), rtol=RTOL, atol=ATOL)
    ref_logits, ref_value = torch_reference(rows, actor, critic, OBS_MODE_MANSY)
    np.testing.assert_allclose(ref_logits, g["actor_logits"], rtol=1e-5, atol=1e-5)     # restatement == reference model


def test_simple_policy_vs_reference_golden():
    g = load_golden("policy_kat.npz")
    actor = seeded_state_dict(_shapes(g["simple_actor_names"], g["simple_actor_shapes"]), 201)
    critic = seeded_state_dict([(n, s) for n, s in _shapes(g["simple_critic_names"], g["simple_critic_shapes"])
                                if not n.startswith("feature_net.")], 202)
    net = PolicyNet(actor, critic, OBS_MODE_SIMPLE)
    rows = np.ascontiguousarray(g["simple_rows"])
    probs, value = net.forward(torch.from_numpy(rows).cuda())
    np.testing.assert_allclose(probs.cpu().numpy()[:, :15], g["simple_probs"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(value.cpu().numpy(), g["simple_value"].reshape(-1), rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("kind,n", [(OBS_MODE_MANSY, 1), (OBS_MODE_MANSY, 1000), (OBS_MODE_SIMPLE, 333), (OBS_MODE_MANSY, 4096)])
def test_policy_vs_torch_fp32(kind, n):
    shapes = mansy_state_dict_shapes() if kind == OBS_MODE_MANSY else simple_state_dict_shapes()
    actor, critic = seeded_state_dict(shapes[0], 7), seeded_state_dict(shapes[1], 8)
    stride = 784 if kind == OBS_MODE_MANSY else 400
    rng = np.random.default_rng(n)
    rows = rng.random((n, stride)).astype(np.float32)
    net = PolicyNet(actor, critic, kind)
    logits, value = net.forward(torch.from_numpy(rows).cuda())
    ref_logits, ref_value = torch_reference(rows, actor, critic, kind)
    np.testing.assert_allclose(logits.cpu().numpy()[:, :15], ref_logits, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(value.cpu().numpy(), ref_value, rtol=RTOL, atol=ATOL)


def test_categorical_sampling():
    shapes = mansy_state_dict_shapes()
    net = PolicyNet(seeded_state_dict(shapes[0], 1), seeded_state_dict(shapes[1], 2), OBS_MODE_MANSY)
    n = 200_000
    logits = torch.zeros((n, 16), device="cuda")
    logits[:, :15] = torch.linspace(-2, 2, 15, device="cuda")[None, :]
    a1, lp1 = net.sample(logits, seed=11, step=3)
    a2, _ = net.sample(logits, seed=11, step=3)
    a3, _ = net.sample(logits, seed=11, step=4)
    assert torch.equal(a1, a2) and not torch.equal(a1, a3)          # counter-based: reproducible, step-dependent
    p = torch.softmax(logits[0, :15], dim=0)
    freq = torch.bincount(a1.long(), minlength=15).float() / n
    assert torch.all((freq - p).abs() < 5 * torch.sqrt(p * (1 - p) / n) + 1e-4)
    np.testing.assert_allclose(lp1.cpu().numpy(), torch.log(p)[a1.long()].cpu().numpy(), rtol=1e-5, atol=1e-6)
