"""Policy / value forward on packed observation rows (CUDA, through the C ABI).

Mirrors the reference's ``FeatureNet`` + ``Actor`` + ``Critic`` forward
(bitrate_selection/models/mansy.py:26-51,63-66,77-80) and the SimpleRL baseline nets
(bitrate_selection/models/simple_rl.py:21-35,46-49,60-63).  Weights come from the reference's
own state-dict layout, so a checkpoint written by ``run_mansy.py`` (``best_policy.pth``: keys
``actor.*`` / ``critic.*``) loads unchanged.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Mapping, Optional, Tuple

import numpy as np
import torch

from . import _capi
from ._capi import MansyError, PolicyWeights, check
from .config import OBS_MODE_MANSY, OBS_MODE_SIMPLE
from .simulator import _require_cuda

# FeatureNet branch order (concat order of models/mansy.py:39-50 / models/simple_rl.py:28-34)
MANSY_BRANCHES = ("conv1d1", "conv1d2", "conv1d3", "conv1d4", "conv1d5", "conv1d6", "conv1d7", "conv1d8", "fc1", "fc2")
MANSY_BRANCH_K = (8, 320, 320, 64, 8, 8, 8, 8, 1, 3)
SIMPLE_BRANCHES = ("conv1d_1", "conv1d_2", "fc1", "fc2", "fc3")
SIMPLE_BRANCH_K = (8, 320, 1, 2, 64)


def _arr(x) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


IDENT_BRANCH_K = (8, 320, 320, 64, 8, 8, 8, 8, 1, 15)      # QoEIdentifierFeatureNet: fc2 reads action_one_hot (mansy.py:101)
NET_IDENTIFIER = 3                                          # MANSY_NET_IDENTIFIER (include/mansy_b200.h)


class PolicyNet:
    """Actor + critic sharing one FeatureNet, evaluated in one kernel."""

    def __init__(self, actor_sd: Mapping[str, object], critic_sd: Mapping[str, object], kind: int = OBS_MODE_MANSY,
                 device: int = 0):
        _require_cuda(device)
        self.lib = _capi.load_library()
        self.kind = kind
        self.device = torch.device("cuda", device)
        names, ks = (MANSY_BRANCHES, MANSY_BRANCH_K) if kind == OBS_MODE_MANSY else (SIMPLE_BRANCHES, SIMPLE_BRANCH_K)
        feat = 128 * len(names)
        keep = []
        w = PolicyWeights()
        w.kind = kind
        for i, (name, k) in enumerate(zip(names, ks)):
            wt = _arr(actor_sd[f"feature_net.{name}.0.weight"]).reshape(128, -1)
            if wt.shape[1] != k:
                raise ValueError(f"feature_net.{name}: expected {k} inputs, got {wt.shape[1]}")
            bs = _arr(actor_sd[f"feature_net.{name}.0.bias"]).reshape(128)
            keep += [wt, bs]
            w.branch_w[i], w.branch_b[i] = wt.ctypes.data, bs.ctypes.data

        def put(field, arr, shape):
            a = _arr(arr).reshape(shape)
            keep.append(a)
            setattr(w, field, a.ctypes.data)

        put("actor_fc_w", actor_sd["fc.0.weight"], (128, feat))
        put("actor_fc_b", actor_sd["fc.0.bias"], (128,))
        put("actor_out_w", actor_sd["out.weight"], (15, 128))
        put("actor_out_b", actor_sd["out.bias"], (15,))
        put("critic_fc_w", critic_sd["fc.0.weight"], (128, feat))
        put("critic_fc_b", critic_sd["fc.0.bias"], (128,))
        put("critic_out_w", critic_sd["out.weight"], (1, 128))
        put("critic_out_b", critic_sd["out.bias"], (1,))
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.mansy_policy_create(C.byref(w), device, C.byref(h)))
        self._h = h

    @classmethod
    def from_policy_state_dict(cls, sd: Mapping[str, object], kind: int = OBS_MODE_MANSY, device: int = 0) -> "PolicyNet":
        """From a tianshou policy state dict (``actor.*`` / ``critic.*`` keys, run_mansy.py:96-104)."""
        actor = {k[len("actor."):]: v for k, v in sd.items() if k.startswith("actor.")}
        critic = {k[len("critic."):]: v for k, v in sd.items() if k.startswith("critic.")}
        if not actor or not critic:
            raise ValueError("state dict has no actor.* / critic.* entries")
        return cls(actor, critic, kind, device)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.mansy_policy_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def forward(self, obs: torch.Tensor, logits: Optional[torch.Tensor] = None,
                value: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """obs ``[N, stride]`` float32 on the device -> (logits ``[N, 16]`` (15 used), value ``[N]``).
        SimpleRL returns action probabilities like the reference's Actor (simple_rl.py:48)."""
        if obs.device != self.device or obs.dtype != torch.float32 or obs.stride(-1) != 1:
            raise ValueError("obs must be a float32 row-major tensor on the policy's device")
        n = obs.shape[0]
        if logits is None:
            logits = torch.empty((n, 16), dtype=torch.float32, device=self.device)
        if value is None:
            value = torch.empty(n, dtype=torch.float32, device=self.device)
        check(self.lib.mansy_policy_forward(self._h, obs.data_ptr(), obs.stride(0), n, logits.data_ptr(),
                                            value.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream))
        return logits, value

    # processing order of the FeatureNet branches inside the tensor-core kernel (csrc/mansy_policy_tc.cu):
    # position i of ``feat_dbg`` holds branch TC_BRANCH_ORDER[kind][i] of the concat order
    TC_BRANCH_ORDER = {OBS_MODE_MANSY: (1, 2, 3, 0, 4, 5, 6, 7, 8, 9), OBS_MODE_SIMPLE: (1, 4, 0, 2, 3)}

    def set_tc_split(self, split: int) -> None:
        """0 = choose from the batch size (default), 1 = one CTA per 128-env tile, 4 = split-K cluster of 4 CTAs per
        tile (``mansy_policy_tc_set_split``; in split mode ``hid_dbg`` excludes the residual)."""
        check(self.lib.mansy_policy_tc_set_split(self._h, int(split)))

    def forward_tc(self, obs: torch.Tensor, logits: Optional[torch.Tensor] = None, value: Optional[torch.Tensor] = None,
                   actions: Optional[torch.Tensor] = None, logp: Optional[torch.Tensor] = None, seed: int = 0,
                   step: int = 0, env_offset: int = 0, sample: bool = True, feat_dbg: Optional[torch.Tensor] = None,
                   hid_dbg: Optional[torch.Tensor] = None):
        """Tensor-core (tcgen05, TF32) forward + categorical sample in one launch.
        Returns (logits ``[N, 16]``, value ``[N]``, actions ``[N]`` int32 | None, logp ``[N]`` | None)."""
        if obs.device != self.device or obs.dtype != torch.float32 or obs.stride(-1) != 1:
            raise ValueError("obs must be a float32 row-major tensor on the policy's device")
        n = obs.shape[0]
        if logits is None:
            logits = torch.empty((n, 16), dtype=torch.float32, device=self.device)
        if value is None:
            value = torch.empty(n, dtype=torch.float32, device=self.device)
        if sample and actions is None:
            actions = torch.empty(n, dtype=torch.int32, device=self.device)
        if sample and logp is None:
            logp = torch.empty(n, dtype=torch.float32, device=self.device)
        ptr = lambda t: None if t is None else t.data_ptr()   # noqa: E731
        check(self.lib.mansy_policy_forward_tc(self._h, obs.data_ptr(), obs.stride(0), n, logits.data_ptr(),
                                               value.data_ptr(), ptr(actions), ptr(logp), int(seed), int(step),
                                               int(env_offset), ptr(feat_dbg), ptr(hid_dbg),
                                               torch.cuda.current_stream(self.device).cuda_stream))
        return logits, value, actions, logp

    def forward_tc_sim(self, sim, obs: torch.Tensor, logits: torch.Tensor, value: torch.Tensor, actions: torch.Tensor,
                       logp: torch.Tensor, seed: int = 0, step: int = 0):
        """``forward_tc`` on the CURRENT observations of ``sim`` (row i = env i): the 320-input table branches come from the
        (video, chunk) memo (``mansy_policy_forward_tc_sim``) -- the path every rollout entry point takes."""
        if obs.device != self.device or obs.dtype != torch.float32 or obs.stride(-1) != 1 or obs.shape[0] != sim.n_envs:
            raise ValueError("obs must hold one float32 row per environment of the simulator, on the policy's device")
        check(self.lib.mansy_policy_forward_tc_sim(self._h, sim._h, obs.data_ptr(), obs.stride(0), logits.data_ptr(), value.data_ptr(),
                                                   actions.data_ptr(), logp.data_ptr(), int(seed), int(step),
                                                   torch.cuda.current_stream(self.device).cuda_stream))
        return logits, value, actions, logp

    def sample(self, logits: torch.Tensor, seed: int, step: int, env_offset: int = 0,
               actions: Optional[torch.Tensor] = None, logp: Optional[torch.Tensor] = None):
        """Categorical(logits).sample() (run_mansy.py:228-229) with a counter-based generator."""
        n = logits.shape[0]
        if actions is None:
            actions = torch.empty(n, dtype=torch.int32, device=self.device)
        if logp is None:
            logp = torch.empty(n, dtype=torch.float32, device=self.device)
        check(self.lib.mansy_policy_sample(logits.data_ptr(), n, 0 if self.kind == OBS_MODE_MANSY else 1, int(seed),
                                           int(step), int(env_offset), actions.data_ptr(), logp.data_ptr(),
                                           torch.cuda.current_stream(self.device).cuda_stream))
        return actions, logp


class IdentifierNet:
    """The QoE identifier (``QoEIdentifierFeatureNet`` + ``QoEIdentifier``, models/mansy.py:83-143) on packed MANSY
    observation rows: ``pred = sigmoid(out(fc(features) + fc2(action_one_hot)))`` -- the rows already carry the last
    action's one-hot (``mansy_env.py:243``), which is what ``PPOPolicy.update`` feeds it (``mansy_ppo.py:44-47``).
    Runs on the same kernels as the policy (tcgen05 TF32 by default, exact fp32 with ``tensor_cores=False``)."""

    def __init__(self, sd: Mapping[str, object], device: int = 0):
        _require_cuda(device)
        self.lib = _capi.load_library()
        self.device = torch.device("cuda", device)
        keep = []
        w = PolicyWeights()
        w.kind = NET_IDENTIFIER
        for i, (name, k) in enumerate(zip(MANSY_BRANCHES, IDENT_BRANCH_K)):
            wt = _arr(sd[f"feature_net.{name}.0.weight"]).reshape(128, -1)
            if wt.shape[1] != k:
                raise ValueError(f"feature_net.{name}: expected {k} inputs, got {wt.shape[1]}")
            bs = _arr(sd[f"feature_net.{name}.0.bias"]).reshape(128)
            keep += [wt, bs]
            w.branch_w[i], w.branch_b[i] = wt.ctypes.data, bs.ctypes.data
        for field, key, shape in (("actor_fc_w", "fc.0.weight", (128, 1280)), ("actor_fc_b", "fc.0.bias", (128,)),
                                  ("actor_out_w", "out.weight", (3, 128)), ("actor_out_b", "out.bias", (3,))):
            a = _arr(sd[key]).reshape(shape)
            keep.append(a)
            setattr(w, field, a.ctypes.data)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.mansy_policy_create(C.byref(w), device, C.byref(h)))
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.mansy_policy_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def forward(self, obs: torch.Tensor, tensor_cores: bool = True, split: int = 0) -> torch.Tensor:
        """obs ``[N, stride >= 784]`` float32 rows -> predicted (normalised) QoE weights ``[N, 3]`` (a view of the
        kernel's ``[N, 16]`` output)."""
        if obs.device != self.device or obs.dtype != torch.float32 or obs.stride(-1) != 1:
            raise ValueError("obs must be a float32 row-major tensor on the identifier's device")
        n = obs.shape[0]
        out = torch.empty((n, 16), dtype=torch.float32, device=self.device)
        if tensor_cores:
            check(self.lib.mansy_policy_tc_set_split(self._h, int(split)))
            check(self.lib.mansy_policy_forward_tc(self._h, obs.data_ptr(), obs.stride(0), n, out.data_ptr(), None, None, None,
                                                   0, 0, 0, None, None, self._stream()))
        else:
            check(self.lib.mansy_policy_forward(self._h, obs.data_ptr(), obs.stride(0), n, out.data_ptr(), None, self._stream()))
        return out[:, :3]

    def reward(self, obs: torch.Tensor, qoe_reward: torch.Tensor, lamb: float, tensor_cores: bool = True):
        """``(identifier_reward float32 [N], (1 - lamb) * qoe_reward + lamb * identifier_reward float64 [N])``:
        ``calculate_indentifier_reward`` (utils/mansy_utils.py:42-49) and the blend of ``PPOPolicy.update``
        (models/mansy_ppo.py:40-49) for a whole batch of transitions in two launches."""
        pred = self.forward(obs, tensor_cores)
        n = obs.shape[0]
        ident = torch.empty(n, dtype=torch.float32, device=self.device)
        mixed = torch.empty(n, dtype=torch.float64, device=self.device)
        q = qoe_reward.to(device=self.device, dtype=torch.float32).contiguous()
        check(self.lib.mansy_identifier_reward(pred.data_ptr(), obs.data_ptr(), obs.stride(0), q.data_ptr(), float(lamb), n,
                                               ident.data_ptr(), mixed.data_ptr(), self._stream()))
        return ident, mixed


def gae_returns(reward: torch.Tensor, value: torch.Tensor, done: torch.Tensor, last_value: torch.Tensor, gamma: float,
                gae_lambda: float):
    """Generalised advantage estimation on the device over ``[T, N]`` rollout rings (``mansy_gae``; what tianshou's
    ``compute_episodic_return`` does on the host).  Returns ``(advantage, returns)`` float32 ``[T, N]``."""
    lib = _capi.load_library()
    T, n = reward.shape
    reward, value = reward.contiguous(), value.contiguous()
    done = done.to(torch.uint8).contiguous()
    last_value = last_value.to(torch.float32).contiguous()
    adv = torch.empty((T, n), dtype=torch.float32, device=reward.device)
    ret = torch.empty((T, n), dtype=torch.float32, device=reward.device)
    check(lib.mansy_gae(reward.data_ptr(), value.data_ptr(), done.data_ptr(), last_value.data_ptr(), T, n, float(gamma),
                        float(gae_lambda), adv.data_ptr(), ret.data_ptr(), torch.cuda.current_stream(reward.device).cuda_stream))
    return adv, ret


def seeded_state_dict(shapes, seed: int) -> Dict[str, np.ndarray]:
    """Deterministic float32 weights from numpy (uniform +-1/sqrt(fan_in)); the same call gives the
    same weights in the build container and on the GPU box."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in shapes:
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else int(shape[0])
        bound = 1.0 / np.sqrt(max(fan_in, 1))
        out[name] = rng.uniform(-bound, bound, size=shape).astype(np.float32)
    return out


def mansy_state_dict_shapes(hidden: int = 128):
    """(name, shape) lists of Actor and Critic-head state dicts (models/mansy.py:14-23,59-60,73-74)."""
    feat = [(f"feature_net.{n}.0.weight", (hidden, 5, 64) if k == 320 else ((hidden, 1, k) if n.startswith("conv") else (hidden, k)))
            for n, k in zip(MANSY_BRANCHES, MANSY_BRANCH_K)]
    feat_all = []
    for (n, s) in feat:
        feat_all += [(n, s), (n.replace("weight", "bias"), (hidden,))]
    actor = feat_all + [("fc.0.weight", (hidden, 10 * hidden)), ("fc.0.bias", (hidden,)),
                        ("out.weight", (15, hidden)), ("out.bias", (15,))]
    critic_head = [("fc.0.weight", (hidden, 10 * hidden)), ("fc.0.bias", (hidden,)),
                   ("out.weight", (1, hidden)), ("out.bias", (1,))]
    return actor, critic_head


def simple_state_dict_shapes():
    """models/simple_rl.py:14-19,42-44,56-58."""
    shp = {"conv1d_1": (128, 1, 8), "conv1d_2": (128, 1, 320), "fc1": (128, 1), "fc2": (128, 2), "fc3": (128, 64)}
    feat_all = []
    for n in SIMPLE_BRANCHES:
        feat_all += [(f"feature_net.{n}.0.weight", shp[n]), (f"feature_net.{n}.0.bias", (128,))]
    actor = feat_all + [("fc.0.weight", (128, 640)), ("fc.0.bias", (128,)), ("out.weight", (15, 128)), ("out.bias", (15,))]
    critic_head = [("fc.0.weight", (128, 640)), ("fc.0.bias", (128,)), ("out.weight", (1, 128)), ("out.bias", (1,))]
    return actor, critic_head
