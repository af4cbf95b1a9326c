"""tests/golden/policy_shipped_kat.npz from the reference's SHIPPED checkpoints -- build container only.

    python -m oracle.make_golden_shipped

The reference ships one trained example run: ``best_policy.pth`` (tianshou PPOPolicy state dict:
``actor.*`` / ``critic.*`` / ``_actor_critic.*`` / ``identifier.*``) and ``best_identifier.pth``.  This script loads them
into the UNMODIFIED reference ``Actor`` / ``Critic`` / ``QoEIdentifier`` (bitrate_selection/models/mansy.py, imported by
oracle/ref_loader.py) and records their fp32 CPU outputs on real-data observation rows (``mansy_real.npz``: rows the
reference's MANSYEnv produced on the shipped Jin2022 / 4G data).  The weights themselves are not committed: the GPU
tests read them from the ``oracle/_ref`` archive (oracle/make_ref.py ``fixtures/``), which travels with gpurun.
"""
from __future__ import annotations

import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mansy_immersivevideostreaming_b200.config import MANSY_OBS_SEGMENTS     # noqa: E402
from oracle.ref_loader import load_reference, read_member                      # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def main():
    import torch
    ref = load_reference()
    g = np.load(os.path.join(GOLDEN, "mansy_real.npz"))
    rows = g["test_obs"][::3][:96].copy()                # reset rows and stepped rows of 6 episodes
    obs = {k: rows[:, off:off + int(np.prod(shape))].reshape((-1,) + tuple(shape)).copy() for k, off, shape in MANSY_OBS_SEGMENTS}
    M = ref.models_mansy
    sd = torch.load(io.BytesIO(read_member("fixtures/best_policy.pth")), map_location="cpu", weights_only=False)
    isd = torch.load(io.BytesIO(read_member("fixtures/best_identifier.pth")), map_location="cpu", weights_only=False)
    fn = M.FeatureNet(8, 64, 5, 128, device="cpu")                       # run_mansy.py:207-209: ONE FeatureNet instance
    actor = M.Actor(fn, 1280, 128, 15, "cpu")
    critic = M.Critic(fn, 1280, 128, "cpu")
    actor.load_state_dict({k[len("actor."):]: v for k, v in sd.items() if k.startswith("actor.")})
    critic.load_state_dict({k[len("critic."):]: v for k, v in sd.items() if k.startswith("critic.")})
    # the checkpoint's critic.feature_net.* and actor.feature_net.* are the same tensors (shared module)
    for k, v in sd.items():
        if k.startswith("critic.feature_net."):
            assert torch.equal(v, sd["actor." + k[len("critic."):]]), k
    ifn = M.QoEIdentifierFeatureNet(8, 64, 5, 15, 128, device="cpu")
    ident = M.QoEIdentifier(ifn, 1280, 128, "cpu")
    ident.load_state_dict(isd)
    ident_in_policy = M.QoEIdentifier(M.QoEIdentifierFeatureNet(8, 64, 5, 15, 128, device="cpu"), 1280, 128, "cpu")
    ident_in_policy.load_state_dict({k[len("identifier."):]: v for k, v in sd.items() if k.startswith("identifier.")})
    with torch.no_grad():
        logits, _ = actor(obs)
        value = critic(obs)
        pred = ident(obs, obs["action_one_hot"])
        pred2 = ident_in_policy(obs, obs["action_one_hot"])
    np.savez_compressed(os.path.join(GOLDEN, "policy_shipped_kat.npz"), rows=rows, actor_logits=logits.numpy(),
                        critic_value=value.numpy(), ident_out=pred.numpy(), ident_in_policy_out=pred2.numpy())
    print(f"policy_shipped_kat: {rows.shape[0]} real-data rows; |logits| max {np.abs(logits.numpy()).max():.3f}, "
          f"value range [{value.min():.3f}, {value.max():.3f}]")


if __name__ == "__main__":
    main()
