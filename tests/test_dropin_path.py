"""The PYTHONPATH-first shim (``dropin_path/envs``): the reference's scripts import ``envs.mansy_env`` /
``envs.simple_rl_env`` / ``envs.expert_env`` by top-level name (run_mansy.py:16, run_simple_rl.py:16, run_expert.py), so
putting ``dropin_path`` on ``PYTHONPATH`` binds the CUDA-backed classes without editing them.  The GPU tests execute the
reference's own test loops (run_mansy.py:161-175, run_simple_rl.py:130-147) VERBATIM -- the loop source is read out of the
unmodified reference files (oracle/_ref archive) -- against the shim and compare the episode log row for row with the
oracle and with the unmodified reference env driven by the same loop."""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "dropin_path")

# the reference loops, quoted for boxes without the oracle/_ref archive (the archive's text wins when present)
_MANSY_LOOP = """\
    with torch.no_grad():
        state = test_env.reset()
        sample_count = test_env.sample_count()
        for i in tqdm(range(sample_count), desc='Testing: '):
            done = False
            while not done:
                for key, value in state.items():
                    state[key] = np.expand_dims(value, 0)
                batch = Batch(obs=state, info={})
                results = policy(batch, state=None)
                logits, act, dist = results.logits, results.act, results.dist
                action = act.item()
                state, reward, done, _ = test_env.step(action)
            state = test_env.reset()
        read_log_file(test_log_path)
"""
_SIMPLE_LOOP = """\
    with torch.no_grad():
        state = test_env.reset()
        sample_count = test_env.sample_count()
        for i in tqdm(range(sample_count), desc='Testing: '):
            video, user, trace = test_env.current_video, test_env.current_user, test_env.current_trace
            done = False
            while not done:
                for key, value in state.items():
                    state[key] = np.expand_dims(value, 0)
                # batch = {'obs': state}
                batch = Batch(obs=state, info={})
                # logits = policy(batch, state=None).logits
                # action = F.softmax(logits, dim=-1).argmax().item()
                action = policy(batch, state=None).act
                action = action.item()
                state, reward, done, _ = test_env.step(action)
            state = test_env.reset()
        read_log_file(test_log_path)
"""


def _loop_source(script: str, first: int, last: int, quoted: str) -> str:
    """Lines first..last (1-based) of the reference script, dedented; the archive's bytes when available."""
    from oracle import ref_loader
    if ref_loader.code_available():
        if ref_loader.reference_available():
            text = open(os.path.join(ref_loader.REFERENCE_ROOT, "bitrate_selection", script)).read()
        else:
            text = ref_loader.read_member(f"bitrate_selection/{script}").decode()
        src = "\n".join(text.splitlines()[first - 1:last]) + "\n"
        assert [l.strip() for l in src.strip().splitlines()] == [l.strip() for l in quoted.strip().splitlines()], \
            "the quoted loop drifted from the reference file"
        return textwrap.dedent(src)
    return textwrap.dedent(quoted)


def _run(code: str, cwd: str, pythonpath: str):
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    env["PYTHONPATH"] = pythonpath
    return subprocess.run([sys.executable, "-c", code], cwd=cwd, env=env, capture_output=True, text=True, timeout=300)


def test_shim_imports_with_only_pythonpath_set(built_library):
    """`python -c "import envs.mansy_env"` with nothing but PYTHONPATH=<repo>/dropin_path."""
    cwd = tempfile.mkdtemp()
    p = _run("import envs.mansy_env, envs.simple_rl_env, envs.expert_env\n"
             "from envs.mansy_env import MANSYEnv\n"
             "from envs.simple_rl_env import SimpleRLEnv\n"
             "from envs.expert_env import ExpertEnv\n"
             "import inspect\n"
             "print(MANSYEnv.__module__, list(inspect.signature(MANSYEnv.__init__).parameters)[1:5])", cwd, SHIM)
    assert p.returncode == 0, p.stderr
    assert p.stdout.split()[0] == "mansy_immersivevideostreaming_b200.dropin.envs.mansy_env"
    assert "'config', 'dataset', 'network_dataset', 'qoe_weights'" in p.stdout


def test_shim_wins_over_the_script_directory(built_library):
    """The reference's ``envs/`` has no ``__init__.py`` (namespace package) and the script directory is ``sys.path[0]``;
    the shim is a regular package, which the import system prefers -- while ``utils`` / ``models`` keep resolving to the
    reference's own directories."""
    bs = os.path.join(tempfile.mkdtemp(), "bitrate_selection")
    for pkg, mod, body in (("envs", "mansy_env", "MANSYEnv = 'reference'\n"), ("utils", "common", "WHO = 'reference utils'\n"),
                           ("models", "mansy", "WHO = 'reference models'\n")):
        os.makedirs(os.path.join(bs, pkg))
        open(os.path.join(bs, pkg, mod + ".py"), "w").write(body)
    open(os.path.join(bs, "run_probe.py"), "w").write(
        "from envs.mansy_env import MANSYEnv\nfrom utils.common import WHO as u\nfrom models.mansy import WHO as m\n"
        "print(MANSYEnv if isinstance(MANSYEnv, str) else MANSYEnv.__module__, '|', u, '|', m)\n")
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    env["PYTHONPATH"] = SHIM
    p = subprocess.run([sys.executable, "run_probe.py"], cwd=bs, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert p.stdout.strip() == "mansy_immersivevideostreaming_b200.dropin.envs.mansy_env | reference utils | reference models"


# ---------------------------------------------------------------------------------------------------------------------
class _Batch:                      # tianshou.data.Batch as the loops use it: keyword construction, attribute access
    def __init__(self, **kw):
        self.__dict__.update(kw)


class _CounterPolicy:
    """Stand-in for ``policy(batch, state=None)``: a deterministic action stream (so two runs of the loop are
    comparable) that also checks what the loop hands it -- every observation array with a leading batch axis of 1."""

    def __init__(self, keys_shapes):
        self.t = 0
        self.keys_shapes = keys_shapes

    def __call__(self, batch, state=None):
        import torch
        for k, shape in self.keys_shapes.items():
            a = batch.obs[k]
            assert a.shape == (1,) + tuple(shape) and a.dtype == np.float32, (k, a.shape)
        act = torch.tensor([(7 * self.t + 3 * (self.t // 5)) % 15])
        self.t += 1
        return _Batch(logits=torch.zeros(1, 15), act=act, dist=None)


def _drive(loop_src, test_env, log_path, keys_shapes):
    import torch
    seen = []
    ns = {"torch": torch, "np": np, "tqdm": lambda it, desc=None: it, "Batch": _Batch, "policy": _CounterPolicy(keys_shapes),
          "test_env": test_env, "test_log_path": log_path, "read_log_file": lambda path: seen.append(path)}
    exec(compile(loop_src, "<reference test loop>", "exec"), ns)
    assert seen == [log_path]
    return open(log_path).read().strip().splitlines()


def _dataset(root):
    from mansy_immersivevideostreaming_b200 import synth
    from mansy_immersivevideostreaming_b200.config import SimConfig
    from oracle import sim_oracle as so
    cfg = SimConfig()
    t = synth.make_synthetic_tables(lambda g, p: so.chunk_masks(g, p, cfg), n_videos=2, n_users=2, n_chunks=24,
                                    n_traces=2, seed=21, trace_len_range=(30, 60), short_tail_frac=0.5)
    return t, synth.write_reference_layout(t, root)


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["mansy", "simple_rl"])
def test_reference_test_loop_runs_verbatim_on_the_shim(which):
    from mansy_immersivevideostreaming_b200.config import (MANSY_OBS_SEGMENTS, OBS_MODE_MANSY, OBS_MODE_SIMPLE, REWARD_QOE,
                                                           SIMPLE_OBS_SEGMENTS)
    from mansy_immersivevideostreaming_b200.dropin.envs._common import tables_for
    from mansy_immersivevideostreaming_b200.refconfig import load_config_yml
    from mansy_immersivevideostreaming_b200.vector_env import episode_log_line
    from oracle import ref_loader
    from oracle import sim_oracle as so
    root = tempfile.mkdtemp()
    _, cfg_path = _dataset(root)
    config = load_config_yml(cfg_path)
    w = config.qoe_split["test"]
    # the module the UNMODIFIED script would import, resolved through the shim directory exactly as PYTHONPATH does
    sys.path.insert(0, SHIM)
    try:
        for k in [k for k in sys.modules if k == "envs" or k.startswith("envs.")]:
            del sys.modules[k]
        if which == "mansy":
            from envs.mansy_env import MANSYEnv as Env
            loop = _loop_source("run_mansy.py", 161, 175, _MANSY_LOOP)
            segs, obs_mode = MANSY_OBS_SEGMENTS, OBS_MODE_MANSY
            make = lambda cls, cfg, log: cls(cfg, "Synth", "SynthNet", w, None, 0.5, log, cfg.startup_download, mode="test", seed=5, device="cpu")   # noqa: E731  run_mansy.py:148-150
        else:
            from envs.simple_rl_env import SimpleRLEnv as Env
            loop = _loop_source("run_simple_rl.py", 130, 147, _SIMPLE_LOOP)
            segs, obs_mode = SIMPLE_OBS_SEGMENTS, OBS_MODE_SIMPLE
            make = lambda cls, cfg, log: cls(cfg, "Synth", "SynthNet", w, log, cfg.startup_download, mode="test", seed=5, device="cpu")   # noqa: E731  run_simple_rl.py:117-119
    finally:
        sys.path.remove(SHIM)
    assert Env.__module__.startswith("mansy_immersivevideostreaming_b200.dropin.envs")
    keys_shapes = {k: shape for k, _, shape in segs}
    log = os.path.join(root, f"{which}_results.csv")
    env = make(Env, config, log)
    env.seed(5)                                                               # run_mansy.py:150
    rows = _drive(loop, env, log, keys_shapes)
    env.close()
    tables = tables_for(config, "Synth", "SynthNet", w, "test", config.startup_download)
    n = tables.n_samples
    assert rows[0] == "video,user,trace,qoe_w1,qoe_w2,qoe_w3,qoe,qoe1,qoe2,qoe3" and len(rows) == 1 + n

    # (1) the oracle (f64 chain, the kernels' contract) driven by the same action stream: identical text, row for row
    orc = so.OracleEnv(tables, obs_mode, REWARD_QOE, "f64", worker_id=0, worker_num=1)
    pol = _CounterPolicy({})
    orc.reset()
    want = []
    for _ in range(n):
        done = False
        while not done:
            _, _, done, _ = orc.step(int(pol(_Batch(obs={})).act.item()))
        e = orc.episodes[-1]
        want.append(episode_log_line(tables, e["sample_id"], *e["sums"], e["steps"]).strip())
        orc.reset()
    assert rows[1:] == want

    # (2) the UNMODIFIED reference env under the same verbatim loop (float32 QoE chain under numpy 2): same episodes in
    # the same order, ids and weights identical, the rounded means within 2e-5 (1e-5 relative + the 5-decimal rounding)
    if ref_loader.code_available():
        ref = ref_loader.load_reference()
        rconfig = ref.common.get_config_from_yml(cfg_path)
        rlog = os.path.join(root, f"{which}_reference.csv")
        RefEnv = ref.mansy_env.MANSYEnv if which == "mansy" else ref.simple_rl_env.SimpleRLEnv
        with ref_loader.silence_prints():
            renv = make(RefEnv, rconfig, rlog)
        renv.seed(5)
        rrows = _drive(loop, renv, rlog, keys_shapes)
        assert len(rrows) == len(rows) and rrows[0] == rows[0]
        for a, b in zip(rows[1:], rrows[1:]):
            fa, fb = a.split(","), b.split(",")
            assert fa[:3] == fb[:3] and [float(x) for x in fa[3:6]] == [float(x) for x in fb[3:6]], (a, b)
            np.testing.assert_allclose([float(x) for x in fa[6:]], [float(x) for x in fb[6:]], rtol=1e-5, atol=2e-5)
