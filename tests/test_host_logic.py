"""Host-side logic on CPU: table packing, sample lists, synthetic generator, observation layouts,
the reference-format writer/packer round trip, and the multi-rank (gloo, world_size 2) plumbing."""
import os
import tempfile

import numpy as np
import pytest
import torch

from helpers import golden_tables, load_golden
from mansy_immersivevideostreaming_b200 import synth
from mansy_immersivevideostreaming_b200.config import (MANSY_OBS_FLOATS, MANSY_OBS_SEGMENTS, MANSY_OBS_STRIDE,
                                                       SIMPLE_OBS_FLOATS, SIMPLE_OBS_SEGMENTS, SIMPLE_OBS_STRIDE,
                                                       SimConfig, rate_out_lut)
from mansy_immersivevideostreaming_b200.rollout import all_gather_stats, broadcast_state_dict, shard_range, summarise_stats
from mansy_immersivevideostreaming_b200.tables import (SimTables, environment_samples, environment_test_samples,
                                                       masks_to_u64, u64_to_masks)
from mansy_immersivevideostreaming_b200.vector_env import episode_log_line
from oracle import sim_oracle as so

CFG = SimConfig()


def _tables(**kw):
    return synth.make_synthetic_tables(lambda g, p: so.chunk_masks(g, p, CFG), n_videos=2, n_users=3, n_traces=4,
                                       seed=5, trace_len_range=(20, 40), **kw)


@pytest.mark.parametrize("segs,floats,stride", [(MANSY_OBS_SEGMENTS, MANSY_OBS_FLOATS, MANSY_OBS_STRIDE),
                                                (SIMPLE_OBS_SEGMENTS, SIMPLE_OBS_FLOATS, SIMPLE_OBS_STRIDE)])
def test_obs_layouts(segs, floats, stride):
    cover = np.zeros(stride, dtype=int)
    for key, off, shape in segs:
        n = int(np.prod(shape))
        cover[off:off + n] += 1
        if n >= 4:
            assert off % 4 == 0, f"{key} must start on a 16-byte boundary"
    assert cover.max() == 1 and cover.sum() == floats          # no overlap, reference payload size
    assert stride % 8 == 0                                      # whole 32-byte sectors


def test_sample_lists_follow_reference_rules():
    # defaults of the reference: V=18, U=45, T=24, Q=4 -> 72 samples, sample i = (i%V, i%U, i%T, i%Q)
    s = environment_samples(18, 45, 24, 4)
    assert s.shape == (72, 4)
    assert np.array_equal(s[50], [50 % 18, 50 % 45, 50 % 24, 50 % 4])
    t = environment_test_samples(3, 15, 8, 4)
    assert t.shape == (1440, 4) and np.array_equal(t[0], [0, 0, 0, 0]) and np.array_equal(t[5], [0, 0, 1, 1])
    assert np.array_equal(t[-1], [2, 14, 7, 3])
    # shipped results.csv order: first test episode is video 21, user 3, trace 31 (index 0 of each split list)


def test_mask_bit_packing_roundtrip():
    rng = np.random.default_rng(0)
    m = (rng.random((7, 5, 64)) < 0.3).astype(np.uint8)
    bits = masks_to_u64(m)
    assert bits.dtype == np.uint64 and np.array_equal(u64_to_masks(bits), m)
    assert int(masks_to_u64(np.eye(64, dtype=np.uint8)[63])) == 1 << 63


def test_synthetic_tables_statistics_and_determinism():
    a, b = _tables(), _tables()
    for k in SimTables._ARRAYS:
        assert np.array_equal(getattr(a, k), getattr(b, k))
    assert a.size.min() >= 3800 and a.size.max() <= 990000
    assert np.all(a.quality[:, :, 3, :] == 16)
    t, lens = synth.synth_traces(np.random.default_rng(1), 40)
    assert lens.min() >= 166 and lens.max() <= 758
    vals = np.concatenate([t[i, :lens[i]] for i in range(40)])
    assert 0.005 < np.mean(vals == 0) < 0.03 and 2.5e6 < vals.mean() < 5.5e6
    pc = np.array([bin(int(x)).count("1") for x in a.vp_gt.reshape(-1) if x])
    assert pc.min() >= 4 and pc.max() <= 40


def test_tables_validation_rejects_bad_inputs():
    t = _tables()
    kw = {k: getattr(t, k).copy() for k in SimTables._ARRAYS}
    bad = dict(kw); bad["trace"] = kw["trace"] * 0
    with pytest.raises(ValueError):
        SimTables(cfg=CFG, n_users=t.n_users, **bad)                      # the reference would never return
    bad = dict(kw); bad["samples"] = kw["samples"].copy(); bad["samples"][0, 0] = 99
    with pytest.raises(ValueError):
        SimTables(cfg=CFG, n_users=t.n_users, **bad)
    bad = dict(kw); bad["vp_start"] = kw["vp_start"] + 10
    with pytest.raises(ValueError):
        SimTables(cfg=CFG, n_users=t.n_users, **bad)                      # simulator.py:44 assert
    with pytest.raises(ValueError):
        SimConfig(tile_num_width=6).validate()


def test_reference_layout_roundtrip():
    """write_reference_layout -> pack_from_reference_layout gives the same tables (no reference needed:
    the generated config.yml is read back with a plain attribute dict)."""
    import yaml
    from mansy_immersivevideostreaming_b200.tables import pack_from_reference_layout

    class Attr(dict):
        __getattr__ = dict.__getitem__

    t = _tables(short_tail_frac=0.4)
    root = tempfile.mkdtemp()
    cfg_path = synth.write_reference_layout(t, root)
    doc = Attr(yaml.safe_load(open(cfg_path)))
    for key in ("viewport_datasets_dir", "video_datasets_dir", "network_datasets_dir"):   # utils/common.py:21-25
        doc[key] = {k: doc["datasets_base_dir"] + v for k, v in doc[key].items()}
    packed = pack_from_reference_layout(doc, "Synth", "SynthNet", list(t.video_ids), list(t.user_ids), list(t.trace_ids),
                                        [[float(x) for x in w] for w in t.qoe_w], mode="train")
    for k in SimTables._ARRAYS:
        assert np.array_equal(getattr(packed, k), getattr(t, k)), k


def test_rate_lut_and_log_line():
    assert rate_out_lut((1, 5, 8, 16, 35))[4] == (4, 4, 3, 2, 2)
    g = load_golden("mansy_real.npz")
    tables = golden_tables(g)
    first = str(g["test_log"]).strip().splitlines()[1]
    env = so.OracleEnv(tables, 1, 0, "f64", worker_id=1, worker_num=7)
    for a in g["test_action"]:
        env.reset() if a < 0 else env.step(int(a))
        if env.episodes:
            break
    e = env.episodes[0]
    line = episode_log_line(tables, e["sample_id"], *e["sums"], e["steps"]).strip()
    assert line.split(",")[:6] == first.split(",")[:6]
    for x, y in zip(line.split(",")[6:], first.split(",")[6:]):
        assert abs(float(x) - float(y)) <= 2e-5


def test_shard_range_and_summary():
    assert shard_range(65536, 3, 8) == (3 * 8192, 8192)
    with pytest.raises(ValueError):
        shard_range(10, 0, 3)
    stats = torch.tensor([[2.0, 1.0, 0.5, 0.25, 50.0, 1.0], [4.0, 2.0, 1.5, 0.75, 50.0, 1.0]], dtype=torch.float64)
    s = summarise_stats(stats)
    assert s["episodes"] == 2 and s["steps"] == 100 and abs(s["mean_qoe"] - 0.06) < 1e-12


def _gloo_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, per = shard_range(16, rank, world)
    local = torch.arange(start * 6, (start + per) * 6, dtype=torch.float64).reshape(per, 6)   # rank-ordered rows
    full = all_gather_stats(local)
    torch.save(full, os.path.join(out_dir, f"r{rank}.pt"))
    # weight broadcast: every rank ends up with rank 0's policy state dict
    from mansy_immersivevideostreaming_b200.policy import mansy_state_dict_shapes, seeded_state_dict
    sd = seeded_state_dict(mansy_state_dict_shapes()[0], 100 + rank)          # ranks start with DIFFERENT weights
    got = broadcast_state_dict(sd, src=0)
    want = seeded_state_dict(mansy_state_dict_shapes()[0], 100)
    ok = list(got) == list(want) and all(np.array_equal(got[k], want[k]) and got[k].shape == want[k].shape for k in want)
    torch.save(torch.tensor(ok), os.path.join(out_dir, f"b{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_stats_world_size_2_gloo():
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_dir = tempfile.mkdtemp()
    mp.spawn(_gloo_worker, args=(2, port, out_dir), nprocs=2, join=True)
    expect = torch.arange(16 * 6, dtype=torch.float64).reshape(16, 6)
    for r in range(2):
        assert torch.equal(torch.load(os.path.join(out_dir, f"r{r}.pt")), expect)   # sharded == unsharded order
        assert bool(torch.load(os.path.join(out_dir, f"b{r}.pt")))                  # broadcast weights == rank 0's


def test_numa_helpers_degrade_to_no_ops():
    """numa.py is best-effort host plumbing: list parsing, and no policy change where there is nothing to choose from."""
    from mansy_immersivevideostreaming_b200 import numa
    assert numa._parse_list("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11} and numa._parse_list("") == set()
    nodes = numa.online_nodes()
    assert nodes and nodes[0] == 0
    with numa.memory_on_node(None) as applied:
        assert applied is False
    with numa.memory_on_node(max(nodes) + 7) as applied:          # a node that does not exist
        assert applied is False
    buf = np.ones(1 << 16, dtype=np.float32)
    where = numa.pages_node(buf.ctypes.data, buf.nbytes)
    assert where == {} or set(where) <= set(nodes) | {-2, -14}     # (negative = errno of an unmapped / foreign page)
