"""Shared test helpers: golden loading, tolerances, replay drivers."""
from __future__ import annotations

import os

import numpy as np

from mansy_immersivevideostreaming_b200.config import (MANSY_OBS_SEGMENTS, OBS_MODE_MANSY, OBS_MODE_SIMPLE,
                                                       SIMPLE_OBS_SEGMENTS)
from mansy_immersivevideostreaming_b200.tables import SimTables

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Observation segments whose value depends on the QoE numeric chain (float32 under numpy >= 2,
# float64 under the reference's pinned numpy 1.24).  Everything else in an
# observation row is chain-independent and must match bit for bit.
CHAIN_DEPENDENT = ("past_viewport_qualities", "past_quality_variances")

# north_star tolerance: 1e-5 relative on download times, buffer levels and QoE rewards.
RTOL = 1e-5


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def golden_tables(g) -> SimTables:
    return SimTables.from_npz_dict(g)


def segments(obs_mode):
    return MANSY_OBS_SEGMENTS if obs_mode == OBS_MODE_MANSY else SIMPLE_OBS_SEGMENTS


def assert_rows_match(row, ref_row, obs_mode, chain_exact: bool, ctx=""):
    """Compare one packed observation row with a reference row.

    chain_exact=True  -> every segment bit-exact (same numeric chain on both sides).
    chain_exact=False -> chain-dependent segments within 1e-5 relative (+1e-6 absolute, the
                         float32 resolution of the reference's own accumulation), rest bit-exact.
    """
    for key, off, shape in segments(obs_mode):
        n = int(np.prod(shape))
        a, b = row[off:off + n], ref_row[off:off + n]
        if chain_exact or key not in CHAIN_DEPENDENT:
            assert np.array_equal(a, b), f"{ctx} segment {key}: {a} != {b}"
        else:
            np.testing.assert_allclose(a, b, rtol=RTOL, atol=1e-6, err_msg=f"{ctx} segment {key}")


def reward_scale(w, q1, q2, q3, norm: bool) -> float:
    """Magnitude of the terms of qoe = w1*q1 - w2*q2 - w3*q3 (cancellation-aware scale)."""
    s = abs(float(w[0]) * q1) + abs(float(w[1]) * q2) + abs(float(w[2]) * q3)
    if norm:
        s /= float(w[0]) + float(w[1]) + float(w[2])
    return max(s, 1e-30)


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-300)
