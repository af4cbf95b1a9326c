"""The C-ABI library on a box WITHOUT a GPU: it must build, load, export every symbol the header
declares, refuse to compute without a device, and its shared scalar building blocks (the same
__host__ __device__ code the kernels run) must agree with the oracle / goldens."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import load_golden
from mansy_immersivevideostreaming_b200 import _capi
from mansy_immersivevideostreaming_b200.config import SimConfig, rate_out_lut
from mansy_immersivevideostreaming_b200.synth import synthetic_actions
from oracle import sim_oracle as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = SimConfig()
RATES = (C.c_int32 * 5)(*CFG.video_rates)


@pytest.fixture(scope="module")
def lib(built_library):
    return _capi.load_library()


def test_header_symbols_exported(lib):
    header = open(os.path.join(ROOT, "include", "mansy_b200.h")).read()
    declared = set(re.findall(r"\b(mansy_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/mansy_b200.h but not exported"
    assert declared == set(_capi.SIGNATURES), "ctypes signature table out of sync with the header"


def test_library_is_sm100a_only(built_library):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", built_library], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mansy_immersivevideostreaming_b200.simulator import BatchSimulator
    g = load_golden("mansy_synth.npz")
    from helpers import golden_tables
    with pytest.raises(_capi.MansyError):
        BatchSimulator(golden_tables(g), 4)
    # and the raw ABI refuses too
    t, c, h = _capi.Tables(), _capi.Cfg(), C.c_void_p()
    assert lib.mansy_create(C.byref(t), C.byref(c), 0, C.byref(h)) != 0
    assert lib.mansy_last_error()


def test_selftest_allocate_matches_golden(lib):
    g = load_golden("allocate_kat.npz")
    out = (C.c_uint8 * 64)()
    for m, vers in zip(g["mask"], g["versions"]):
        for a in range(16):
            assert lib.mansy_selftest_allocate(int(m), a, C.byref(RATES), C.byref(out)) == 0
            assert np.array_equal(np.frombuffer(out, dtype=np.uint8), vers[a]), (hex(int(m)), a)
    # out-of-table actions behave like (0, 0) (utils/common.py:103)
    assert lib.mansy_selftest_allocate(int(g["mask"][5]), -3, C.byref(RATES), C.byref(out)) == 0
    assert np.array_equal(np.frombuffer(out, dtype=np.uint8), so.allocate_tile_versions(0, 0, int(g["mask"][5]), CFG.video_rates))


def test_selftest_allocate_random_vs_oracle(lib):
    rng = np.random.default_rng(3)
    out = (C.c_uint8 * 64)()
    for rates in ((1, 5, 8, 16, 35), (2, 3, 10, 40, 41), (1, 2, 4, 8, 16)):
        cr = (C.c_int32 * 5)(*rates)
        for _ in range(60):
            m = int(rng.integers(0, 2**63)) & int(rng.integers(0, 2**63))
            a = int(rng.integers(0, 15))
            lib.mansy_selftest_allocate(m, a, C.byref(cr), C.byref(out))
            assert np.array_equal(np.frombuffer(out, dtype=np.uint8),
                                  so.allocate_tile_versions(*so.action_to_rates(a), m, rates))
    assert rate_out_lut(CFG.video_rates) == ((0, 0, 0, 0, 0), (1, 1, 0, 0, 0), (2, 2, 1, 0, 0), (3, 3, 2, 1, 1), (4, 4, 3, 2, 2))


def test_selftest_fov_mask_matches_golden(lib):
    g = load_golden("geometry_kat.npz")
    mask, valid = C.c_uint64(), C.c_int32()
    for x, y, m in zip(g["x"], g["y"], g["mask"]):
        lib.mansy_selftest_fov_mask(int(x), int(y), 2560, 1440, 600, 300, C.byref(mask), C.byref(valid))
        assert valid.value == 1 and mask.value == int(m), (x, y)
    lib.mansy_selftest_fov_mask(-1, 5, 2560, 1440, 600, 300, C.byref(mask), C.byref(valid))
    assert valid.value == 0


def test_selftest_fov_exhaustive_axes_vs_oracle(lib):
    """Every x on the horizontal axis and every y on the vertical axis (the mask is an outer product)."""
    mask, valid = C.c_uint64(), C.c_int32()
    for x in range(0, 2561):
        lib.mansy_selftest_fov_mask(x, 700, 2560, 1440, 600, 300, C.byref(mask), C.byref(valid))
        assert mask.value == so.mask_bits(so.fov_tile_mask(x, 700, CFG))
    for y in range(0, 1441):
        lib.mansy_selftest_fov_mask(1000, y, 2560, 1440, 600, 300, C.byref(mask), C.byref(valid))
        assert mask.value == so.mask_bits(so.fov_tile_mask(1000, y, CFG))


def test_selftest_centre_to_pixel(lib):
    rng = np.random.default_rng(1)
    vals = np.concatenate([rng.random(2000).astype(np.float32),
                           np.array([0.0, 1.0, 0.5, 0.49999997, 0.99999994, 0.125], dtype=np.float32)])
    for v in vals:
        assert lib.mansy_selftest_centre_to_pixel(float(v), 2560) == so.centre_to_pixels(v, v, CFG)[0]
        assert lib.mansy_selftest_centre_to_pixel(float(v), 1440) == so.centre_to_pixels(v, v, CFG)[1]


def test_selftest_download_matches_golden(lib):
    g = load_golden("trace_kat.npz")
    for thr, L, sizes, rec in zip(g["thr"], g["lens"], g["sizes"], g["rec"]):
        thr = np.ascontiguousarray(thr)
        idx, tm, buf, dl, rb = C.c_int32(0), C.c_double(0.0), C.c_double(3.0), C.c_double(), C.c_double()
        for s, (rdl, ridx, rtm, rrb, rbuf) in zip(sizes, rec):
            assert lib.mansy_selftest_download(thr.ctypes.data, int(L), int(s), C.byref(idx), C.byref(tm), C.byref(buf),
                                               C.byref(dl), C.byref(rb)) == 0
            # bit-exact float64: same operation order as CPython, no FMA contraction
            assert (dl.value, idx.value, tm.value, rb.value, buf.value) == (rdl, int(ridx), rtm, rrb, rbuf)


def test_hashed_action_stream_matches_numpy(lib):
    for seed in (0, 1234, 2**40 + 17):
        for step in (0, 1, 77, 10**6):
            acts = synthetic_actions(64, step, seed=seed, env_offset=1000)
            for k in (0, 5, 63):
                assert lib.mansy_selftest_hashed_action(seed, 1000 + k, step) == int(acts[k])
    assert set(np.unique(synthetic_actions(4096, 3))) == set(range(15))


def test_new_entry_points_validate_arguments_without_a_device(lib):
    """MTIO / expert / regression entry points reject bad arguments with MANSY_E_INVALID (-1) before touching CUDA, and
    mansy_mtio_create refuses to run without a device (no CPU fallback)."""
    import torch
    w, h = _capi.MtioWeights(), C.c_void_p()
    assert lib.mansy_mtio_create(None, 0, 16, C.byref(h)) == -1
    w.n_enc, w.n_dec, w.his_window, w.fut_window, w.pe_rows = 0, 2, 5, 15, 16
    assert lib.mansy_mtio_create(C.byref(w), 0, 16, C.byref(h)) == -1 and b"n_enc" in lib.mansy_last_error()
    w.n_enc, w.fut_window = 2, 40
    assert lib.mansy_mtio_create(C.byref(w), 0, 16, C.byref(h)) == -1 and b"fut_window" in lib.mansy_last_error()
    w.fut_window = 15
    assert lib.mansy_mtio_create(C.byref(w), 0, 16, C.byref(h)) == -1 and b"NULL" in lib.mansy_last_error()   # no weights
    assert lib.mansy_mtio_sample(None, None, None, 1, 0, 0, None, None, None) == -1
    assert lib.mansy_mtio_sample_host(None, None, None, 1, 0, 0, None, None) == -1
    assert lib.mansy_expert_actions(None, 4, None, None, None) == -1
    assert lib.mansy_linreg_sample(None, None, 1, 5, 15, None, None) == -1
    assert lib.mansy_mtio_destroy(None) == 0
    if not torch.cuda.is_available():
        keep = [np.zeros(8, np.float32)]
        for name, _ in _capi.MtioWeights._fields_:
            if name in ("emb_w", "emb_b", "pe", "enc_norm_w", "dec_norm_w", "conv_w", "conv_b", "bn_w", "bn_b", "bn_mean",
                        "bn_var", "pred_w", "pred_b"):
                setattr(w, name, keep[0].ctypes.data)
        assert lib.mansy_mtio_create(C.byref(w), 0, 16, C.byref(h)) == -2       # MANSY_E_CUDA: no device


def test_reciprocal_division_is_an_ieee_division(lib):
    """csrc/mansy_core.cuh ddiv_rcp (what the kernels use for x / throughput, x / tile count, x / 35, x / 5e6, x / 5): bit-identical
    to a / b on the divisors the simulator meets -- random numerators, and numerators built to land next to a rounding
    midpoint of the quotient (the hard cases of a division)."""
    rng = np.random.default_rng(20260102)
    n = 2_000_000
    cases = []
    # (1) throughput entries: integers up to 1.4e7 B/s and rescaled (fractional) traces; numerators: remaining bytes
    b = np.concatenate([np.floor(rng.uniform(1, 1.4e7, n)), rng.uniform(1e-3, 1.4e7, n)])
    a = np.concatenate([np.floor(rng.uniform(0, 1e7, n)), rng.uniform(0, 1e7, n)])
    cases.append((a, b))
    # (2) the constant divisors with wide numerators
    consts = np.array([35.0, 5e6, 5.0] + list(range(1, 65)), dtype=np.float64)
    b = rng.choice(consts, 2 * n)
    a = np.ldexp(rng.uniform(0.5, 1.0, 2 * n), rng.integers(-30, 30, 2 * n))
    cases.append((a, b))
    # (3) near-midpoint quotients: a = b * (q + ulp(q) / 2) nudged by -1 / 0 / +1 ulp
    b = np.concatenate([rng.choice(consts, n), np.floor(rng.uniform(1, 1.4e7, n))])
    q = np.ldexp(rng.uniform(1.0, 2.0, 2 * n), rng.integers(-15, 15, 2 * n))
    a = q * b + (np.nextafter(q, np.inf) - q) * 0.5 * b
    nudge = rng.integers(0, 3, 2 * n)
    a = np.where(nudge == 1, np.nextafter(a, 0.0), np.where(nudge == 2, np.nextafter(a, np.inf), a))
    cases.append((a, b))
    for a, b in cases:
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        out = np.empty_like(a)
        assert lib.mansy_selftest_ddiv_rcp(a.ctypes.data, b.ctypes.data, a.size, out.ctypes.data) == 0
        assert np.array_equal(out, a / b)
