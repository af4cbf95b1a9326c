"""ctypes binding of the C ABI declared in ``include/mansy_b200.h``.

The shared library is built in-tree (``mansy_immersivevideostreaming_b200/csrc/
libmansy_b200.so``) by ``__graft_entry__.build()`` / ``python -m
mansy_immersivevideostreaming_b200.build``.  There is no fallback: if the library is missing
or a CUDA call fails, the product raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# MANSY_LIB_PATH: another build of the same library (A/B timing of two builds on one box); never a different implementation
LIB_PATH = os.environ.get("MANSY_LIB_PATH") or os.path.join(_HERE, "csrc", "libmansy_b200.so")

AUX_DOUBLES = 16
STATS_DOUBLES = 16
TOTALS_DOUBLES = 6          # mansy_episode_totals / mansy_peer_allgather_stats columns
PEER_HANDLE_BYTES = 64


# csrc/mansy_sim.cuh EnvState (one 128-byte record per environment)
ENV_STATE_FIELDS = [
    ("cur_time", "<f8"), ("buf", "<f8"), ("prev_vq", "<f8"), ("next_chunk", "<i4"), ("cur_idx", "<i4"),
    ("ep_step", "<i4"), ("cursor", "<i4"), ("video", "<i4"), ("pair", "<i4"),
    ("trace", "<i4"), ("end_chunk", "<i4"), ("sample_id", "<i4"), ("flags", "<i4"),
    ("w0", "<f4"), ("w1", "<f4"), ("w2", "<f4"), ("start_chunk", "<i4"),
    ("sum_qoe", "<f8"), ("sum_q1", "<f8"), ("sum_q2", "<f8"), ("sum_q3", "<f8"),
    ("ep_return", "<f8"), ("reserved", "<f8"),
]


class MansyError(RuntimeError):
    pass


class Tables(C.Structure):
    _fields_ = [
        ("size", C.c_void_p), ("quality", C.c_void_p), ("video_time", C.c_void_p),
        ("vp_gt", C.c_void_p), ("vp_pred", C.c_void_p), ("vp_acc", C.c_void_p),
        ("vp_start", C.c_void_p), ("vp_end", C.c_void_p),
        ("trace", C.c_void_p), ("trace_len", C.c_void_p),
        ("qoe_w", C.c_void_p), ("samples", C.c_void_p),
        ("n_videos", C.c_int32), ("n_chunks", C.c_int32), ("n_users", C.c_int32), ("n_vp_chunks", C.c_int32),
        ("n_traces", C.c_int32), ("trace_stride", C.c_int32), ("n_qoe", C.c_int32), ("n_samples", C.c_int32),
    ]


class Cfg(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int32), ("env_offset", C.c_int32), ("worker_num", C.c_int32), ("seed", C.c_int32),
        ("obs_mode", C.c_int32), ("reward_mode", C.c_int32), ("video_rates", C.c_int32 * 5),
        ("startup_download", C.c_int32), ("chunk_length", C.c_int32), ("max_size", C.c_int32),
        ("max_throughput", C.c_int32),
    ]


class Out(C.Structure):
    _fields_ = [
        ("obs", C.c_void_p), ("obs_stride", C.c_int64), ("reward", C.c_void_p), ("done", C.c_void_p),
        ("aux", C.c_void_p), ("tile_versions", C.c_void_p),
    ]


class Rollout(C.Structure):
    _fields_ = [
        ("obs", C.c_void_p), ("obs_stride", C.c_int64), ("slabs", C.c_int32),
        ("actions", C.c_void_p), ("logp", C.c_void_p), ("value", C.c_void_p), ("reward", C.c_void_p),
        ("done", C.c_void_p), ("logits", C.c_void_p),
    ]


class RolloutHost(C.Structure):
    _fields_ = [
        ("host_slabs", C.c_int32), ("obs", C.c_void_p), ("actions", C.c_void_p), ("logp", C.c_void_p),
        ("value", C.c_void_p), ("reward", C.c_void_p), ("done", C.c_void_p),
    ]


ROLLOUT_FP32_POLICY = 1
ROLLOUT_TIME_KERNELS = 2
ROLLOUT_NO_PDL = 4
ROLLOUT_TWO_KERNELS = 8
ROLLOUT_NO_ZERO_COPY = 16


class PolicyWeights(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("branch_w", C.c_void_p * 10), ("branch_b", C.c_void_p * 10),
        ("actor_fc_w", C.c_void_p), ("actor_fc_b", C.c_void_p), ("actor_out_w", C.c_void_p),
        ("actor_out_b", C.c_void_p), ("critic_fc_w", C.c_void_p), ("critic_fc_b", C.c_void_p),
        ("critic_out_w", C.c_void_p), ("critic_out_b", C.c_void_p),
    ]


MTIO_MAX_LAYERS = 4
MTIO_FP32 = 1
MTIO_TIME_KERNELS = 2


class MtioAttn(C.Structure):
    _fields_ = [("in_proj_w", C.c_void_p), ("in_proj_b", C.c_void_p), ("out_w", C.c_void_p), ("out_b", C.c_void_p)]


class MtioLayer(C.Structure):
    _fields_ = [
        ("self_attn", MtioAttn), ("cross_attn", MtioAttn),
        ("lin1_w", C.c_void_p), ("lin1_b", C.c_void_p), ("lin2_w", C.c_void_p), ("lin2_b", C.c_void_p),
        ("norm1_w", C.c_void_p), ("norm1_b", C.c_void_p), ("norm2_w", C.c_void_p), ("norm2_b", C.c_void_p),
        ("norm3_w", C.c_void_p), ("norm3_b", C.c_void_p),
    ]


class MtioWeights(C.Structure):
    _fields_ = [
        ("n_enc", C.c_int32), ("n_dec", C.c_int32), ("his_window", C.c_int32), ("fut_window", C.c_int32),
        ("pe_rows", C.c_int32), ("reserved", C.c_int32),
        ("emb_w", C.c_void_p), ("emb_b", C.c_void_p), ("pe", C.c_void_p),
        ("enc", MtioLayer * MTIO_MAX_LAYERS), ("dec", MtioLayer * MTIO_MAX_LAYERS),
        ("enc_norm_w", C.c_void_p), ("enc_norm_b", C.c_void_p), ("dec_norm_w", C.c_void_p), ("dec_norm_b", C.c_void_p),
        ("conv_w", C.c_void_p), ("conv_b", C.c_void_p),
        ("bn_w", C.c_void_p), ("bn_b", C.c_void_p), ("bn_mean", C.c_void_p), ("bn_var", C.c_void_p),
        ("pred_w", C.c_void_p), ("pred_b", C.c_void_p),
    ]


# name -> (restype, argtypes); every symbol declared in include/mansy_b200.h
_vp = C.c_void_p
SIGNATURES = {
    "mansy_last_error": (C.c_char_p, []),
    "mansy_abi_version": (C.c_int, []),
    "mansy_kernel_launches": (C.c_int64, []),
    "mansy_create": (C.c_int, [C.POINTER(Tables), C.POINTER(Cfg), C.c_int, C.POINTER(_vp)]),
    "mansy_destroy": (C.c_int, [_vp]),
    "mansy_seed": (C.c_int, [_vp, C.c_int32, _vp]),
    "mansy_reset": (C.c_int, [_vp, _vp, C.c_int32, _vp, C.c_int64, _vp]),
    "mansy_step": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.POINTER(Out), _vp]),
    "mansy_step_host": (C.c_int, [_vp, _vp, C.c_int32, _vp, _vp, _vp, _vp]),
    "mansy_reset_host": (C.c_int, [_vp, _vp, _vp]),
    "mansy_rollout_random": (C.c_int, [_vp, C.c_int32, C.c_uint64, C.c_int64, C.c_int32, C.POINTER(Out), _vp]),
    "mansy_episode_stats": (C.c_int, [_vp, _vp, _vp]),
    "mansy_stats_clear": (C.c_int, [_vp, _vp]),
    "mansy_episode_totals": (C.c_int, [_vp, _vp, _vp]),
    "mansy_set_outcome_table": (C.c_int, [_vp, C.c_int32]),
    "mansy_peer_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_int, C.POINTER(_vp)]),
    "mansy_peer_export": (C.c_int, [_vp, _vp]),
    "mansy_peer_connect": (C.c_int, [_vp, _vp]),
    "mansy_peer_barrier": (C.c_int, [_vp, _vp]),
    "mansy_peer_allgather_stats": (C.c_int, [_vp, _vp, _vp, C.POINTER(_vp)]),
    "mansy_peer_timed_out": (C.c_int, [_vp, C.POINTER(C.c_int32)]),
    "mansy_peer_destroy": (C.c_int, [_vp]),
    "mansy_state_snapshot": (C.c_int, [_vp, _vp, _vp]),
    "mansy_error_flag": (C.c_int, [_vp, C.POINTER(C.c_int32)]),
    "mansy_expert_actions": (C.c_int, [_vp, C.c_int32, _vp, _vp, _vp]),
    "mansy_viewport_tiles": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       _vp, _vp, _vp, _vp]),
    "mansy_allocate_tile_versions": (C.c_int, [_vp, _vp, C.c_int64, C.POINTER(C.c_int32 * 5), _vp, _vp]),
    "mansy_policy_create": (C.c_int, [C.POINTER(PolicyWeights), C.c_int, C.POINTER(_vp)]),
    "mansy_policy_destroy": (C.c_int, [_vp]),
    "mansy_policy_forward": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, _vp, _vp, _vp]),
    "mansy_policy_sample": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_uint64, C.c_int64, C.c_int32, _vp, _vp, _vp]),
    "mansy_policy_forward_tc": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, _vp, _vp, _vp, _vp, C.c_uint64, C.c_int64,
                                          C.c_int32, _vp, _vp, _vp]),
    "mansy_policy_forward_tc_sim": (C.c_int, [_vp, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp, C.c_uint64, C.c_int64, _vp]),
    "mansy_policy_forward_tc_timeline": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, _vp, _vp, _vp, _vp, C.c_uint64, C.c_int64,
                                                   C.c_int32, _vp, _vp, _vp, _vp]),
    "mansy_policy_tc_set_split": (C.c_int, [_vp, C.c_int32]),
    "mansy_debug_progress": (C.c_int, [_vp]),
    "mansy_identifier_reward": (C.c_int, [_vp, _vp, C.c_int64, _vp, C.c_double, C.c_int32, _vp, _vp, _vp]),
    "mansy_gae": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int32, C.c_int32, C.c_double, C.c_double, _vp, _vp, _vp]),
    "mansy_debug_fused_timeline": (C.c_int, [_vp, C.c_int32]),
    "mansy_rollout_policy": (C.c_int, [_vp, _vp, C.POINTER(Rollout), C.c_int32, C.c_int64, C.c_uint64, C.c_int32, _vp]),
    "mansy_rollout_policy_host": (C.c_int, [_vp, _vp, C.POINTER(Rollout), C.POINTER(RolloutHost), C.c_int32, C.c_int64,
                                            C.c_uint64, C.c_int32, _vp]),
    "mansy_rollout_reserve_timing": (C.c_int, [_vp, C.c_int32]),
    "mansy_rollout_kernel_ms": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "mansy_selftest_allocate": (C.c_int, [C.c_uint64, C.c_int32, C.POINTER(C.c_int32 * 5), C.POINTER(C.c_uint8 * 64)]),
    "mansy_selftest_fov_mask": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                          C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]),
    "mansy_selftest_centre_to_pixel": (C.c_int, [C.c_float, C.c_int32]),
    "mansy_selftest_download": (C.c_int, [_vp, C.c_int32, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_double),
                                          C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mansy_selftest_ddiv_rcp": (C.c_int, [_vp, _vp, C.c_int64, _vp]),
    "mansy_selftest_hashed_action": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64]),
    "mansy_mtio_create": (C.c_int, [C.POINTER(MtioWeights), C.c_int, C.c_int32, C.POINTER(_vp)]),
    "mansy_mtio_destroy": (C.c_int, [_vp]),
    "mansy_mtio_sample": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, _vp]),
    "mansy_mtio_sample_host": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp, _vp]),
    "mansy_linreg_sample": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, C.c_int32, _vp, _vp]),
    "mansy_mtio_kernel_ms": (C.c_int, [_vp, C.POINTER(C.c_double * 3), C.POINTER(C.c_int32 * 3)]),
}

_lib: Optional[C.CDLL] = None


def load_library(path: Optional[str] = None) -> C.CDLL:
    """Load ``libmansy_b200.so`` and attach signatures; raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise MansyError(
            f"{p} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()'). "
            "There is no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.mansy_abi_version() != 1:
        raise MansyError("libmansy_b200.so ABI version mismatch")
    if path is None:
        _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load_library().mansy_last_error()
        raise MansyError(f"libmansy_b200 error {rc}: {msg.decode() if msg else '?'}")
