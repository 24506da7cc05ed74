"""ctypes binding of include/b2rl.h.  Fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb2rl.so")
MAX_FIELDS = 8

c_i64, c_i32, c_u64, c_u32 = C.c_int64, C.c_int32, C.c_uint64, C.c_uint32
c_f32, c_f64, c_vp = C.c_float, C.c_double, C.c_void_p


class ReplayDesc(C.Structure):
    _fields_ = [("capacity", c_i64), ("n_fields", c_i32), ("device", c_i32),
                ("field_bytes", c_i64 * MAX_FIELDS)]


# name -> (restype, argtypes); must list every symbol include/b2rl.h declares.
SIGNATURES = {
    "b2rl_last_error": (C.c_char_p, []),
    "b2rl_version": (C.c_int, []),
    "b2rl_replay_create": (C.c_int, [C.POINTER(ReplayDesc), C.POINTER(c_vp)]),
    "b2rl_replay_destroy": (C.c_int, [c_vp]),
    "b2rl_replay_size": (C.c_int, [c_vp, C.POINTER(c_i64), C.POINTER(c_i64), C.POINTER(c_i64)]),
    "b2rl_replay_field_ptr": (C.c_int, [c_vp, c_i32, C.POINTER(c_vp)]),
    "b2rl_replay_push": (C.c_int, [c_vp, C.POINTER(c_vp), c_vp, c_i64, c_vp]),
    "b2rl_replay_reserve": (C.c_int, [c_vp, c_i64, C.POINTER(c_i64), c_vp]),
    "b2rl_replay_copy_payload": (C.c_int, [c_vp, C.POINTER(c_vp), c_i64, c_i64, c_vp]),
    "b2rl_replay_commit": (C.c_int, [c_vp, c_vp, c_i64, c_vp]),
    "b2rl_replay_ingest_pipelined": (C.c_int, [c_vp, C.POINTER(c_vp), c_vp, c_i64, c_vp]),
    "b2rl_replay_evict": (C.c_int, [c_vp, c_i64, c_vp]),
    "b2rl_replay_fill_hash": (C.c_int, [c_vp, c_i64, c_u32, c_vp]),
    "b2rl_tree_build": (C.c_int, [c_vp, c_vp, c_i64, c_vp]),
    "b2rl_tree_sample": (C.c_int, [c_vp, c_vp, c_u64, c_u64, c_i64, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "b2rl_replay_seed": (C.c_int, [c_vp, c_u64, c_u64, c_vp]),
    "b2rl_tree_sample_stream": (C.c_int, [c_vp, c_i64, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "b2rl_tree_sample_fetch": (C.c_int, [c_vp, c_i64, c_f32, c_vp, c_vp, c_vp, c_vp, C.POINTER(c_vp), c_vp]),
    "b2rl_philox_uniforms": (C.c_int, [c_u64, c_u64, c_i64, c_vp, c_vp]),
    "b2rl_tree_update": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    "b2rl_tree_stats": (C.c_int, [c_vp, c_f32, c_vp, c_vp, c_vp]),
    "b2rl_tree_leaves": (C.c_int, [c_vp, c_i64, c_i64, c_vp, c_vp]),
    "b2rl_replay_gather": (C.c_int, [c_vp, c_vp, c_i64, C.POINTER(c_vp), c_vp]),
    "b2rl_apex_target": (C.c_int, [c_vp] * 7 + [c_i32, c_i32, c_f32, c_f32] + [c_vp] * 6),
    "b2rl_r2d2_target": (C.c_int, [c_vp] * 6 + [c_i32, c_i32, c_i32, c_i32, c_f64, c_f32, c_i32]
                         + [c_vp] * 6),
    "b2rl_vtrace": (C.c_int, [c_vp] * 5 + [c_i32, c_i32, c_f32, c_f32, c_f32, c_f32] + [c_vp] * 3),
    "b2rl_conv1_pack": (C.c_int, [c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "b2rl_conv1_pack_jobs": (C.c_int, [C.POINTER(c_vp), C.POINTER(c_i32), C.POINTER(c_i32), C.POINTER(c_vp),
                                       C.POINTER(c_vp), c_i32, c_i32, c_vp]),
    "b2rl_conv1_fused": (C.c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i32, c_i32, c_vp, c_i32, c_vp]),
    "b2rl_conv1_wgrad_workspace_floats": (c_i64, [c_i32]),
    "b2rl_conv1_wgrad": (C.c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i32, c_vp, c_vp, c_i32, c_vp]),
    "b2rl_rmsprop_step": (C.c_int, [C.POINTER(c_vp), C.POINTER(c_vp), C.POINTER(c_vp), C.POINTER(c_vp),
                                    C.POINTER(c_i64), c_i32, c_f64, c_f64, c_f64, c_i32, c_vp, c_vp, c_vp]),
    "b2rl_rmsprop_norm_finish": (C.c_int, [c_vp, c_i32, c_vp, c_vp]),
    "b2rl_peer_allreduce_max_ctas": (c_i32, []),
    "b2rl_peer_allreduce_mean": (C.c_int, [c_vp, c_vp, c_i32, c_i32, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "b2rl_peer_allreduce_mean_big": (C.c_int, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp]),
    "b2rl_gemm_packed_floats": (c_i64, [c_i64, c_i64, c_i32]),
    "b2rl_gemm_split_pack": (C.c_int, [c_vp, c_i64, c_i64, c_i64, c_i32, c_i32, c_vp, c_vp]),
    "b2rl_gemm_split_pack_into": (C.c_int, [c_vp, c_i64, c_i64, c_i64, c_i32, c_i32, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp]),
    "b2rl_gemm_pack_act_nhwc": (C.c_int, [c_vp, c_i64, c_i64, c_i64, c_i32, c_i32, c_vp, c_vp]),
    "b2rl_unflatten_relu_mask": (C.c_int, [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "b2rl_gemm_workspace_floats": (c_i64, [c_i64, c_i64, c_i64, c_i64]),
    "b2rl_gemm_tf32x3": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "b2rl_dueling_forward": (C.c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "b2rl_dueling_backward": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "b2rl_dueling_backward_w": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "b2rl_launch_count": (c_i64, []),
}

_lib = None


class B2RLError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libb2rl.so.  No fallback: a missing library is a hard error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise B2RLError(
            f"{LIB_PATH} is missing — build it with `python -m distributed_rl_b200.build` "
            "(nvcc, sm_100a).  There is deliberately no CPU fallback for the product path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise B2RLError(f"libb2rl.so does not export {name}")
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise B2RLError(f"libb2rl error {rc}: {load().b2rl_last_error().decode()}")
