"""ctypes binding of libmeshgen_b200.so (C ABI declared in include/meshgen_b200.h).

The library is the product: if it is missing or cannot be loaded this module raises -- there is
no Python/CPU fallback for the environment dynamics.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.environ.get("MESHGEN_LIB") or os.path.join(_PKG, "lib", "libmeshgen_b200.so")  # MESHGEN_LIB: tuning variants
CSRC = os.path.join(_PKG, "csrc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]

SYMBOLS = [
    "mg_create", "mg_set_domains", "mg_set_random", "mg_set_auto_reset", "mg_reset", "mg_step", "mg_move", "mg_step_host", "mg_step_host_begin", "mg_step_host_end",
    "mg_set_obs_delta", "mg_set_host_delta", "mg_last_host_bytes", "mg_sample_actions", "mg_sample_actions_seq", "mg_get_state", "mg_get_elements", "mg_debug_polygon",
    "mg_stats", "mg_stats_async", "mg_set_log_capacity", "mg_log_capacity", "mg_replay_add", "mg_snapshot_bytes",
    "mg_snapshot_save", "mg_snapshot_load", "mg_set_option", "mg_set_kernel_timing", "mg_kernel_times", "mg_num_envs", "mg_max_verts",
    "mg_launch_count", "mg_destroy", "mg_last_error", "mg_version",
]


class PolygenCfg(C.Structure):
    _fields_ = [("ctr_x", C.c_double), ("ctr_y", C.c_double), ("ave_radius", C.c_double), ("irregularity", C.c_double),
                ("spikeyness", C.c_double), ("min_coarse", C.c_int32), ("max_coarse", C.c_int32),
                ("min_verts", C.c_int32), ("max_verts", C.c_int32)]


class EpisodeStats(C.Structure):
    _fields_ = [("episodes", C.c_int64), ("completed", C.c_int64), ("truncated", C.c_int64), ("steps", C.c_int64),
                ("successes", C.c_int64), ("elements", C.c_int64), ("sum_n", C.c_int64), ("sum_n_success", C.c_int64),
                ("ring_items", C.c_int64), ("sum_n_ring", C.c_int64),
                ("sum_return", C.c_double), ("sum_length", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class StateView(C.Structure):
    _fields_ = [("n", C.c_int32), ("ref_index", C.c_int32), ("n_elements", C.c_int32), ("failed_num", C.c_int32),
                ("n0", C.c_int32), ("memo_flags", C.c_int32), ("base_length", C.c_double), ("current_area", C.c_double),
                ("original_area", C.c_double), ("area_min", C.c_double), ("area_crit", C.c_double),
                ("xy_host", C.c_void_p), ("vertex_id_host", C.c_void_p), ("cand_key_host", C.c_void_p),
                ("cand_stamp_host", C.c_void_p)]


def build(force: bool = False, extra_flags=()) -> str:
    """Compile csrc/mg_abi.cu (which includes the kernels) into lib/libmeshgen_b200.so for sm_100a."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(_ROOT, "include", "meshgen_b200.h")]
    newest = max(os.path.getmtime(s) for s in srcs)
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= newest:
        return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    cmd = ["nvcc", *NVCC_FLAGS, *extra_flags, "-o", LIB_PATH, os.path.join(CSRC, "mg_abi.cu")]
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def load():
    """Load the shared library and declare the prototypes; raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback for the environment.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
    L.mg_create.argtypes = [C.POINTER(vp), i32, i32, i32]
    L.mg_set_domains.argtypes = [vp, vp, vp, i32, vp, vp]
    L.mg_set_random.argtypes = [vp, u64, C.POINTER(PolygenCfg), i64]
    L.mg_set_auto_reset.argtypes = [vp, i32]
    L.mg_reset.argtypes = [vp, vp, vp, vp]
    L.mg_step.argtypes = [vp] + [vp] * 7 + [vp]
    L.mg_move.argtypes = [vp] + [vp] * 7 + [vp]
    L.mg_step_host.argtypes = [vp] + [vp] * 7
    L.mg_step_host_begin.argtypes = [vp] + [vp] * 7
    L.mg_step_host_end.argtypes = [vp]
    L.mg_set_obs_delta.argtypes = [vp, i32]
    L.mg_set_host_delta.argtypes = [vp, i32]
    L.mg_last_host_bytes.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.mg_sample_actions.argtypes = [vp, u64, u64, vp, vp]
    L.mg_sample_actions_seq.argtypes = [vp, u64, vp, vp, vp]
    L.mg_get_state.argtypes = [vp, i32, C.POINTER(StateView)]
    L.mg_get_elements.argtypes = [vp, i32, vp, i32, C.POINTER(C.c_int32), vp, i32, C.POINTER(C.c_int32)]
    L.mg_debug_polygon.argtypes = [vp, i32, i32, vp, i32, C.POINTER(C.c_int32), C.POINTER(C.c_double), vp, C.POINTER(C.c_int32),
                                   C.POINTER(C.c_double)]
    L.mg_stats.argtypes = [vp, C.POINTER(EpisodeStats), i32]
    L.mg_stats_async.argtypes = [vp, vp, i32, vp]
    L.mg_set_log_capacity.argtypes = [vp, i32, i32]
    L.mg_log_capacity.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.mg_replay_add.argtypes = [vp, i64, i64] + [vp] * 14
    L.mg_snapshot_bytes.argtypes = [vp]
    L.mg_snapshot_bytes.restype = i64
    L.mg_snapshot_save.argtypes = [vp, vp, vp]
    L.mg_snapshot_load.argtypes = [vp, vp, i64, vp]
    L.mg_set_option.argtypes = [vp, C.c_char_p, i32]
    L.mg_set_kernel_timing.argtypes = [vp, i32]
    L.mg_kernel_times.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.mg_num_envs.argtypes = [vp]
    L.mg_max_verts.argtypes = [vp]
    L.mg_launch_count.argtypes = [vp]
    L.mg_launch_count.restype = i64
    L.mg_destroy.argtypes = [vp]
    L.mg_last_error.argtypes = [vp]
    L.mg_last_error.restype = C.c_char_p
    L.mg_version.restype = C.c_char_p
    for s in SYMBOLS:
        if getattr(L, s).restype is C.c_int or s in ("mg_create",):
            getattr(L, s).restype = C.c_int
    _lib = L
    return L


class MeshgenError(RuntimeError):
    pass


def check(rc: int, handle=None, what: str = ""):
    if rc != 0:
        msg = load().mg_last_error(handle)
        raise MeshgenError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")
