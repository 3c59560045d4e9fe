"""ctypes binding of libcmr_b200.so (include/cmr_b200.h).

There is no CPU fallback: if the shared library is missing, or a call is made
without a CUDA device, this module raises.  PyTorch supplies device memory and
the current stream only.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CMR_B200_LIB", os.path.join(_HERE, "libcmr_b200.so"))

_c_int = ctypes.c_int
_c_vp = ctypes.c_void_p
_c_f = ctypes.c_float
_c_sz = ctypes.c_size_t
_c_i64 = ctypes.c_int64

# name -> (restype, argtypes); mirrors include/cmr_b200.h one to one
SIGNATURES = {
    "cmr_abi_version": (_c_int, []),
    "cmr_error_string": (ctypes.c_char_p, [_c_int]),
    "cmr_launch_count": (ctypes.c_ulonglong, []),
    "cmr_take_fault": (_c_int, [_c_vp]),
    "cmr_workspace_bytes": (_c_sz, [_c_int] * 4),
    "cmr_cloud_mean": (_c_int, [_c_vp, _c_int, _c_int, _c_vp, _c_vp]),
    "cmr_episode_prepare": (_c_int, [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    "cmr_observe": (_c_int, [_c_vp] * 7 + [_c_int] * 5 + [_c_vp] * 5),
    "cmr_cost_volume_workspace_bytes": (ctypes.c_size_t, [_c_int] * 5),
    "cmr_cost_volume_prepare": (_c_int, [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    "cmr_cost_volume_warp": (_c_int, [_c_vp] * 5 + [_c_int] * 7 + [_c_vp, _c_vp]),
    "cmr_sample_workspace_bytes": (ctypes.c_size_t, [_c_int] * 3),
    "cmr_sample_prepare": (_c_int, [_c_vp, _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    "cmr_sample_image_features": (_c_int, [_c_vp] * 5 + [_c_int] * 5 + [_c_vp, _c_vp, _c_vp]),
    "cmr_fps_f64": (_c_int, [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_vp, _c_vp, _c_vp]),
    "cmr_nearest_f64": (_c_int, [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    "cmr_episode_scan": (_c_int, [_c_vp, _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    "cmr_episode_compact": (_c_int, [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    "cmr_project": (_c_int, [_c_vp] * 6 + [_c_int] * 5 + [_c_vp] * 6 + [_c_int, _c_vp]),
    "cmr_tile_scatter": (_c_int, [_c_vp] * 3 + [_c_int] * 6 + [_c_vp] * 2),
    "cmr_to_disentangled": (_c_int, [_c_vp, _c_vp, _c_int, _c_vp]),
    "cmr_step": (_c_int, [_c_vp] * 5 + [_c_int] * 3 + [_c_vp]),
    "cmr_expert": (_c_int, [_c_vp] * 4 + [_c_int] * 3 + [_c_vp] * 3),
    "cmr_reward_scratch_bytes": (_c_sz, [_c_int]),
    "cmr_reward": (_c_int, [_c_vp] * 6 + [_c_int] * 3 + [_c_vp] * 4),
    "cmr_square_distance": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    "cmr_index_points": (_c_int, [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    "cmr_index_points_backward": (_c_int, [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    "cmr_farthest_point_sample": (_c_int, [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    "cmr_farthest_point_sample_grid": (_c_int, [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_vp, _c_vp, _c_vp]),
    "cmr_knn": (_c_int, [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    "cmr_knn_grid_workspace_bytes": (_c_sz, [_c_int, _c_int]),
    "cmr_knn_grid": (_c_int, [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp, _c_vp]),
    "cmr_query_ball_point": (_c_int, [_c_vp, _c_vp, _c_f, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    "cmr_query_ball_point_grid": (_c_int, [_c_vp, _c_vp, _c_f, _c_f, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp, _c_vp]),
    "cmr_group_points": (_c_int, [_c_vp] * 4 + [_c_int] * 5 + [_c_vp, _c_vp]),
    "cmr_reward_compare": (_c_int, [_c_vp, _c_vp, _c_int, _c_vp, _c_vp, _c_vp]),
    "cmr_iteration": (_c_int, [_c_vp] * 10),
    "cmr_session_create": (_c_int, [_c_vp, _c_vp]),
    "cmr_session_destroy": (None, [_c_vp]),
    "cmr_session_submit": (_c_int, [_c_vp, _c_vp, _c_vp]),
    "cmr_session_wait": (_c_int, [_c_vp, ctypes.c_longlong, _c_vp, _c_vp, _c_vp, _c_vp]),
    "cmr_session_stats": (_c_int, [_c_vp, _c_vp, _c_vp]),
    "cmr_session_last_observation": (_c_int, [_c_vp, ctypes.c_longlong, _c_vp, _c_vp, _c_vp]),
    "cmr_tower_blob_bytes": (_c_sz, [_c_int]),
    "cmr_tower_pack": (_c_int, [_c_int] + [_c_vp] * 8),
    "cmr_tower_workspace_bytes": (_c_sz, [_c_int, _c_int]),
    "cmr_tower_forward": (_c_int, [_c_vp] * 6 + [_c_int, _c_int, _c_vp, _c_vp]),
    "cmr_conv_epilogue": (_c_int, [_c_vp, _c_vp, _c_vp, _c_f, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    "cmr_to_channels_last": (_c_int, [_c_vp, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    "cmr_deterministic_action": (_c_int, [_c_vp, _c_int, _c_i64, _c_vp, _c_int, _c_i64, _c_int, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp]),
    "cmr_grouped_linear": (_c_int, [_c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_f, _c_int, _c_vp, _c_int,
                                    _c_vp]),
}

_lib = None


class CmrError(RuntimeError):
    pass


def load():
    """Load libcmr_b200.so (once).  Raises if it has not been built - never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CmrError(
                f"{LIB_PATH} is missing: build it with `python -m cmr_agent_b200.build` "
                "(or __graft_entry__.build()); cmr_agent_b200 has no CPU or PyTorch fallback")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.cmr_abi_version() != 5:
            raise CmrError("libcmr_b200.so ABI version mismatch; rebuild it")
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().cmr_error_string(rc).decode()
        raise CmrError(f"{what} failed: {msg} (code {rc})")


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


try:                                     # the raw handle without building a torch.cuda.Stream object (~1 us saved per call)
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:                   # pragma: no cover
    _raw_stream = None


def stream_handle(device_index):
    """cudaStream_t of torch's current stream on `device_index`, as an int."""
    if _raw_stream is not None:
        return _raw_stream(device_index)
    return torch.cuda.current_stream(device_index).cuda_stream


def bind(name):
    """The ctypes function itself (hot paths bind once and check the return code themselves)."""
    return getattr(load(), name)


def fail(rc, what):
    msg = load().cmr_error_string(rc).decode()
    raise CmrError(f"{what} failed: {msg} (code {rc})")


class SessionConfig(ctypes.Structure):
    """cmr_session_config of include/cmr_b200.h."""
    _fields_ = [(n, ctypes.c_int) for n in ("B", "N", "C", "H", "W", "iters", "dof6", "reward_mode", "depth",
                                            "features_resident", "nbins")] + \
               [("rot_tab", ctypes.c_void_p), ("t_tab", ctypes.c_void_p)]


class RolloutInputs(ctypes.Structure):
    """cmr_rollout_inputs of include/cmr_b200.h."""
    _fields_ = [(n, ctypes.c_void_p) for n in ("pc", "overlap", "feat", "img_feat", "K", "P", "pc_in_cam", "pc_mask",
                                               "action_r", "action_t")]


class IterationArgs(ctypes.Structure):
    """cmr_iteration_args of include/cmr_b200.h."""
    _fields_ = [(n, ctypes.c_void_p) for n in ("pc", "overlap", "img_feat", "K", "mean", "workspace", "rot_tab", "t_tab",
                                               "target", "mask", "reward_scratch", "dist_cached")] + \
               [(n, ctypes.c_int) for n in ("B", "N", "C", "H", "W", "nbins", "dof6", "reward_mode")]


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def require_cuda(t, name, dtype=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise CmrError(f"{name} must be a CUDA tensor: cmr_agent_b200 has no CPU fallback")
    if dtype is not None and t.dtype != dtype:
        raise CmrError(f"{name} must be {dtype}, got {t.dtype}")
    return t


def call(name, *args):
    check(getattr(load(), name)(*args), name)


def launch_count():
    return int(load().cmr_launch_count())


def take_fault():
    """Read-and-clear the sticky device fault flag (synchronises the current stream)."""
    return int(load().cmr_take_fault(stream()))
