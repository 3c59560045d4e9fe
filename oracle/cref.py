"""ctypes binding of oracle/cmr_oracle.c (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libcmr_oracle.so")
    src = os.path.join(_HERE, "cmr_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libcmr_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.cmr_oracle_p2p.restype = ctypes.c_double
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def project(pc, mean, RT, K, H, W, fused=None):
    """pc [3,N], mean [3], RT [4,4], K [3,3] -> (idx int32 [N], in_cam uint8 [N]).
    fused: whether the bmm these columns belong to is in torch's FMA-chain regime (>= 45 columns);
    defaults to the regime of an N-column product."""
    pc, mean, RT, K = _f32(pc), _f32(mean).reshape(3), _f32(RT).reshape(16), _f32(K).reshape(9)
    N = pc.shape[1]
    idx = np.empty(N, np.int32)
    inc = np.empty(N, np.uint8)
    fused = (N >= 45) if fused is None else bool(fused)
    lib().cmr_oracle_project(_p(pc), _p(mean), _p(RT), _p(K), N, H, W, int(fused), _p(idx), _p(inc))
    return idx, inc


def scatter_mean(feat, overlap, idx, P):
    """feat [C,N], overlap [N] bool/uint8, idx int32 [N] -> [C,P]."""
    feat = _f32(feat)
    overlap = np.ascontiguousarray(overlap, dtype=np.uint8)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    C, N = feat.shape
    out = np.empty((C, P), np.float32)
    lib().cmr_oracle_scatter_mean(_p(feat), _p(overlap), _p(idx), N, C, P, _p(out))
    return out


def p2p(target, pc, mask, mean, RT, intended):
    target, pc = _f32(target), _f32(pc)
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    mean, RT = _f32(mean).reshape(3), _f32(RT).reshape(16)
    return lib().cmr_oracle_p2p(_p(target), _p(pc), _p(mask), _p(mean), _p(RT), pc.shape[1], int(intended))


def fps(xyz, npoint, start):
    """xyz [N,3] -> int64 [npoint]."""
    xyz = _f32(xyz)
    out = np.empty(npoint, np.int64)
    lib().cmr_oracle_fps(_p(xyz), xyz.shape[0], npoint, ctypes.c_int64(int(start)), _p(out))
    return out


def knn(q, ref, k):
    q, ref = _f32(q), _f32(ref)
    out = np.empty((q.shape[0], k), np.int64)
    lib().cmr_oracle_knn(_p(q), _p(ref), q.shape[0], ref.shape[0], k, _p(out))
    return out


def ball(q, ref, radius, nsample):
    q, ref = _f32(q), _f32(ref)
    out = np.empty((q.shape[0], nsample), np.int64)
    r2 = np.float32(float(radius) ** 2)
    lib().cmr_oracle_ball(_p(q), _p(ref), q.shape[0], ref.shape[0], ctypes.c_float(float(r2)), nsample, _p(out))
    return out


def sqdist(q, ref):
    q, ref = _f32(q), _f32(ref)
    out = np.empty((q.shape[0], ref.shape[0]), np.float32)
    lib().cmr_oracle_sqdist(_p(q), _p(ref), q.shape[0], ref.shape[0], _p(out))
    return out


def apply_step(pose, Rnew, move_t):
    pose = _f32(pose).copy().reshape(16)
    Rnew, move_t = _f32(Rnew).reshape(9), _f32(move_t).reshape(3)
    lib().cmr_oracle_apply_step(_p(pose), _p(Rnew), _p(move_t))
    return pose.reshape(4, 4)


def compose_xyz(Rx, Ry, Rz):
    out = np.empty(9, np.float32)
    Rx, Ry, Rz = _f32(Rx).reshape(9), _f32(Ry).reshape(9), _f32(Rz).reshape(9)
    lib().cmr_oracle_compose_xyz(_p(Rx), _p(Ry), _p(Rz), _p(out))
    return out.reshape(3, 3)


def to_disentangled(pose, mean):
    pose = _f32(pose).copy().reshape(16)
    mean = _f32(mean).reshape(3)
    lib().cmr_oracle_to_disentangled(_p(pose), _p(mean))
    return pose.reshape(4, 4)
