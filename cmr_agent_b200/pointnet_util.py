"""Drop-in for the reference's ``models/pointnet_util.py`` function API on B200.

Same names, argument order, shapes and dtypes as
/root/reference/models/pointnet_util.py (``square_distance:19``,
``index_points:36``, ``farthest_point_sample:50``, ``query_ball_point:73``,
``sample_and_group:96``, ``sample_and_group_all:136``).  ``models/PointNN.py``
imports ``index_points, square_distance`` from here unchanged after
``cmr_agent_b200.install()``.  Every function runs a hand-written sm_100a kernel
from libcmr_b200.so on the current CUDA stream; CUDA tensors only.

Extensions (not in the reference): ``knn_point`` - the fused, never-materialised
top-k that ``square_distance(...).argsort()[:, :, :k]`` amounts to, in the stable
order (distance, index); ``sample_and_group(knn=True)`` uses it.
"""
import ctypes

import torch

from . import _lib


def _f32c(t, name):
    t = _lib.require_cuda(t, name, torch.float32)
    return t if t.is_contiguous() else t.contiguous()


def _i64c(t, name):
    t = _lib.require_cuda(t, name)
    if t.dtype != torch.int64:
        t = t.long()
    return t if t.is_contiguous() else t.contiguous()


def _square_distance_raw(src, dst):
    B, S, Cs = src.shape
    Bd, N, Cd = dst.shape
    if Cs != 3 or Cd != 3 or B != Bd:
        raise _lib.CmrError("square_distance expects src [B,S,3] and dst [B,N,3]")
    out = torch.empty(B, S, N, device=src.device, dtype=torch.float32)
    if out.numel() == 0:
        return out
    ss = (ctypes.c_int64 * 3)(*src.stride())
    ds = (ctypes.c_int64 * 3)(*dst.stride())
    _lib.call("cmr_square_distance", _lib.ptr(src), ctypes.cast(ss, ctypes.c_void_p), _lib.ptr(dst),
              ctypes.cast(ds, ctypes.c_void_p), B, S, N, _lib.ptr(out), _lib.stream())
    return out


class _SquareDistance(torch.autograd.Function):
    """d[s,n] = |src_s - dst_n|^2; d/dsrc_s = 2 sum_n g[s,n] (src_s - dst_n), d/ddst_n = -2 sum_s g[s,n] (src_s - dst_n).
    The reference differentiates the same expression through broadcasting (pointnet_util.py:33); its callers only
    ever pass coordinates, so this backward exists for gradient equivalence, not for speed."""

    @staticmethod
    def forward(ctx, src, dst):
        ctx.save_for_backward(src, dst)
        return _square_distance_raw(src, dst)

    @staticmethod
    def backward(ctx, g):
        src, dst = ctx.saved_tensors
        g = g.contiguous()
        gs = gd = None
        if ctx.needs_input_grad[0]:
            gs = 2.0 * (g.sum(dim=2, keepdim=True) * src - torch.bmm(g, dst))
        if ctx.needs_input_grad[1]:
            gd = 2.0 * (g.sum(dim=1).unsqueeze(-1) * dst - torch.bmm(g.transpose(1, 2), src))
        return gs, gd


def square_distance(src, dst):
    """pointnet_util.py:19-33: src [B,S,3], dst [B,N,3] (any strides) -> [B,S,N] f32."""
    _lib.require_cuda(src, "src", torch.float32)
    _lib.require_cuda(dst, "dst", torch.float32)
    if torch.is_grad_enabled() and (src.requires_grad or dst.requires_grad):
        return _SquareDistance.apply(src, dst)
    return _square_distance_raw(src, dst)


class _IndexPoints(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx_flat):
        B, N, C = points.shape
        S = idx_flat.shape[1]
        out = torch.empty(B, S, C, device=points.device, dtype=points.dtype)
        _lib.call("cmr_index_points", _lib.ptr(points), _lib.ptr(idx_flat), B, N, S, C * points.element_size(),
                  _lib.ptr(out), _lib.stream())
        ctx.save_for_backward(idx_flat)
        ctx.shape = (B, N, C)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx_flat,) = ctx.saved_tensors
        B, N, C = ctx.shape
        if grad_out.dtype != torch.float32:
            raise _lib.CmrError("index_points backward supports float32 only")
        grad_out = grad_out.contiguous()
        grad = torch.zeros(B, N, C, device=grad_out.device, dtype=torch.float32)
        _lib.call("cmr_index_points_backward", _lib.ptr(grad_out), _lib.ptr(idx_flat), B, N, idx_flat.shape[1], C,
                  _lib.ptr(grad), _lib.stream())
        return grad, None


def index_points(points, idx):
    """pointnet_util.py:36-47: points [B,N,C], idx [B,S] or [B,S,K] int64 -> [B,S,(K),C]."""
    _lib.require_cuda(points, "points")
    if points.dim() != 3:
        raise _lib.CmrError("points must be [B,N,C]")
    pts = points if points.is_contiguous() else points.contiguous()
    raw = idx.size()
    flat = _i64c(idx, "idx").reshape(raw[0], -1)
    if flat.shape[1] == 0:
        return pts.new_empty(*raw, pts.shape[-1])
    out = _IndexPoints.apply(pts, flat)
    return out.reshape(*raw, -1)


def farthest_point_sample(xyz, npoint):
    """pointnet_util.py:50-70: xyz [B,N,3] -> centroids [B,npoint] int64.  The first index is drawn
    with the reference's own call on the CPU generator (:62), so a seeded run picks the same seeds."""
    xyz = _f32c(xyz, "xyz")
    B, N, C = xyz.shape
    if C != 3:
        raise _lib.CmrError("farthest_point_sample expects xyz [B,N,3]")
    start = torch.randint(0, N, (B,), dtype=torch.long).to(xyz.device)
    return farthest_point_sample_from(xyz, npoint, start)


FPS_GRID_MAX_POINTS = 41900       # the pruned sampler keeps a cloud's running distances in one SM's shared memory


def farthest_point_sample_from(xyz, npoint, start, method="auto"):
    """Same as ``farthest_point_sample`` with explicit start indices [B] int64 (extension).
    method: "cluster" = every point every round, one thread-block cluster per cloud; "grid" = cell pruning on a uniform
    grid, one SM per cloud (same indices); "auto" = grid for 2048 <= N <= 41900."""
    xyz = _f32c(xyz, "xyz")
    B, N, _ = xyz.shape
    start = _i64c(start, "start")
    out = torch.empty(B, npoint, device=xyz.device, dtype=torch.int64)
    if npoint <= 0:
        return out
    if method not in ("auto", "grid", "cluster"):
        raise ValueError(method)
    if method == "grid" or (method == "auto" and KNN_GRID_MIN_POINTS <= N <= FPS_GRID_MAX_POINTS):
        ws = _grid_workspace(B, N, xyz.device)
        _lib.call("cmr_farthest_point_sample_grid", _lib.ptr(xyz), _lib.ptr(start), B, N, npoint, _lib.ptr(ws), _lib.ptr(out),
                  _lib.stream())
    else:
        _lib.call("cmr_farthest_point_sample", _lib.ptr(xyz), _lib.ptr(start), B, N, npoint, _lib.ptr(out),
                  _lib.stream())
    return out


_knn_ws = {}          # (device, bytes) -> workspace tensor of the grid searches (reused: the build is part of the call)
KNN_GRID_MIN_POINTS = 2048


def _grid_workspace(B, N, device):
    nbytes = _lib.load().cmr_knn_grid_workspace_bytes(B, N)
    key = (device, nbytes)
    ws = _knn_ws.get(key)
    if ws is None:
        if len(_knn_ws) > 4:
            _knn_ws.clear()
        ws = _knn_ws[key] = torch.empty(nbytes, dtype=torch.uint8, device=device)
    return ws


def query_ball_point(radius, nsample, xyz, new_xyz, method="auto"):
    """pointnet_util.py:73-93 -> group_idx [B,S,nsample] int64.
    method: "scan" = every query against every point; "grid" = only the cells of a uniform grid that the ball touches
    (same result); "auto" = grid from 2048 points on (nsample <= 128)."""
    xyz = _f32c(xyz, "xyz")
    new_xyz = _f32c(new_xyz, "new_xyz")
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    # `sqrdists > radius ** 2`: the python double is rounded to fp32 for the comparison
    r2 = float(torch.tensor(radius ** 2, dtype=torch.float32))
    n_eff = min(nsample, N)   # sort()[..., :nsample] cannot return more than N columns
    out = torch.empty(B, S, n_eff, device=xyz.device, dtype=torch.int64)
    if n_eff == 0 or S == 0:
        return out
    if method not in ("auto", "grid", "scan"):
        raise ValueError(method)
    grid_ok = n_eff <= 128 and radius >= 0
    if grid_ok and (method == "grid" or (method == "auto" and N >= KNN_GRID_MIN_POINTS)):
        ws = _grid_workspace(B, N, xyz.device)
        _lib.call("cmr_query_ball_point_grid", _lib.ptr(new_xyz), _lib.ptr(xyz), ctypes.c_float(r2),
                  ctypes.c_float(float(abs(radius))), n_eff, B, S, N, _lib.ptr(ws), _lib.ptr(out), _lib.stream())
    else:
        _lib.call("cmr_query_ball_point", _lib.ptr(new_xyz), _lib.ptr(xyz), ctypes.c_float(r2), n_eff, B, S, N,
                  _lib.ptr(out), _lib.stream())
    return out




def knn_point(k, xyz, new_xyz, method="auto"):
    """``square_distance(new_xyz, xyz).argsort()[:, :, :k]`` (pointnet_util.py:115-116) without the
    [B,S,N] matrix, ties ordered by index -> [B,S,min(k,N)] int64.

    method: "brute" = every query against every point (cmr_knn); "grid" = uniform grid over the cloud, rings of cells
    around the query (cmr_knn_grid; the same result bit for bit); "auto" = grid from 2048 points on."""
    xyz = _f32c(xyz, "xyz")
    new_xyz = _f32c(new_xyz, "new_xyz")
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    k_eff = min(k, N)
    if k_eff > 128:
        raise _lib.CmrError("knn_point supports k <= 128")
    out = torch.empty(B, S, k_eff, device=xyz.device, dtype=torch.int64)
    if k_eff == 0 or S == 0:
        return out
    if method not in ("auto", "grid", "brute"):
        raise ValueError(method)
    if method == "grid" or (method == "auto" and N >= KNN_GRID_MIN_POINTS):
        ws = _grid_workspace(B, N, xyz.device)
        _lib.call("cmr_knn_grid", _lib.ptr(new_xyz), _lib.ptr(xyz), B, S, N, k_eff, _lib.ptr(ws), _lib.ptr(out), _lib.stream())
    else:
        _lib.call("cmr_knn", _lib.ptr(new_xyz), _lib.ptr(xyz), B, S, N, k_eff, _lib.ptr(out), _lib.stream())
    return out


def _group_points_raw(xyz, points, new_xyz, idx):
    B, N, _ = xyz.shape
    _, S, K = idx.shape
    D = 0 if points is None else points.shape[-1]
    out = torch.empty(B, S, K, 3 + D, device=xyz.device, dtype=torch.float32)
    if out.numel():
        _lib.call("cmr_group_points", _lib.ptr(xyz), _lib.ptr(points), _lib.ptr(new_xyz), _lib.ptr(idx), B, N, S, K,
                  D, _lib.ptr(out), _lib.stream())
    return out


def _scatter_rows(grad_rows, idx_flat, N):
    """sum of grad_rows [B,S',C] into rows idx_flat [B,S'] of a zero [B,N,C] (cmr_index_points_backward)."""
    B, S, C = grad_rows.shape
    grad = torch.zeros(B, N, C, device=grad_rows.device, dtype=torch.float32)
    if S and C:
        _lib.call("cmr_index_points_backward", _lib.ptr(grad_rows.contiguous()), _lib.ptr(idx_flat), B, N, S, C,
                  _lib.ptr(grad), _lib.stream())
    return grad


class _GroupPoints(torch.autograd.Function):
    """Gradient of the fused gather - centroid || gather: what autograd gives the reference through
    index_points, the subtraction and the cat (pointnet_util.py:120-129)."""

    @staticmethod
    def forward(ctx, xyz, points, new_xyz, idx):
        ctx.save_for_backward(idx)
        ctx.n = xyz.shape[1]
        ctx.has_points = points is not None
        return _group_points_raw(xyz, points, new_xyz, idx)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        B, S, K = idx.shape
        flat = idx.reshape(B, S * K)
        g = g.float()
        gxyz = gpts = gnew = None
        if ctx.needs_input_grad[0]:
            gxyz = _scatter_rows(g[..., :3].reshape(B, S * K, 3), flat, ctx.n)
        if ctx.has_points and ctx.needs_input_grad[1]:
            gpts = _scatter_rows(g[..., 3:].reshape(B, S * K, -1), flat, ctx.n)
        if ctx.needs_input_grad[2]:
            gnew = -g[..., :3].sum(dim=2)
        return gxyz, gpts, gnew, None


def group_points(xyz, points, new_xyz, idx):
    """Fused tail of sample_and_group (pointnet_util.py:120-129):
    cat(xyz[idx] - new_xyz[:, :, None], points[idx]) -> [B,S,K,3+D]."""
    xyz = _f32c(xyz, "xyz")
    new_xyz = _f32c(new_xyz, "new_xyz")
    idx = _i64c(idx, "idx")
    if points is not None:
        points = _f32c(points, "points")
    if torch.is_grad_enabled() and (xyz.requires_grad or new_xyz.requires_grad or
                                    (points is not None and points.requires_grad)):
        return _GroupPoints.apply(xyz, points, new_xyz, idx)
    return _group_points_raw(xyz, points, new_xyz, idx)


def sample_and_group(npoint, radius, nsample, xyz, points, returnfps=False, knn=False):
    """pointnet_util.py:96-133 -> new_xyz [B,npoint,3], new_points [B,npoint,nsample,3+D]."""
    B, N, C = xyz.shape
    fps_idx = farthest_point_sample(xyz, npoint)
    new_xyz = index_points(xyz, fps_idx)
    if knn:
        idx = knn_point(nsample, xyz, new_xyz)
    else:
        idx = query_ball_point(radius, nsample, xyz, new_xyz)
    new_points = group_points(xyz, points, new_xyz, idx)
    if returnfps:
        grouped_xyz = index_points(xyz, idx)
        return new_xyz, new_points, grouped_xyz, fps_idx
    return new_xyz, new_points


def sample_and_group_all(xyz, points):
    """pointnet_util.py:136-153 (pure views/concat; nothing to accelerate)."""
    B, N, C = xyz.shape
    new_xyz = torch.zeros(B, 1, C, device=xyz.device)
    grouped_xyz = xyz.view(B, 1, N, C)
    if points is not None:
        new_points = torch.cat([grouped_xyz, points.view(B, 1, N, -1)], dim=-1)
    else:
        new_points = grouped_xyz
    return new_xyz, new_points
