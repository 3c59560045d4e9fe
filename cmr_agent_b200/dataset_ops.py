"""Drop-ins for the two geometric operations the reference's datasets run per sample on the host
(SURVEY.md section 8f, rank 3): farthest-point sampling of the node set and the nearest node of every point.

    FarthestSampler.sample      dataset/KittiDataset.py:107-126  (= dataset/NuScenesDataset.py:25-44)
    nearest_index(points, nodes) replaces ``cKDTree(node_np.T).query(pc.T, k=1)[1]`` (KittiDataset.py:365-366,
                                 NuScenesDataset.py:284-285)

Both compute in float64 like the numpy / scipy originals and return the same indices.  ``FarthestSampler`` keeps
the reference's signature (numpy in, numpy out, the start index drawn from ``np.random`` exactly as :118 does);
the ``*_batch`` functions take CUDA tensors for callers that collate first and sample on the device.
There is no CPU fallback: without a GPU these raise ``CmrError``.
"""
import numpy as np
import torch

from . import _lib


def _device():
    if not torch.cuda.is_available():
        raise _lib.CmrError("cmr_agent_b200.dataset_ops needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def farthest_point_sample_batch(pts, k, start):
    """pts [B,3,M] f64 CUDA, start [B] i64 CUDA -> (idx [B,k] i64, sampled [B,3,k] f64), KittiDataset.py:115-126."""
    pts = _lib.require_cuda(pts, "pts", torch.float64).contiguous()
    start = _lib.require_cuda(start, "start", torch.int64).contiguous()
    B, three, M = pts.shape
    if three != 3 or start.shape != (B,):
        raise _lib.CmrError("pts must be [B,3,M] and start [B]")
    idx = torch.empty(B, k, dtype=torch.int64, device=pts.device)
    out = torch.empty(B, 3, k, dtype=torch.float64, device=pts.device)
    _lib.call("cmr_fps_f64", _lib.ptr(pts), _lib.ptr(start), B, M, int(k), _lib.ptr(idx), _lib.ptr(out), _lib.stream())
    return idx, out


def nearest_index_batch(points, nodes):
    """points [B,3,N], nodes [B,3,S] f64 CUDA -> [B,N] i64: index of the nearest node of every point."""
    points = _lib.require_cuda(points, "points", torch.float64).contiguous()
    nodes = _lib.require_cuda(nodes, "nodes", torch.float64).contiguous()
    B, three, N = points.shape
    if three != 3 or nodes.shape[0] != B or nodes.shape[1] != 3:
        raise _lib.CmrError("points must be [B,3,N] and nodes [B,3,S]")
    out = torch.empty(B, N, dtype=torch.int64, device=points.device)
    _lib.call("cmr_nearest_f64", _lib.ptr(points), _lib.ptr(nodes), B, N, nodes.shape[2], _lib.ptr(out), _lib.stream())
    return out


class FarthestSampler:
    """dataset/KittiDataset.py:107-126 with the same constructor, method and return values."""

    def __init__(self, dim=3):
        if dim != 3:
            raise _lib.CmrError("FarthestSampler: only dim=3 (the reference never uses another)")
        self.dim = dim

    def calc_distances(self, p0, points):
        return ((p0 - points) ** 2).sum(axis=0)

    def sample(self, pts, k):
        """pts [3,M] numpy -> (farthest_pts [3,k] f64, farthest_pts_idx [k] i64)."""
        dev = _device()
        # :118 draws ``np.random.randint(len(pts))`` - len of a [3,M] array is 3; kept as it is, so that the
        # global numpy generator advances exactly like the reference's
        init_idx = np.random.randint(len(pts))
        p = torch.from_numpy(np.ascontiguousarray(pts, dtype=np.float64)).to(dev).unsqueeze(0)
        start = torch.tensor([init_idx], dtype=torch.int64, device=dev)
        idx, out = farthest_point_sample_batch(p, k, start)
        return out[0].cpu().numpy(), idx[0].cpu().numpy()


def nearest_index(points, nodes):
    """points [3,N], nodes [3,S] numpy -> [N] int64 (what ``cKDTree(nodes.T).query(points.T, k=1)[1]`` returns)."""
    dev = _device()
    p = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float64)).to(dev).unsqueeze(0)
    n = torch.from_numpy(np.ascontiguousarray(nodes, dtype=np.float64)).to(dev).unsqueeze(0)
    return nearest_index_batch(p, n)[0].cpu().numpy()
