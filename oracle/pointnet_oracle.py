"""torch-CPU restatement ("port") of the reference PointNet++ utilities.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows
/root/reference/models/pointnet_util.py; each function cites its lines.
Pinned against the real module by tests/test_oracle_vs_reference.py and the
fixtures in tests/golden/ (the reference itself ships no tests).

kNN note: the reference takes ``argsort()[:, :, :k]`` which is
an UNSTABLE sort, so on exact distance ties (duplicate-padded clouds) its order
is arbitrary.  ``knn`` here takes ``stable=True`` (order = (distance, index));
the tie-aware comparator ``knn_equivalent`` accepts any order within an
equal-distance class and is what the raw reference output is checked with.
"""
import torch


def square_distance(src, dst):
    """pointnet_util.py:19-33: direct differences, sum over the last axis -> [B,S,N]."""
    return torch.sum((src[:, :, None] - dst[:, None]) ** 2, dim=-1)


def index_points(points, idx):
    """pointnet_util.py:36-47: batched row gather, idx [B,S] or [B,S,K] -> [B,S,(K),C]."""
    shape = idx.size()
    flat = idx.reshape(shape[0], -1)
    out = torch.gather(points, 1, flat[..., None].expand(-1, -1, points.size(-1)))
    return out.reshape(*shape, -1)


def farthest_point_sample(xyz, npoint, start=None):
    """pointnet_util.py:50-70.  ``start`` [B] int64 replaces the CPU-RNG draw of :62 so the
    caller can share one draw between oracle and device (same call when omitted)."""
    B, N, _ = xyz.shape
    centroids = torch.zeros(B, npoint, dtype=torch.long)
    distance = torch.ones(B, N) * 1e10
    farthest = torch.randint(0, N, (B,), dtype=torch.long) if start is None else start.clone()
    rows = torch.arange(B, dtype=torch.long)
    for i in range(npoint):
        centroids[:, i] = farthest
        c = xyz[rows, farthest, :].view(B, 1, 3)
        d = torch.sum((xyz - c) ** 2, -1)
        distance = torch.min(distance, d)
        farthest = torch.max(distance, -1)[1]
    return centroids


def query_ball_point(radius, nsample, xyz, new_xyz):
    """pointnet_util.py:73-93: first ``nsample`` in-radius indices (ascending), padded with the
    first hit; N everywhere when nothing is in radius."""
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    group = torch.arange(N, dtype=torch.long).view(1, 1, N).repeat([B, S, 1])
    d = square_distance(new_xyz, xyz)
    group[d > radius ** 2] = N
    group = group.sort(dim=-1)[0][:, :, :nsample]
    first = group[:, :, 0].view(B, S, 1).repeat([1, 1, nsample])
    pad = group == N
    group[pad] = first[pad]
    return group


def knn(query, ref, k, stable=True):
    """pointnet_util.py:115-116 / PointNN.py:215-216: k smallest of square_distance rows."""
    d = square_distance(query, ref)
    return d.argsort(dim=-1, stable=stable)[:, :, :k]


def knn_equivalent(idx_a, idx_b, query, ref):
    """True when two [B,S,k] index sets are equal up to permutation inside equal-distance classes
    (SURVEY.md A.7): the sorted distance rows must be bit-identical and each class a set match."""
    if idx_a.shape != idx_b.shape:
        return False
    N = ref.shape[1]
    for b in range(query.shape[0]):
        d = square_distance(query[b:b + 1], ref[b:b + 1])[0]
        da = torch.gather(d, 1, idx_a[b])
        db = torch.gather(d, 1, idx_b[b])
        # both are ascending selections: the distance rows must agree bit for bit
        if not torch.equal(da, db):
            return False
        # members strictly closer than the k-th distance are forced; the boundary class
        # (distance == k-th distance) may be any members of that class
        forced = da < da[:, -1:]
        ia = torch.where(forced, idx_a[b], torch.full_like(idx_a[b], N)).sort(dim=1)[0]
        ib = torch.where(forced, idx_b[b], torch.full_like(idx_b[b], N)).sort(dim=1)[0]
        if not torch.equal(ia, ib):
            return False
        # no duplicates inside a row
        if (idx_b[b].sort(dim=1)[0].diff(dim=1) == 0).any():
            return False
    return True


def sample_and_group(npoint, radius, nsample, xyz, points, returnfps=False, knn_mode=False, start=None):
    """pointnet_util.py:96-133."""
    B, N, C = xyz.shape
    S = npoint
    fps_idx = farthest_point_sample(xyz, npoint, start=start)
    new_xyz = index_points(xyz, fps_idx)
    if knn_mode:
        idx = knn(new_xyz, xyz, nsample)
    else:
        idx = query_ball_point(radius, nsample, xyz, new_xyz)
    grouped_xyz = index_points(xyz, idx)
    grouped_xyz_norm = grouped_xyz - new_xyz.view(B, S, 1, C)
    if points is not None:
        new_points = torch.cat([grouped_xyz_norm, index_points(points, idx)], dim=-1)
    else:
        new_points = grouped_xyz_norm
    if returnfps:
        return new_xyz, new_points, grouped_xyz, fps_idx
    return new_xyz, new_points


def sample_and_group_all(xyz, points):
    """pointnet_util.py:136-153."""
    B, N, C = xyz.shape
    new_xyz = torch.zeros(B, 1, C)
    grouped = xyz.view(B, 1, N, C)
    if points is not None:
        return new_xyz, torch.cat([grouped, points.view(B, 1, N, -1)], dim=-1)
    return new_xyz, grouped
