"""GPU parity of the PointNet++ front-end kernels against the reference's golden outputs and the
oracle.  All outputs here are integers (indices) or exact copies/differences: BIT-EXACT, with the
tie-aware comparator of SURVEY.md A.7 where the reference itself is unstable (raw argsort)."""
import numpy as np
import pytest
import torch

from cmr_agent_b200 import synth
from oracle import cref, pointnet_oracle as po
from tests import helpers as hp

pytestmark = pytest.mark.gpu


def _pn():
    from cmr_agent_b200 import pointnet_util
    return pointnet_util


@pytest.mark.parametrize("tag,unique", [("dup", (2600, 2600)), ("plain", None)])
def test_front_end_matches_reference_golden(cuda, tag, unique):
    pn = _pn()
    g = hp.golden("pointnet")
    xyz_cpu = synth.make_cloud_batch(2, num_pt=4096, seed=hp.SEED, unique=unique)
    assert hp.sha(xyz_cpu) == str(g[f"xyz_sha_{tag}"])
    xyz = xyz_cpu.to(cuda)
    fps_want = torch.from_numpy(g[f"fps_{tag}"])
    fps = pn.farthest_point_sample_from(xyz, 128, fps_want[:, 0].to(cuda))
    assert torch.equal(fps.cpu(), fps_want)
    new_xyz = pn.index_points(xyz, fps)
    assert torch.equal(new_xyz.cpu(), torch.from_numpy(g[f"new_xyz_{tag}"]))
    d = pn.square_distance(new_xyz, xyz)
    assert torch.equal(d.cpu(), po.square_distance(new_xyz.cpu(), xyz_cpu))
    assert np.allclose(d.double().sum(-1).cpu().numpy(), g[f"sqdist_rowsum_{tag}"], rtol=1e-12)
    knn = pn.knn_point(16, xyz, new_xyz).cpu()
    assert torch.equal(knn, torch.from_numpy(g[f"knn16_stable_{tag}"]))
    assert po.knn_equivalent(torch.from_numpy(g[f"knn16_raw_{tag}"]), knn, new_xyz.cpu(), xyz_cpu)
    for r in (0.5, 2.0):
        ball = pn.query_ball_point(r, 32, xyz, new_xyz)
        assert torch.equal(ball.cpu(), torch.from_numpy(g[f"ball_{r}_{tag}"]))
    # sample_and_group: same CPU-RNG draw as the reference (pointnet_util.py:62)
    torch.manual_seed(hp.SEED + 1)
    nx, npts, gxyz, fidx = pn.sample_and_group(64, 1.5, 16, xyz, xyz * 0.5 + 1.0, returnfps=True)
    assert torch.equal(fidx.cpu(), torch.from_numpy(g[f"sag_fps_{tag}"]))
    assert torch.equal(nx.cpu(), torch.from_numpy(g[f"sag_new_xyz_{tag}"]))
    assert torch.equal(npts.cpu(), torch.from_numpy(g[f"sag_new_points_{tag}"]))
    assert torch.equal(gxyz.cpu(), po.index_points(xyz_cpu, pn.query_ball_point(1.5, 16, xyz, nx).cpu()))


def test_fps_full_size_matches_reference_golden(cuda):
    """BASELINE config 4 shape for one cloud: 40960 -> 1280, all 1280 indices identical to the reference."""
    pn = _pn()
    g = hp.golden("pointnet")
    xyz_cpu = synth.make_cloud_batch(1, num_pt=40960, seed=hp.SEED + 100)
    assert hp.sha(xyz_cpu) == str(g["xyz_sha_full"])
    want = torch.from_numpy(g["fps_full"])
    got = pn.farthest_point_sample_from(xyz_cpu.to(cuda), 1280, want[:, 0].to(cuda))
    assert torch.equal(got.cpu(), want)


@pytest.mark.parametrize("N,npoint,B,unique", [
    (100, 17, 3, None), (1000, 1000, 2, None), (2048, 300, 2, (700, 700)), (5000, 64, 5, None),
    (10240, 1280, 2, None), (20000, 128, 3, None), (40960, 256, 4, (26000, 34000)), (65536, 64, 2, None),
    (98304, 32, 1, None), (150000, 16, 1, None), (64, 100, 2, None),
])
def test_fps_sizes_vs_c_oracle(cuda, N, npoint, B, unique):
    """Every register/cluster configuration of k_fps, including npoint > #unique points (all-zero
    distance ties) and npoint > N."""
    pn = _pn()
    xyz_cpu = synth.make_cloud_batch(B, num_pt=N, seed=N, unique=unique)
    start = torch.randint(0, N, (B,), generator=torch.Generator().manual_seed(N))
    got = pn.farthest_point_sample_from(xyz_cpu.to(cuda), npoint, start.to(cuda)).cpu()
    for b in range(B):
        want = cref.fps(xyz_cpu[b].numpy(), npoint, int(start[b]))
        assert np.array_equal(got[b].numpy(), want), f"cloud {b}: first diff at {np.argmax(got[b].numpy() != want)}"


def test_fps_draws_start_like_the_reference(cuda):
    pn = _pn()
    xyz_cpu = synth.make_cloud_batch(3, num_pt=2048, seed=1)
    torch.manual_seed(99)
    want = po.farthest_point_sample(xyz_cpu, 50)
    torch.manual_seed(99)
    got = pn.farthest_point_sample(xyz_cpu.to(cuda), 50)
    assert got.dtype == torch.int64 and torch.equal(got.cpu(), want)


@pytest.mark.parametrize("N,S,k,unique", [
    (40960, 1280, 64, None),            # config 4 shape
    (40960, 333, 64, (26000, 34000)),   # duplicate-padded: every tie class matters
    (1280, 1280, 16, None),             # the live PointNN.KnnPointTransformer shape (PointNN.py:215-216)
    (5000, 77, 1, None), (300, 50, 128, None), (100, 9, 100, None), (33, 5, 32, None), (2048, 64, 33, (500, 500)),
])
def test_knn_vs_c_oracle(cuda, N, S, k, unique):
    pn = _pn()
    B = 2
    xyz_cpu = synth.make_cloud_batch(B, num_pt=N, seed=N + k, unique=unique)
    q_cpu = xyz_cpu[:, torch.randperm(N, generator=torch.Generator().manual_seed(k))[:S]].contiguous()
    got = pn.knn_point(k, xyz_cpu.to(cuda), q_cpu.to(cuda)).cpu()
    assert got.shape == (B, S, min(k, N)) and got.dtype == torch.int64
    for b in range(B):
        want = cref.knn(q_cpu[b].numpy(), xyz_cpu[b].numpy(), min(k, N))
        assert np.array_equal(got[b].numpy(), want)


def test_live_pointnn_pattern(cuda):
    """models/PointNN.py:213-221: permuted (strided) inputs, square_distance + argsort + index_points."""
    pn = _pn()
    g = torch.Generator().manual_seed(4)
    node = torch.randn(2, 3, 1280, generator=g)            # b x 3 x n as PointNN holds it
    feat = torch.randn(2, 64, 1280, generator=g)
    xyz_d = node.to(cuda).permute(0, 2, 1)                 # non-contiguous view, like :213
    f_d = feat.to(cuda).permute(0, 2, 1)
    d = pn.square_distance(xyz_d, xyz_d)
    want = po.square_distance(node.permute(0, 2, 1), node.permute(0, 2, 1))
    assert torch.equal(d.cpu(), want)
    idx = d.argsort()[:, :, :16]
    assert po.knn_equivalent(idx.cpu(), want.argsort(stable=True)[:, :, :16], node.permute(0, 2, 1), node.permute(0, 2, 1))
    assert torch.equal(pn.index_points(xyz_d, idx).cpu(), po.index_points(node.permute(0, 2, 1), idx.cpu()))
    assert torch.equal(pn.index_points(f_d, idx).cpu(), po.index_points(feat.permute(0, 2, 1), idx.cpu()))


@pytest.mark.parametrize("N,S,nsample,radius", [(40960, 1280, 32, 1.0), (4096, 100, 64, 0.2), (1000, 33, 8, 5.0),
                                                (500, 20, 16, 0.0), (31, 4, 40, 100.0)])
def test_ball_query_vs_c_oracle(cuda, N, S, nsample, radius):
    pn = _pn()
    xyz_cpu = synth.make_cloud_batch(2, num_pt=N, seed=N, unique=None)
    q_cpu = xyz_cpu[:, :S].contiguous().clone()
    q_cpu[:, -1] += 5000.0                                   # one query with nothing in radius -> all N
    got = pn.query_ball_point(radius, nsample, xyz_cpu.to(cuda), q_cpu.to(cuda)).cpu()
    n_eff = min(nsample, N)
    for b in range(2):
        want = cref.ball(q_cpu[b].numpy(), xyz_cpu[b].numpy(), radius, n_eff)
        assert np.array_equal(got[b].numpy(), want)
    if radius < 1000:
        assert bool((got[:, -1] == N).all())


def test_index_points_shapes_dtypes_and_faults(cuda):
    pn = _pn()
    from cmr_agent_b200 import _lib
    g = torch.Generator().manual_seed(0)
    for C, dtype in ((3, torch.float32), (64, torch.float32), (5, torch.float16), (7, torch.int64), (1, torch.uint8)):
        pts = (torch.randn(3, 500, C, generator=g) * 100).to(dtype)
        for shape in ((3, 40), (3, 10, 16), (3, 0)):
            idx = torch.randint(0, 500, shape, generator=g)
            got = pn.index_points(pts.to(cuda), idx.to(cuda))
            if idx.numel() == 0:          # the reference's reshape(-1) cannot infer C for an empty gather
                assert got.shape == (*shape, C) and got.dtype == dtype
                continue
            assert got.dtype == dtype and torch.equal(got.cpu(), po.index_points(pts, idx))
    assert _lib.take_fault() == 0
    bad = torch.tensor([[0, 500, 2]] * 3)
    out = pn.index_points(torch.ones(3, 500, 4).to(cuda), bad.to(cuda))
    assert _lib.take_fault() == 1 and float(out[:, 1].abs().sum()) == 0.0


def test_index_points_backward(cuda):
    pn = _pn()
    g = torch.Generator().manual_seed(2)
    pts = torch.randn(2, 300, 32, generator=g)
    idx = torch.randint(0, 300, (2, 50, 8), generator=g)     # repeated rows accumulate
    w = torch.randn(2, 50, 8, 32, generator=g)
    a = pts.clone().requires_grad_(True)
    (po.index_points(a, idx) * w).sum().backward()
    b = pts.to(cuda).requires_grad_(True)
    (pn.index_points(b, idx.to(cuda)) * w.to(cuda)).sum().backward()
    # a row's gradient is a float32 sum of up to ~10 terms added in atomic (arbitrary) order, here and in torch's own
    # CUDA backward: the bound is relative to the sum of the terms' magnitudes, not to the (possibly cancelling) result
    scale = torch.zeros_like(pts).index_put_((torch.arange(2).view(2, 1, 1).expand_as(idx), idx), w.abs(), accumulate=True)
    assert bool(((b.grad.cpu() - a.grad).abs() <= 2e-6 * scale + 1e-7).all())
    assert bool((b.grad.cpu()[scale == 0] == 0).all())


def test_group_points_and_knn_grouping(cuda):
    pn = _pn()
    xyz_cpu = synth.make_cloud_batch(2, num_pt=3000, seed=8)
    feats = torch.randn(2, 3000, 13, generator=torch.Generator().manual_seed(1))
    torch.manual_seed(5)
    a = po.sample_and_group(40, 1.0, 12, xyz_cpu, feats, knn_mode=True)
    torch.manual_seed(5)
    b = pn.sample_and_group(40, 1.0, 12, xyz_cpu.to(cuda), feats.to(cuda), knn=True)
    assert torch.equal(b[0].cpu(), a[0]) and torch.equal(b[1].cpu(), a[1])
    torch.manual_seed(5)
    c = pn.sample_and_group(40, 1.0, 12, xyz_cpu.to(cuda), None)
    torch.manual_seed(5)
    d = po.sample_and_group(40, 1.0, 12, xyz_cpu, None)
    assert torch.equal(c[1].cpu(), d[1])
    n1, p1 = pn.sample_and_group_all(xyz_cpu.to(cuda), feats.to(cuda))
    n2, p2 = po.sample_and_group_all(xyz_cpu, feats)
    assert torch.equal(n1.cpu(), n2) and torch.equal(p1.cpu(), p2)


def _lidar_like(B, N, seed, unique=None):
    return synth.make_cloud_batch(B, num_pt=N, seed=seed, unique=unique)


@pytest.mark.parametrize("case", ["kitti", "padded_duplicates", "uniform_cube", "one_point_repeated", "line", "tiny",
                                  "k_equals_n", "queries_outside", "k128"])
def test_grid_knn_equals_brute_force_bit_for_bit(cuda, case):
    """cmr_knn_grid (uniform grid + ring search) must return exactly what cmr_knn returns - the stable
    (distance, index) order - on clouds that stress the search: LiDAR slabs, duplicate-padded clouds (exact ties),
    a cube (the third axis is not binned), degenerate extents, queries far outside the cloud."""
    from cmr_agent_b200 import pointnet_util as pn
    g = torch.Generator().manual_seed(123)
    k, S = 16, 257
    if case == "kitti":
        xyz = _lidar_like(2, 40960, 7)
        k, S = 64, 640
    elif case == "padded_duplicates":
        xyz = _lidar_like(2, 8192, 9, unique=(2600, 2700))
    elif case == "uniform_cube":
        xyz = torch.rand(2, 6000, 3, generator=g) * 50 - 25
    elif case == "one_point_repeated":
        xyz = torch.ones(1, 3000, 3) * 3.25
    elif case == "line":
        xyz = torch.zeros(2, 5000, 3)
        xyz[:, :, 2] = torch.rand(2, 5000, generator=g) * 100
    elif case == "tiny":
        xyz = torch.rand(3, 40, 3, generator=g)
        k, S = 7, 40
    elif case == "k_equals_n":
        xyz = torch.rand(1, 100, 3, generator=g) * 10
        k, S = 100, 30
    elif case == "queries_outside":
        xyz = _lidar_like(1, 10000, 11)
    else:
        xyz = _lidar_like(1, 20000, 13)
        k = 128
    xyz = xyz.contiguous().to(cuda)
    if case == "queries_outside":
        q = (torch.rand(1, S, 3, generator=g) * 2000 - 1000).to(cuda)
    else:
        idx = torch.randint(0, xyz.shape[1], (xyz.shape[0], S), generator=g).to(cuda)
        q = pn.index_points(xyz, idx) + (torch.rand(xyz.shape[0], S, 3, generator=g).to(cuda) - 0.5) * 0.5
    a = pn.knn_point(k, xyz, q, method="brute")
    b = pn.knn_point(k, xyz, q, method="grid")
    torch.cuda.synchronize()
    assert a.shape == b.shape
    assert torch.equal(a, b), f"{case}: {(a != b).sum().item()} of {a.numel()} neighbour slots differ"


@pytest.mark.parametrize("case", ["kitti", "padded_duplicates", "cube", "zero_radius", "huge_radius", "nothing_in_radius", "tiny"])
def test_grid_ball_query_equals_the_scan(cuda, case):
    """cmr_query_ball_point_grid against cmr_query_ball_point (which the goldens pin): the first nsample indices in
    ascending order, padded with the first hit, N where nothing lies within the radius."""
    from cmr_agent_b200 import pointnet_util as pn
    g = torch.Generator().manual_seed(77)
    radius, nsample, S = 1.0, 32, 300
    if case == "kitti":
        xyz = _lidar_like(2, 40960, 17)
        radius, nsample, S = 2.0, 64, 640
    elif case == "padded_duplicates":
        xyz = _lidar_like(2, 8192, 19, unique=(2600, 2700))
        radius = 1.5
    elif case == "cube":
        xyz = torch.rand(2, 6000, 3, generator=g) * 20 - 10
        radius, nsample = 2.5, 128
    elif case == "zero_radius":
        xyz = _lidar_like(1, 5000, 23)
        radius = 0.0
    elif case == "huge_radius":
        xyz = _lidar_like(1, 5000, 29)
        radius, nsample = 1e4, 48
    elif case == "nothing_in_radius":
        xyz = _lidar_like(1, 5000, 31)
        radius = 1e-3
    else:
        xyz = torch.rand(2, 50, 3, generator=g)
        radius, nsample, S = 0.4, 16, 50
    xyz = xyz.contiguous().to(cuda)
    idx = torch.randint(0, xyz.shape[1], (xyz.shape[0], S), generator=g).to(cuda)
    q = pn.index_points(xyz, idx)
    if case in ("nothing_in_radius",):
        q = q + 500.0
    elif case not in ("zero_radius",):
        q = q + (torch.rand(xyz.shape[0], S, 3, generator=g).to(cuda) - 0.5) * 0.3
    a = pn.query_ball_point(radius, nsample, xyz, q, method="scan")
    b = pn.query_ball_point(radius, nsample, xyz, q, method="grid")
    torch.cuda.synchronize()
    assert torch.equal(a, b), f"{case}: {(a != b).sum().item()} of {a.numel()} slots differ"
    if case == "nothing_in_radius":
        assert bool((a == xyz.shape[1]).all())


@pytest.mark.parametrize("case", ["kitti", "padded_duplicates", "cube", "one_point_repeated", "line", "small", "all_points"])
def test_grid_fps_equals_the_cluster_kernel(cuda, case):
    """cmr_farthest_point_sample_grid (cell pruning) against cmr_farthest_point_sample (pinned by the goldens): the
    same indices in the same order, including ties (duplicate-padded clouds) and degenerate extents."""
    from cmr_agent_b200 import pointnet_util as pn
    g = torch.Generator().manual_seed(5)
    npoint = 256
    if case == "kitti":
        xyz = _lidar_like(3, 40960, 41)
        npoint = 1280
    elif case == "padded_duplicates":
        xyz = _lidar_like(2, 8192, 43, unique=(2600, 2700))
        npoint = 3000                                   # more samples than unique points: zero distances, index ties
    elif case == "cube":
        xyz = torch.rand(2, 6000, 3, generator=g) * 50 - 25
    elif case == "one_point_repeated":
        xyz = torch.ones(1, 3000, 3) * 3.25
        npoint = 64
    elif case == "line":
        xyz = torch.zeros(2, 5000, 3)
        xyz[:, :, 2] = torch.rand(2, 5000, generator=g) * 100
    elif case == "small":
        xyz = torch.rand(4, 300, 3, generator=g) * 10
        npoint = 128
    else:
        xyz = _lidar_like(1, 2500, 47)
        npoint = 2500
    xyz = xyz.contiguous().to(cuda)
    start = torch.randint(0, xyz.shape[1], (xyz.shape[0],), generator=g).to(cuda)
    a = pn.farthest_point_sample_from(xyz, npoint, start, method="cluster")
    b = pn.farthest_point_sample_from(xyz, npoint, start, method="grid")
    torch.cuda.synchronize()
    assert torch.equal(a, b), f"{case}: first difference at column {int((a != b).any(dim=0).float().argmax())}"
