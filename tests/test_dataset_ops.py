"""Dataset-side geometric operations: FarthestSampler.sample and the nearest node of every
point, float64.  CPU tests pin the oracle (oracle/dataset_oracle.py) against the real reference class and scipy
(build container only) and against the committed golden fixture; GPU tests compare the CUDA kernels, through the
C ABI and the drop-in class, with the oracle and the fixture.  Bar: indices BIT-EXACT, sampled points bit-exact."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle import dataset_oracle as do
from oracle import reference_loader

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden  # noqa: E402  (its dataset_inputs() regenerates the seeded inputs)

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataset.npz"))
CASES = ("kitti", "small", "ties")


def _inputs(case):
    pc, sub, k = make_golden.dataset_inputs(case)
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(pc).tobytes())
    h.update(np.ascontiguousarray(sub).tobytes())
    assert h.hexdigest().encode() == GOLDEN[case + "_sha"].tobytes(), "seeded inputs differ from the fixture's"
    return pc, sub, k


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_golden(case):
    pc, sub, k = _inputs(case)
    pts, idx = do.farthest_sample(sub, k, int(GOLDEN[case + "_init"]))
    assert np.array_equal(idx, GOLDEN[case + "_fps_idx"]) and np.array_equal(pts, GOLDEN[case + "_fps_pts"])
    if case != "ties":
        assert np.array_equal(do.nearest_index(pc, pts), GOLDEN[case + "_nearest"])


@pytest.mark.skipif(not reference_loader.available(), reason="needs /root/reference (build container only)")
def test_oracle_matches_reference_class_and_scipy():
    from scipy.spatial import cKDTree
    kd = reference_loader.kitti_dataset()
    rng = np.random.RandomState(5)
    for M, k in ((500, 40), (4096, 300)):
        pts = rng.randn(3, M) * np.array([[30.0], [2.0], [30.0]])
        np.random.seed(M)
        want_pts, want_idx = kd.FarthestSampler().sample(pts, k)
        got_pts, got_idx = do.farthest_sample(pts, k, int(want_idx[0]))
        assert np.array_equal(got_idx, want_idx) and np.array_equal(got_pts, want_pts)
        cloud = rng.randn(3, 5000) * 20.0
        assert np.array_equal(do.nearest_index(cloud, want_pts), cKDTree(want_pts.T).query(cloud.T, k=1)[1])


def test_dataset_ops_reject_cpu_only_hosts_and_bad_shapes():
    from cmr_agent_b200 import _lib, dataset_ops
    with pytest.raises(_lib.CmrError):
        dataset_ops.farthest_point_sample_batch(torch.zeros(1, 3, 8, dtype=torch.float64), 2,
                                                torch.zeros(1, dtype=torch.int64))
    with pytest.raises(_lib.CmrError):
        dataset_ops.FarthestSampler(dim=2)
    lib = _lib.load()
    assert lib.cmr_fps_f64(None, None, 1, 8, 2, None, None, None) == -1
    assert lib.cmr_nearest_f64(None, None, 1, 8, 2, None, None) == -1


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_matches_golden_and_oracle(cuda, case):
    from cmr_agent_b200 import dataset_ops
    pc, sub, k = _inputs(case)
    init = int(GOLDEN[case + "_init"])
    p = torch.from_numpy(sub).to(cuda).unsqueeze(0)
    idx, pts = dataset_ops.farthest_point_sample_batch(p, k, torch.tensor([init], device=cuda))
    assert np.array_equal(idx[0].cpu().numpy(), GOLDEN[case + "_fps_idx"])
    assert np.array_equal(pts[0].cpu().numpy(), GOLDEN[case + "_fps_pts"])
    near = dataset_ops.nearest_index_batch(torch.from_numpy(pc).to(cuda).unsqueeze(0), pts)
    want = GOLDEN[case + "_nearest"] if case != "ties" else do.nearest_index(pc, GOLDEN[case + "_fps_pts"])
    assert np.array_equal(near[0].cpu().numpy(), want)


@pytest.mark.gpu
def test_gpu_drop_in_class_consumes_numpy_rng_like_the_reference(cuda):
    from cmr_agent_b200 import dataset_ops
    rng = np.random.RandomState(11)
    pts = rng.randn(3, 2500) * 15.0
    np.random.seed(123)
    init = np.random.randint(3)            # KittiDataset.py:118 draws randint(len(pts)) = randint(3)
    after = np.random.randint(1 << 30)
    np.random.seed(123)
    got_pts, got_idx = dataset_ops.FarthestSampler().sample(pts, 100)
    assert np.random.randint(1 << 30) == after          # the generator advanced by exactly one draw
    want_pts, want_idx = do.farthest_sample(pts, 100, init)
    assert got_idx.dtype == np.int64 and np.array_equal(got_idx, want_idx) and np.array_equal(got_pts, want_pts)
    cloud = rng.randn(3, 7001) * 15.0
    assert np.array_equal(dataset_ops.nearest_index(cloud, got_pts), do.nearest_index(cloud, want_pts))


@pytest.mark.gpu
def test_gpu_batched_ragged_sizes(cuda):
    from cmr_agent_b200 import dataset_ops
    rng = np.random.RandomState(3)
    for B, M, k, N in ((3, 1025, 17, 999), (2, 16384, 64, 4100), (1, 5, 5, 3)):
        pts = rng.randn(B, 3, M) * 5.0
        start = rng.randint(0, M, size=B)
        idx, out = dataset_ops.farthest_point_sample_batch(torch.from_numpy(pts).to(cuda), k,
                                                           torch.from_numpy(start).to(cuda))
        cloud = rng.randn(B, 3, N) * 5.0
        near = dataset_ops.nearest_index_batch(torch.from_numpy(cloud).to(cuda), out)
        for b in range(B):
            wp, wi = do.farthest_sample(pts[b], k, int(start[b]))
            assert np.array_equal(idx[b].cpu().numpy(), wi) and np.array_equal(out[b].cpu().numpy(), wp)
            assert np.array_equal(near[b].cpu().numpy(), do.nearest_index(cloud[b], wp))
    from cmr_agent_b200 import _lib
    with pytest.raises(_lib.CmrError):      # more points than one CTA keeps distances for
        dataset_ops.farthest_point_sample_batch(torch.zeros(1, 3, 20000, dtype=torch.float64, device=cuda), 4,
                                                torch.zeros(1, dtype=torch.int64, device=cuda))
