"""Size-independent properties at BASELINE.json's full sizes (where the torch/C oracle would take minutes):
conservation ("checksum of checksums") for the scatter-mean, idempotence/determinism, sortedness and
uniqueness for the index kernels."""
import pytest
import torch

from cmr_agent_b200 import synth
from tests import helpers as hp

pytestmark = pytest.mark.gpu


def test_full_batch_observation_properties(cuda):
    """Config 2 shapes: 32 KITTI episodes.  (1) two calls give identical bits (deterministic, no atomics on
    floats); (2) conservation: sum over pixels of mean*count == sum of the features of the visible
    predicted-overlap points, per episode and channel; (3) empty pixels are exactly zero; (4) the image half
    and obs3d are exact copies."""
    from cmr_agent_b200 import environment as env
    B = 32
    cpu = synth.make_batch(B, seed=2023, num_pt=40960, img_h=160, img_w=512)
    data = hp.to_device(cpu, cuda)
    cfg = synth.StepConfig(device=cuda)
    a_r, a_t = synth.make_actions(B, 3, seed=1)
    pose, _ = env.init(data)
    for it in range(3):
        o2, o3, pix, mvis = env.observation_from_a_pose(data, pose, return_pixels=True)
        p2, p3 = env.observation_from_a_pose(data, pose)
        assert torch.equal(o2, p2) and torch.equal(o3, p3)
        P = 40 * 128
        vis = data["pc_overlap_pred"] & (pix < P)
        assert torch.equal(vis.sum(dim=1).int(), mvis)
        assert torch.equal(o3[:, 4], (pix < P).float())
        assert torch.equal(o2[:, :64], data["img_geo_feat"]) and torch.equal(o3[:, :3], data["pc"])
        cnt = torch.zeros(B, P + 1, device=cuda)
        cnt.scatter_add_(1, torch.where(vis, pix, torch.full_like(pix, P)).long(), vis.float())
        cnt = cnt[:, :P]
        proj = o2[:, 64:].reshape(B, 64, P)
        assert float(proj[(cnt == 0).unsqueeze(1).expand_as(proj)].abs().max()) == 0.0
        lhs = (proj.double() * cnt.unsqueeze(1).double()).sum(dim=2)                       # [B,64]
        rhs = (data["pc_geo_feat"].double() * vis.unsqueeze(1).double()).sum(dim=2)         # [B,64]
        assert hp.rel_err(lhs.cpu(), rhs.cpu(), floor=1e-5) <= 1e-5
        env.step(a_r[it].to(cuda), a_t[it].to(cuda), pose, cfg)


def test_config4_front_end_properties(cuda):
    """Config 4 shapes (reduced batch): FPS 40960 -> 1280 gives 1280 distinct indices starting at the seed and
    a non-increasing selection distance; kNN k=64 rows are sorted by (distance, index), unique, and contain
    the query itself first; ball-query rows are ascending up to the padding."""
    from cmr_agent_b200 import pointnet_util as pn
    B, N, S, K = 8, 40960, 1280, 64
    xyz = synth.make_cloud_batch(B, num_pt=N, seed=99).to(cuda)
    start = torch.arange(B, device=cuda) * 977 % N
    fps = pn.farthest_point_sample_from(xyz, S, start)
    assert fps.shape == (B, S) and torch.equal(fps[:, 0], start)
    assert all(len(set(fps[b].tolist())) == S for b in range(B))
    new_xyz = pn.index_points(xyz, fps)
    # distance of every selected point to the set selected before it never increases
    for b in range(2):
        d = pn.square_distance(new_xyz[b:b + 1], new_xyz[b:b + 1])[0]
        lower = torch.tril(torch.ones_like(d), diagonal=-1).bool()
        dmin = torch.where(lower, d, torch.full_like(d, float("inf"))).min(dim=1)[0][1:]
        assert bool((dmin[1:] <= dmin[:-1]).all())
    knn = pn.knn_point(K, xyz, new_xyz)
    assert knn.shape == (B, S, K) and knn.dtype == torch.int64
    assert torch.equal(knn[:, :, 0], fps)                       # distance 0 to itself, lowest index of any tie
    nb = pn.index_points(xyz, knn)
    dist = ((nb - new_xyz[:, :, None]) ** 2)
    dist = (dist[..., 0] + dist[..., 1]) + dist[..., 2]
    assert bool((dist[:, :, 1:] >= dist[:, :, :-1]).all())
    tie = dist[:, :, 1:] == dist[:, :, :-1]
    assert bool((knn[:, :, 1:][tie] > knn[:, :, :-1][tie]).all())
    assert bool((knn.sort(dim=-1)[0].diff(dim=-1) != 0).all())
    ball = pn.query_ball_point(1.0, 32, xyz, new_xyz)
    first = ball[:, :, :1]
    assert bool(((ball[:, :, 1:] > ball[:, :, :-1]) | (ball[:, :, 1:] == first)).all())
    g = pn.index_points(xyz, ball)
    gd = ((g - new_xyz[:, :, None]) ** 2).sum(-1)
    assert bool((gd <= 1.0 + 1e-6).all())


def test_episode_sharding_is_invariant(cuda):
    """Config 3 shape (NuScenes-like, duplicate-padded clouds, 40x80 grid): running the episodes in shards
    (what each of N GPUs does, cmr_agent_b200.dist.shard_range) gives bit-identical per-episode outputs
    to running the whole batch, for every world size; the all-reduced metric sums agree as well."""
    from cmr_agent_b200 import dist as cdist
    from cmr_agent_b200 import environment as env
    total, iters = 16, 3
    shape = dict(num_pt=40960, img_h=160, img_w=320, unique=(26000, 34000))
    cpu = synth.make_batch(total, seed=7, **shape)
    cfg = synth.StepConfig(device=cuda)
    a_r, a_t = synth.make_actions(total, iters, seed=7)

    def rollout(lo, hi):
        data = {k: (v[lo:hi] if isinstance(v, torch.Tensor) else v) for k, v in cpu.items()}
        data = hp.to_device(data, cuda)
        pose, target = env.init(data)
        env.to_disentangled(target, data["pc"])
        outs = []
        for it in range(iters):
            o2, o3 = env.observation_from_a_pose(data, pose)
            env.step(a_r[it, lo:hi].to(cuda), a_t[it, lo:hi].to(cuda), pose, cfg)
            _, dist = env.reward(pose, data, None)
            outs.append((o2.cpu(), o3.cpu(), dist.cpu()))
        err_t = (pose[:, :3, 3] - target[:, :3, 3]).norm(dim=1).cpu()
        return outs, pose.cpu(), err_t

    whole, pose_whole, err_whole = rollout(0, total)
    ref = cdist.MetricSums().add(err_whole * 0 + 1.0, err_whole).summary()
    for world in (2, 8):
        sums = cdist.MetricSums()
        for rank in range(world):
            lo, hi = cdist.shard_range(total, rank, world)
            part, pose_part, err_part = rollout(lo, hi)
            assert torch.equal(pose_part, pose_whole[lo:hi])
            for it in range(iters):
                for a, b in zip(part[it], whole[it]):
                    assert torch.equal(a, b[lo:hi])
            sums.add(err_part * 0 + 1.0, err_part)        # what each rank would contribute before the all-reduce
        got = sums.summary()
        assert got["episodes"] == total and abs(got["rte_mean"] - ref["rte_mean"]) < 1e-9
