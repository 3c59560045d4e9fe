"""Pins the oracle (torch port + C restatement) against the REAL reference, loaded by file path.
Runs only where /root/reference exists (the build container); the golden fixtures carry the same
pin to the GPU box."""
import numpy as np
import pytest
import torch

from cmr_agent_b200 import synth
from oracle import cref, env_oracle as eo, pointnet_oracle as po, reference_loader as rl

pytestmark = pytest.mark.skipif(not rl.available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def env():
    return rl.environment()


@pytest.fixture(scope="module")
def pn():
    return rl.pointnet_util()


@pytest.mark.parametrize("shape", [dict(num_pt=4096, img_h=64, img_w=256), dict(num_pt=1531, img_h=36, img_w=100)])
@pytest.mark.parametrize("dof6", [False, True])
def test_rollout_port_and_c_match_reference(env, shape, dof6):
    B, iters = 2, 3
    data = synth.make_batch(B, seed=11, **shape)
    cfg = synth.StepConfig(is_6_DoF=dof6)
    a_r, a_t = synth.make_actions(B, iters, seed=11, dof6=dof6)
    H, W = shape["img_h"] // 4, shape["img_w"] // 4
    ps, pt = env.init(data)
    qs, qt = eo.init(data)
    assert torch.equal(ps, qs) and torch.equal(pt, qt)
    assert torch.equal(env.to_disentangled(pt, data["pc"]), eo.to_disentangled(qt, data["pc"]))
    mean = data["pc"].mean(dim=2)
    prev_r = prev_o = None
    for it in range(iters):
        o2, o3 = env.observation_from_a_pose(data, ps)
        p2, p3 = eo.observation_from_a_pose(data, qs)
        assert torch.equal(o2, p2) and torch.equal(o3, p3)
        idx, inc = eo.projected_pixels(data, qs)
        for b in range(B):
            ci, cc = cref.project(data["pc"][b].numpy(), mean[b].numpy(), qs[b].numpy(), data["K"][b].numpy(), H, W)
            assert np.array_equal(ci, idx[b].numpy())
            assert np.array_equal(cc.astype(np.float32), o3[b, 4].numpy())
            sm = cref.scatter_mean(data["pc_geo_feat"][b].numpy(), data["pc_overlap_pred"][b].numpy(), ci, H * W)
            # sequential point-order sums reproduce torch's CPU scatter_add_ bit for bit
            assert np.array_equal(sm.reshape(64, H, W), o2[b, 64:].numpy())
        before = qs.clone()
        env.step(a_r[it], a_t[it], ps, cfg)
        eo.step(a_r[it], a_t[it], qs, cfg)
        assert torch.equal(ps, qs)
        # C restatement of the 3x3 FMA chains, fed the reference's own per-axis matrices
        move_r = torch.zeros(B, 3)
        move_t = torch.zeros(B, 3)
        if dof6:
            for ax in range(3):
                move_r[:, ax] = cfg.r_steps[a_r[it][:, ax]]
                move_t[:, ax] = cfg.t_steps[a_t[it][:, ax]]
        else:
            move_r[:, 1] = cfg.r_steps[a_r[it][:, 0]]
            move_t[:, 0] = cfg.t_steps[a_t[it][:, 0]]
            move_t[:, 2] = cfg.t_steps[a_t[it][:, 1]]
        Rn = env.euler_angles_to_matrix(move_r, "XYZ")
        for b in range(B):
            got = cref.apply_step(before[b].numpy(), Rn[b].numpy(), move_t[b].numpy())
            assert np.array_equal(got, ps[b].numpy())
        r1, d1 = env.reward(ps, data, prev_r)
        r2, d2 = eo.reward(qs, data, prev_o)
        assert torch.equal(r1, r2) and torch.equal(d1, d2)
        for b in range(B):
            c = cref.p2p(data["pc_in_cam_space"][b].numpy(), data["pc"][b].numpy(), data["pc_mask"][b].numpy(),
                         mean[b].numpy(), qs[b].numpy(), 0)
            assert abs(c - float(d1[b])) <= 1e-5 * abs(c)
        prev_r, prev_o = d1, d2


def test_expert_port_matches_reference(env):
    B = 16
    data = synth.make_batch(B, seed=3, num_pt=512, img_h=32, img_w=64)
    for dof6 in (False, True):
        cfg = synth.StepConfig(is_6_DoF=dof6)
        ps, pt = env.init(data)
        env.to_disentangled(pt, data["pc"])
        a = env.expert(ps, pt, cfg, data)
        b = eo.expert(ps, pt, cfg, data)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


@pytest.mark.parametrize("unique", [None, (2600, 2600)])
def test_pointnet_port_and_c_match_reference(pn, unique):
    xyz = synth.make_cloud_batch(2, num_pt=4096, seed=5, unique=unique)
    B = xyz.shape[0]
    torch.manual_seed(1)
    f1 = pn.farthest_point_sample(xyz, 96)
    torch.manual_seed(1)
    f2 = po.farthest_point_sample(xyz, 96)
    assert torch.equal(f1, f2)
    for b in range(B):
        assert np.array_equal(cref.fps(xyz[b].numpy(), 96, int(f1[b, 0])), f1[b].numpy())
    new_xyz = pn.index_points(xyz, f1)
    assert torch.equal(new_xyz, po.index_points(xyz, f1))
    d = pn.square_distance(new_xyz, xyz)
    assert torch.equal(d, po.square_distance(new_xyz, xyz))
    assert np.array_equal(cref.sqdist(new_xyz[0].numpy(), xyz[0].numpy()), d[0].numpy())
    raw = d.argsort()[:, :, :16]
    stable = po.knn(new_xyz, xyz, 16)
    assert po.knn_equivalent(raw, stable, new_xyz, xyz)
    for b in range(B):
        assert np.array_equal(cref.knn(new_xyz[b].numpy(), xyz[b].numpy(), 16), stable[b].numpy())
    for r in (0.0, 0.5, 2.0):
        q1 = pn.query_ball_point(r, 24, xyz, new_xyz)
        assert torch.equal(q1, po.query_ball_point(r, 24, xyz, new_xyz))
        for b in range(B):
            assert np.array_equal(cref.ball(new_xyz[b].numpy(), xyz[b].numpy(), r, 24), q1[b].numpy())
    far = new_xyz + 1000.0          # nothing in radius -> every slot is N (pointnet_util.py:88-92)
    q = pn.query_ball_point(0.5, 8, xyz, far)
    assert bool((q == xyz.shape[1]).all())
    assert np.array_equal(cref.ball(far[0].numpy(), xyz[0].numpy(), 0.5, 8), q[0].numpy())
    for knn_mode in (False, True):
        torch.manual_seed(3)
        a = pn.sample_and_group(48, 1.0, 16, xyz, xyz * 2, knn=knn_mode)
        torch.manual_seed(3)
        b_ = po.sample_and_group(48, 1.0, 16, xyz, xyz * 2, knn_mode=knn_mode)
        assert torch.equal(a[0], b_[0])
        if not knn_mode or unique is None:
            assert torch.equal(a[1], b_[1])


def test_knn_comparator_rejects_wrong_neighbours():
    xyz = synth.make_cloud_batch(1, num_pt=512, seed=9)
    q = xyz[:, :8]
    good = po.knn(q, xyz, 8)
    bad = good.clone()
    bad[0, 0, 3] = (bad[0, 0, 3] + 200) % 512
    assert po.knn_equivalent(good, good, q, xyz)
    assert not po.knn_equivalent(good, bad, q, xyz)


def test_small_product_regime_of_cpu_bmm(env):
    """torch's CPU bmm is the plain unfused loop when rows*cols*k < 400 (3 x n, n <= 44) and the FMA
    chain above (measured here; DESIGN.md 'arithmetic contract').  The C restatement follows both."""
    cfg = synth.StepConfig(is_6_DoF=True)
    g = torch.Generator().manual_seed(0)
    # (a) per-axis matrices composed as (Rx @ Ry) @ Rz and R <- Rnew @ R
    ang = torch.rand(64, 3, generator=g) * 6 - 3
    Rn = env.euler_angles_to_matrix(ang, "XYZ")
    for b in range(64):
        mats = [env._axis_angle_rotation(a, ang[b, i]) for i, a in enumerate("XYZ")]
        assert np.array_equal(cref.compose_xyz(*[m.numpy() for m in mats]), Rn[b].numpy())
    # (b) to_disentangled
    poses = torch.eye(4).repeat(64, 1, 1)
    poses[:, :3, :3] = Rn
    poses[:, :3, 3] = torch.randn(64, 3, generator=g) * 7
    pcd = torch.randn(64, 3, 200, generator=g) * 20 + 3
    mean = pcd.mean(dim=2)
    want = env.to_disentangled(poses.clone(), pcd)
    for b in range(64):
        assert np.array_equal(cref.to_disentangled(poses[b].numpy(), mean[b].numpy()), want[b].numpy())
    # (c) clouds of 44 / 45 points sit on either side of the regime switch
    for n, fused in ((44, False), (45, True), (7, False)):
        data = synth.make_batch(2, seed=21, num_pt=n, img_h=160, img_w=512)
        data["pc_overlap_pred"][:] = True
        pose = poses[:2].clone()
        pose[:, :3, 3] *= 0.1
        o2, o3 = env.observation_from_a_pose(data, pose)
        m = data["pc"].mean(dim=2)
        for b in range(2):
            ci, cc = cref.project(data["pc"][b].numpy(), m[b].numpy(), pose[b].numpy(), data["K"][b].numpy(), 40, 128)
            assert np.array_equal(cc.astype(np.float32), o3[b, 4].numpy())
            sm = cref.scatter_mean(data["pc_geo_feat"][b].numpy(), np.ones(n, np.uint8), ci, 5120)
            assert np.array_equal(sm.reshape(64, 40, 128), o2[b, 64:].numpy())
