"""The oracle (torch port + C restatement) against the golden fixtures the REAL reference produced
(tests/golden/make_golden.py).  Runs everywhere, GPU box included, without /root/reference."""
import numpy as np
import pytest
import torch

from cmr_agent_b200 import synth
from oracle import cref, env_oracle as eo, pointnet_oracle as po
from tests import helpers as hp


@pytest.mark.parametrize("name", list(hp.ENV_CASES))
def test_env_oracle_reproduces_golden(name):
    data, g, iters, full, shape = hp.env_inputs(name)
    H, W = shape["img_h"] // 4, shape["img_w"] // 4
    cfg = synth.StepConfig()
    B = data["pc"].shape[0]
    mean = data["pc"].mean(dim=2)
    assert np.array_equal(mean.numpy(), g["mean"])
    pose, target = eo.init(data)
    assert np.array_equal(eo.to_disentangled(target, data["pc"]).numpy(), g["pose_target_disentangled"])
    for b in range(B):
        assert np.array_equal(cref.to_disentangled(data["P"][b].numpy(), mean[b].numpy()),
                              g["pose_target_disentangled"][b])
    prev = None
    for it in range(iters):
        assert np.array_equal(pose.numpy(), g[f"pose_{it}"])
        o2, o3 = eo.observation_from_a_pose(data, pose)
        idx, inc = eo.projected_pixels(data, pose)
        assert np.array_equal(idx.numpy(), g[f"idx_{it}"])
        assert np.array_equal(np.packbits(inc.numpy(), axis=1), g[f"incam_{it}"])
        for b in range(B):
            ci, cc = cref.project(data["pc"][b].numpy(), mean[b].numpy(), pose[b].numpy(), data["K"][b].numpy(), H, W)
            assert np.array_equal(ci, g[f"idx_{it}"][b])
            if full:
                sm = cref.scatter_mean(data["pc_geo_feat"][b].numpy(), data["pc_overlap_pred"][b].numpy(), ci, H * W)
                assert np.array_equal(sm.reshape(64, H, W), g[f"obs2d_proj_{it}"][b])
        if full:
            assert np.array_equal(o2[:, 64:].numpy(), g[f"obs2d_proj_{it}"])
        else:
            assert np.array_equal(o2[:, 64:].double().sum(dim=1).numpy(), g[f"obs2d_proj_chansum_{it}"])
        eo.step(torch.from_numpy(g["a_r"][it]), torch.from_numpy(g["a_t"][it]), pose, cfg)
        rew, dist = eo.reward(pose, data, prev)
        assert np.array_equal(dist.numpy(), g[f"dist_{it}"]) and np.array_equal(rew.numpy(), g[f"reward_{it}"])
        prev = dist
    assert np.array_equal(pose.numpy(), g["pose_final"])


@pytest.mark.parametrize("tag", ["3", "6"])
def test_step_oracle_reproduces_golden(tag):
    from cmr_agent_b200 import environment as drop_in
    g = hp.golden("step")
    cfg = synth.StepConfig(is_6_DoF=(tag == "6"))
    pose = torch.from_numpy(g[f"pose_in_{tag}"]).clone()
    a_r, a_t = torch.from_numpy(g[f"a_r_{tag}"]), torch.from_numpy(g[f"a_t_{tag}"])
    assert np.array_equal(eo.step(a_r, a_t, pose.clone(), cfg).numpy(), g[f"pose_out_{tag}"])
    # host logic of the product (table construction) + C restatement of the kernel arithmetic
    rot, tt = drop_in.build_step_tables(cfg.r_steps, cfg.t_steps)
    nb = tt.shape[0]
    assert rot.shape == (3, nb + 1, 3, 3) and torch.equal(rot[:, nb], torch.eye(3).expand(3, 3, 3))
    for b in range(pose.shape[0]):
        if tag == "6":
            ir = [int(a_r[b, i]) for i in range(3)]
            mv = [float(tt[int(a_t[b, i])]) for i in range(3)]
        else:
            ir = [nb, int(a_r[b, 0]), nb]
            mv = [float(tt[int(a_t[b, 0])]), 0.0, float(tt[int(a_t[b, 1])])]
        Rn = cref.compose_xyz(rot[0, ir[0]].numpy(), rot[1, ir[1]].numpy(), rot[2, ir[2]].numpy())
        got = cref.apply_step(pose[b].numpy(), Rn, np.array(mv, np.float32))
        assert np.array_equal(got, g[f"pose_out_{tag}"][b])


@pytest.mark.parametrize("tag,unique", [("dup", (2600, 2600)), ("plain", None)])
def test_pointnet_oracle_reproduces_golden(tag, unique):
    g = hp.golden("pointnet")
    xyz = synth.make_cloud_batch(2, num_pt=4096, seed=hp.SEED, unique=unique)
    assert hp.sha(xyz) == str(g[f"xyz_sha_{tag}"])
    fps = torch.from_numpy(g[f"fps_{tag}"])
    assert torch.equal(po.farthest_point_sample(xyz, 128, start=fps[:, 0]), fps)
    new_xyz = po.index_points(xyz, fps)
    assert np.array_equal(new_xyz.numpy(), g[f"new_xyz_{tag}"])
    stable = po.knn(new_xyz, xyz, 16)
    assert np.array_equal(stable.numpy(), g[f"knn16_stable_{tag}"])
    assert po.knn_equivalent(torch.from_numpy(g[f"knn16_raw_{tag}"]), stable, new_xyz, xyz)
    for b in range(2):
        assert np.array_equal(cref.fps(xyz[b].numpy(), 128, int(fps[b, 0])), g[f"fps_{tag}"][b])
        assert np.array_equal(cref.knn(new_xyz[b].numpy(), xyz[b].numpy(), 16), g[f"knn16_stable_{tag}"][b])
        for r in (0.5, 2.0):
            assert np.array_equal(cref.ball(new_xyz[b].numpy(), xyz[b].numpy(), r, 32), g[f"ball_{r}_{tag}"][b])
    torch.manual_seed(hp.SEED + 1)
    nx, npts, gxyz, fidx = po.sample_and_group(64, 1.5, 16, xyz, xyz * 0.5 + 1.0, returnfps=True)
    assert np.array_equal(npts.numpy(), g[f"sag_new_points_{tag}"]) and np.array_equal(fidx.numpy(), g[f"sag_fps_{tag}"])


def test_fps_full_size_c_oracle_reproduces_golden():
    g = hp.golden("pointnet")
    xyz = synth.make_cloud_batch(1, num_pt=40960, seed=hp.SEED + 100)
    assert hp.sha(xyz) == str(g["xyz_sha_full"])
    assert np.array_equal(cref.fps(xyz[0].numpy(), 1280, int(g["fps_full"][0, 0])), g["fps_full"][0])
